"""The peer-memory optimiser exchange (csrc/peer_adam.cu, NgpTrainer(grad_comm='peer')) with TWO processes: CUDA IPC, the flag protocol and the fused reduce / Adam / all-gather
kernel are the real ones. The ranks share cuda:0 when the box has one GPU (IPC between processes works on one device too; host-side collectives then go through gloo), and use
cuda:0 / cuda:1 over NCCL when it has two. Checked: every rank ends with the bit-identical fp16 working table, and parameters agree with the dense all-reduce path (the reference's
DDP semantics) up to the bf16 rounding of the exchanged gradients."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, two_gpus, out):
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    local = rank if two_gpus else 0
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if two_gpus:
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group('gloo', rank=rank, world_size=world)
    from xrnerf_b200 import synth
    from xrnerf_b200.ngp import NgpField
    from xrnerf_b200.train import NgpTrainer
    n = 8192
    bf = torch.from_numpy(synth.bitfield_from_grid_numpy(synth.lego_like_density_grid(0))[0]).to(dev)
    batches = [tuple(torch.from_numpy(x).to(dev) for x in synth.ray_batch(n, seed=100 * rank + b)[:2]) for b in range(2)]
    t_, d_, c_ = synth.ngp_weights(seed=0)
    tgt = torch.rand((n, 3), generator=torch.Generator().manual_seed(3 + rank)).to(dev); bg = torch.zeros((n, 3), device=dev)
    res = {}
    for mode in ('allreduce', 'peer'):
        f = NgpField(n_packed_levels=6).to(dev)
        with torch.no_grad():
            f.hash_params.copy_(torch.from_numpy(t_).to(dev)); f.density_params.copy_(torch.from_numpy(d_).to(dev)); f.color_params.copy_(torch.from_numpy(c_).to(dev))
        f.mark_dirty(); f.refresh()
        tr = NgpTrainer(f, bf, n, target_batch_size=1 << 18, grad_comm=mode, ema_momentum=0.05)
        assert tr.grad_comm == mode
        for i in range(4):
            tr.step(*batches[i % 2], tgt, bg)
        torch.cuda.synchronize()
        if tr.px is not None:
            tr.px.check()
        t16 = f._table16.float().clone()
        sig = torch.stack([t16.double().sum(), t16.double().abs().sum(), (t16.double() * (torch.arange(t16.numel(), device=dev) % 97).double()).sum()]).cpu()
        sigs = [None] * world
        dist.all_gather_object(sigs, sig.tolist())
        res[mode] = dict(t16=t16.cpu().numpy(), dens=f.density_params.detach().cpu().numpy(), color=f.color_params.detach().cpu().numpy(), same=all(s == sigs[0] for s in sigs))
        tr.close()
        assert f._table16.data_ptr() != 0 and float((f._table16.float() - t16).abs().max()) == 0.0     # the field owns an ordinary copy again
        del tr, f
    t0 = torch.from_numpy(t_).half().float().numpy()
    upd = np.linalg.norm(res['allreduce']['t16'] - t0)
    out[rank] = dict(same_allreduce=res['allreduce']['same'], same_peer=res['peer']['same'], upd=float(upd),
                     rel_t16=float(np.linalg.norm(res['peer']['t16'] - res['allreduce']['t16']) / (upd + 1e-30)),
                     rel_dens=float(np.linalg.norm(res['peer']['dens'] - res['allreduce']['dens']) / np.linalg.norm(res['allreduce']['dens'])),
                     rel_color=float(np.linalg.norm(res['peer']['color'] - res['allreduce']['color']) / np.linalg.norm(res['allreduce']['color'])))
    dist.barrier()
    dist.destroy_process_group()


def test_peer_exchange_two_processes():
    import torch.multiprocessing as mp
    two = torch.cuda.device_count() >= 2
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, two, out), nprocs=2, join=True)
    assert len(out) == 2
    for r in range(2):
        o = out[r]
        assert o['same_peer'], 'the ranks hold different fp16 tables after the peer exchange'
        assert o['upd'] > 0 and o['rel_t16'] < 5e-2 and o['rel_dens'] < 1e-2 and o['rel_color'] < 1e-2, o
