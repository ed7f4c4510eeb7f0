"""N-GPU check of the peer-memory optimiser exchange (grad_comm='peer') against the NCCL reduce-scatter / all-gather path ('sharded') and the dense all-reduce:
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/peer_probe.py [steps]
Every rank trains on different ray batches from the same initial weights; per mode: ms/step (CUDA events, max over ranks), whether every rank ends with the identical fp16 working table,
and the distance of the final parameters to the 'allreduce' run (the reference's semantics: mean gradient, same Adam on every rank)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xrnerf_b200 import synth  # noqa: E402
from xrnerf_b200.ngp import NgpField  # noqa: E402
from xrnerf_b200.train import NgpTrainer  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
SAME = bool(os.environ.get('PROBE_SAME_GPU'))     # debugging aid: every rank on cuda:0, gloo for the host-side collectives (the peer exchange itself needs no NCCL)
if SAME:
    local = 0
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if SAME:
    dist.init_process_group('gloo')
else:
    dist.init_process_group('nccl', device_id=dev)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
N = 65536
bf = torch.from_numpy(synth.bitfield_from_grid_numpy(synth.lego_like_density_grid(0))[0]).to(dev)
batches = [tuple(torch.from_numpy(x).to(dev) for x in synth.ray_batch(N, seed=100 * rank + b)[:2]) for b in range(4)]
t_, d_, c_ = synth.ngp_weights(seed=0)
g = torch.Generator(device='cpu').manual_seed(7 + rank)
tgt = torch.rand((N, 3), generator=g).to(dev); bg = torch.zeros((N, 3), device=dev)
results = {}
for mode in (('allreduce', 'peer') if SAME else ('allreduce', 'sharded', 'peer')):
    f = NgpField(n_packed_levels=6).to(dev)
    with torch.no_grad():
        f.hash_params.copy_(torch.from_numpy(t_).to(dev)); f.density_params.copy_(torch.from_numpy(d_).to(dev)); f.color_params.copy_(torch.from_numpy(c_).to(dev))
    f.mark_dirty(); f.refresh()
    tr = NgpTrainer(f, bf, N, target_batch_size=1 << 20, grad_comm=mode, ema_momentum=0.05)
    for i in range(3):
        tr.step(*batches[i % 4], tgt, bg, next_rays=batches[(i + 1) % 4])
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        loss = tr.step(*batches[(3 + i) % 4], tgt, bg, next_rays=batches[(4 + i) % 4])
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / K], device='cpu' if SAME else dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if tr.px is not None:
        tr.px.check()
    t16 = f._table16.float()
    sig = torch.stack([t16.double().sum(), t16.double().abs().sum(), (t16.double() * (torch.arange(t16.numel(), device=dev) % 97).double()).sum()])
    sig = sig.cpu() if SAME else sig
    sigs = [torch.zeros_like(sig) for _ in range(world)]
    dist.all_gather(sigs, sig)
    same = all(bool((s == sigs[0]).all()) for s in sigs)
    ema = None
    if not SAME:
        tr.sync_master()
        ema = tr.ema_full()
    results[mode] = dict(ms=float(ms), same=same, table=f.hash_params.detach().clone(), dens=f.density_params.detach().clone(), color=f.color_params.detach().clone(), t16=t16.clone(),
                         ema=None if ema is None else ema.clone(), loss=float(loss))
    tr.close()
    del tr, f
    torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    ref = results['allreduce']
    t0_16 = torch.from_numpy(t_).to(dev).half().float()
    for mode, r in results.items():
        def rel(a, b):
            return float((a - b).norm() / (b.norm() + 1e-30))
        print(f"{mode:9s}: {r['ms']:.3f} ms/step  identical fp16 table on every rank: {r['same']}  loss {r['loss']:.5f}  "
              f"|t16 - t16(allreduce)| / |t16 update| = {float((r['t16'] - ref['t16']).norm() / ((ref['t16'] - t0_16).norm() + 1e-30)):.2e}  rel dens {rel(r['dens'], ref['dens']):.2e}  rel color {rel(r['color'], ref['color']):.2e}"
              f"  rel ema {rel(r['ema'], ref['ema']) if r['ema'] is not None and ref['ema'] is not None else float('nan'):.2e}", flush=True)
dist.destroy_process_group()
