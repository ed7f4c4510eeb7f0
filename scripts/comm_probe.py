"""2+ ranks: time the gradient-exchange collectives alone and the training step with / without them.  torchrun --nproc-per-node N scripts/comm_probe.py"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
dev = torch.device('cuda', local)
n = 12196240
per = -(-(-(-n // world)) // 8) * 8
pad = per * world
g32 = torch.randn(pad, device=dev); g16 = g32.bfloat16(); mine16 = torch.empty(per, dtype=torch.bfloat16, device=dev)
t16 = torch.zeros(pad, dtype=torch.float16, device=dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, (time.perf_counter() - t0) / reps * 1e6


res = {}
res['all_reduce fp32 48.8MB'] = timed(lambda: dist.all_reduce(g32))
res['reduce_scatter bf16 24.4MB'] = timed(lambda: dist.reduce_scatter_tensor(mine16, g16))
res['all_gather fp16 24.4MB'] = timed(lambda: dist.all_gather_into_tensor(t16, t16[rank * per:(rank + 1) * per]))
res['all_reduce fp32 41KB (MLP weights)'] = timed(lambda: dist.all_reduce(g32[:10240]))
if rank == 0:
    for k, (dev_us, wall_us) in res.items():
        print('%-40s device %.1f us  wall %.1f us per call' % (k, dev_us, wall_us), flush=True)

from xrnerf_b200 import synth
from xrnerf_b200.ngp import NgpField
from xrnerf_b200.train import NgpTrainer
N = 65536
bf = torch.from_numpy(synth.bitfield_from_grid_numpy(synth.lego_like_density_grid(0))[0]).to(dev)
batches = [tuple(torch.from_numpy(x).to(dev) for x in synth.ray_batch(N, seed=1000 * rank + b)[:2]) for b in range(4)]
tgt = torch.rand((N, 3), device=dev); bg = torch.zeros((N, 3), device=dev)
for mode in ('none', 'allreduce', 'sharded'):
    f = NgpField().to(dev)
    tr = NgpTrainer(f, bf, N, target_batch_size=1 << 20, grad_comm=mode)
    if mode == 'none':
        tr.grad_comm = 'none'
    step = lambda i=[0]: (tr.step(*batches[i[0] % 4], tgt, bg, next_rays=batches[(i[0] + 1) % 4]), i.__setitem__(0, i[0] + 1))
    d, w = timed(step, reps=20)
    if rank == 0:
        print('train step grad_comm=%-9s device %.1f us  wall %.1f us' % (mode, d, w), flush=True)
    del tr, f
dist.destroy_process_group()
