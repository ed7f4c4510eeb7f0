"""Whole march (count + scan + emit through rays_sampler_api) on 65 536-ray batches of the bench scene, CUDA events.
python scripts/march_bench.py [clean]   -- `clean`: the same scene without the seeded speckle (a converged grid without floaters)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xrnerf_b200 import raymarch_cuda as rm  # noqa: E402
from xrnerf_b200 import synth  # noqa: E402

clean = len(sys.argv) > 1 and sys.argv[1] == 'clean'
grid = synth.lego_like_density_grid(0, speckle=0.0 if clean else 0.002)
bf, _ = synth.bitfield_from_grid_numpy(grid)
N = 65536
batches = [synth.ray_batch(N, seed=b)[:2] for b in range(8)]
dev = 'cuda'
bft = torch.from_numpy(bf).to(dev)
rays = [(torch.from_numpy(np.ascontiguousarray(o)).to(dev), torch.from_numpy(np.ascontiguousarray(d)).to(dev)) for o, d in batches]
cap = N * 48
coords = torch.zeros((cap, 7), dtype=torch.float32, device=dev)
ridx = torch.zeros((N, 1), dtype=torch.int32, device=dev); ns = torch.zeros((N, 2), dtype=torch.int32, device=dev); cnt = torch.zeros(2, dtype=torch.int32, device=dev)


def run(o, d):
    cnt.zero_()
    rm.rays_sampler_api(o, d, bft, None, None, None, 0.0, 1.0, 0.05, 1.0 / 256, coords, ridx, ns, cnt)


rm.reset_rng(ray_sampler=0)
for o, d in rays[:3]:
    run(o, d)
torch.cuda.synchronize()
ts = []
for rep in range(5):
    for o, d in rays:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(o, d); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
rm.reset_rng(ray_sampler=0)
run(*rays[0]); torch.cuda.synchronize()
sig = (int(cnt[1]), int(ns[:, 0].sum()), float(coords[:int(cnt[1])].double().sum()))
print(f'scene={"clean" if clean else "speckled"}: march (count + scan + emit) median {np.median(ts):.1f} us, '
      f'min {np.min(ts):.1f} us; samples/ray {sig[0] / N:.2f}; signature {sig}')
