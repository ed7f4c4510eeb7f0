"""Developer timing probe (not the contract bench): per-stage device times of the NGP path on the synthetic scene."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from xrnerf_b200 import synth
from xrnerf_b200.ngp import NgpField, NgpRenderer
import xrnerf_b200.raymarch_cuda as rm


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    impls = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else '0,1').split(',')]
    N = 65536
    grid = synth.lego_like_density_grid(0)
    bf, _ = synth.bitfield_from_grid_numpy(grid)
    o, d, img, poses = synth.ray_batch(N, seed=1)
    o, d, bf = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), torch.from_numpy(bf).cuda()
    f = NgpField().cuda()
    cap = N * 64
    coords = torch.zeros((cap, 7), device='cuda'); ridx = torch.zeros((N, 1), dtype=torch.int32, device='cuda')
    ns = torch.zeros((N, 2), dtype=torch.int32, device='cuda'); cnt = torch.zeros(2, dtype=torch.int32, device='cuda')

    def march():
        cnt.zero_()
        rm.rays_sampler_api(o, d, bf, None, None, None, 0.0, 1.0, 0.05, 1 / 256, coords, ridx, ns, cnt)
    t = timeit(march)
    S = int(cnt[1].item())
    print(f'march: {t:.3f} ms  rays={N} samples={S} ({S / N:.1f}/ray)', flush=True)
    c = coords[:S]
    for impl in impls:
        t = timeit(lambda: f.run_mlp(c[:, :3], c[:, 4:], impl=impl), n=5, warm=2)
        print(f'field impl{impl}: {t:.3f} ms  {S / t / 1e3:.1f} Msamples/s  gather {S * 512 / t / 1e6:.1f} GB/s', flush=True)
    raw = f.run_mlp(c[:, :3], c[:, 4:], impl=impls[-1])
    rgb = torch.zeros((N, 3), device='cuda'); alpha = torch.zeros((N, 1), device='cuda')
    t = timeit(lambda: rm.calc_rgb_influence_api(raw, c, ns, torch.zeros(3), 2, 3, 0.0, 1.0, rgb, alpha))
    print(f'composite: {t:.3f} ms', flush=True)
    if 1 in impls:
        r = NgpRenderer(f)
        t = timeit(lambda: r.render(o, d, bf))
        print(f'render (fused chain): {t:.3f} ms  {N / t / 1e3:.2f} Mrays/s', flush=True)


if __name__ == '__main__':
    main()
