"""Developer probe: host-side cost of one NgpRenderer.render / render_fused call (queue not full) and GPU gaps."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xrnerf_b200 import synth, _C
from xrnerf_b200.ngp import NgpField, NgpRenderer
N = 65536
bf, _ = synth.bitfield_from_grid_numpy(synth.lego_like_density_grid(0)); bf = torch.from_numpy(bf).cuda()
f = NgpField().cuda()
o, d, _, _ = synth.ray_batch(N, seed=1); o, d = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
for name in ('render_fused', 'render'):
    r = NgpRenderer(f)
    fn = getattr(r, name)
    for _ in range(5): fn(o, d, bf)
    torch.cuda.synchronize()
    for reps in (5, 50, 200):
        ts = []
        t0 = time.perf_counter()
        for i in range(reps):
            a = time.perf_counter(); fn(o, d, bf); ts.append(time.perf_counter() - a)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        ts = np.array(ts) * 1e3
        print(f'{name}: reps {reps}: host/call mean {ts.mean():.3f} p50 {np.median(ts):.3f} max {ts.max():.3f} ms; issue {1e3*(t1-t0)/reps:.3f} ms/call, incl. drain {1e3*(t2-t0)/reps:.3f} ms/call', flush=True)
# raw C call without the python wrapper work
r = NgpRenderer(f); r.render_fused(o, d, bf); torch.cuda.synchronize()
rgb, alpha, ns = r._out[('fused', N, o.device)]
args = (f.cfg, _C.ptr(f._table16), _C.ptr(f._image), _C.ptr(bf), _C.ptr(o), _C.ptr(d), N, 0.0, 1.0, 0.05, 1 / 256, 9121, 7, _C.float3((0, 0, 0)), 2, 3, _C.ptr(rgb), _C.ptr(alpha), _C.ptr(ns), _C.ptr(r._ws_fused), _C.stream())
t0 = time.perf_counter()
for i in range(200): _C.lib.xrb_ngp_render_fused(*args)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f'raw C call: issue {1e3*(t1-t0)/200:.3f} ms/call, incl. drain {1e3*(t2-t0)/200:.3f} ms/call')
