"""Developer probe: why is the fused arm slower inside bench.py than in bench_fused.py? Replays bench.py's loop with switches."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from xrnerf_b200 import _C
from xrnerf_b200.ngp import NgpField, NgpRenderer
mode = sys.argv[1] if len(sys.argv) > 1 else 'plain'
dev = torch.device('cuda', 0); torch.cuda.set_device(0)
nbat = 8 if 'fewbatches' in mode else bench.N_BATCHES
bf_np, batches, (table, dens, color) = bench.make_scene(0, nbat)
bf = torch.from_numpy(bf_np).to(dev)
field = NgpField().to(dev)
P = 1
renderers = [NgpRenderer(field, samples_per_ray_budget=bench.BUDGET) for _ in range(P)]
streams = [torch.cuda.Stream(device=dev) for _ in range(P)]
dev_batches = [(torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)) for (o, d) in batches]
if 'pinned' in mode:
    host_batches = [(torch.from_numpy(o).pin_memory(), torch.from_numpy(d).pin_memory()) for (o, d) in batches[:8]]
main = torch.cuda.current_stream()
if 'chain' in mode:
    for i in range(60):
        with torch.cuda.stream(streams[0]):
            renderers[0].render(*dev_batches[i % nbat], bf)
    torch.cuda.synchronize()
if 'smi' in mode:
    c = bench.ClockSampler(0); c.start(); time.sleep(1.0); print(c.stop())
K = 100
for i in range(10):
    with torch.cuda.stream(streams[0]):
        renderers[0].render_fused(*dev_batches[i % nbat], bf)
torch.cuda.synchronize()
f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
f0.record(main)
e = torch.cuda.Event(); e.record(main); streams[0].wait_event(e)
t0 = time.perf_counter()
for i in range(K):
    with torch.cuda.stream(streams[0]):
        renderers[0].render_fused(*dev_batches[(10 + i) % nbat], bf)
t1 = time.perf_counter()
e = torch.cuda.Event(); e.record(streams[0]); main.wait_event(e)
f1.record(main)
torch.cuda.synchronize()
print(f'{mode}: {f0.elapsed_time(f1) / K:.3f} ms/step (host issue {1e3 * (t1 - t0) / K:.3f})')
