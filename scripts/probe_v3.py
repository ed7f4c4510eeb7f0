"""Developer probe: attribution of the NerfMLP v3 kernel time (XRB_NM_DBG bits: 1 no weight TMA, 2 no MMA, 4 no epilogue math) and v2 for comparison."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xrnerf_b200 import registry as R, _C
from xrnerf_b200.nerf_mlp import nerf_mlp_forward_tiles, pack_nerf_mlp_v2, pack_nerf_mlp_v3

MLP = dict(type='NerfMLP', skips=[4], netdepth=8, netwidth=256, netchunk=1024 * 32, output_ch=5, use_viewdirs=True, embedder=dict(type='BaseEmbedder', i_embed=0, multires=10, multires_dirs=4))


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


mlp = R.build_mlp(MLP).cuda()
rows = 32768 * 64
emb = torch.randn((rows, 90), device='cuda')
flop = rows * 593408 * 2
enc = torch.empty(_C.lib.xrb_nerf_enc_image_bytes(rows, 63), dtype=torch.uint8, device='cuda')
_C.check(_C.lib.xrb_nerf_pack_embedded(_C.ptr(emb), rows, 63, 27, _C.ptr(enc), _C.stream()))
raw = torch.empty((rows, 4), device='cuda')
packs = {2: pack_nerf_mlp_v2(mlp), 3: pack_nerf_mlp_v3(mlp)}
only = os.environ.get('PROBE_ONLY')
for v in ((3,) if os.environ.get('PROBE_V3_ONLY') else (3, 2)):
    for dbg in ([int(only)] if only else [0, 1, 2, 4, 7]):
        os.environ['XRB_NM_DBG'] = str(dbg)
        image, bias = packs[v]
        t = timeit(lambda: nerf_mlp_forward_tiles(image, bias, enc, rows, 63, 27, raw, version=v))
        print(f'v{v} dbg={dbg} (tma {"off" if dbg & 1 else "on "} mma {"off" if dbg & 2 else "on "} epi {"off" if dbg & 4 else "on "}): {t:.3f} ms -> {flop / t / 1e9:.1f} TFLOP/s-equivalent', flush=True)
    if only:
        break

for sr in ('0', '1'):
    os.environ['XRB_N3_SHARED_RING'] = sr; os.environ['XRB_NM_DBG'] = '0'
    image, bias = packs[3]
    t = timeit(lambda: nerf_mlp_forward_tiles(image, bias, enc, rows, 63, 27, raw, version=3), n=10)
    print(f'v3 shared_ring={sr}: {t:.3f} ms -> {flop / t / 1e9:.1f} TFLOP/s-equivalent', flush=True)
os.environ['XRB_N3_SHARED_RING'] = os.environ.get('PROBE_SR', '0')
for stg in ():
    os.environ['XRB_NM_DBG'] = '0'; os.environ['XRB_N3_STAGGER'] = str(stg)
    image, bias = packs[3]
    t = timeit(lambda: nerf_mlp_forward_tiles(image, bias, enc, rows, 63, 27, raw, version=3), n=10)
    print(f'v3 stagger={stg}: {t:.3f} ms -> {flop / t / 1e9:.1f} TFLOP/s-equivalent', flush=True)
os.environ['XRB_N3_STAGGER'] = '3000'
# ---- timeline of one tile (dbg bit4): issuer / poller / compute time stamps per layer, in cycles relative to the layer-0 issuer start
import ctypes, numpy as np
_C.lib.xrb_internal_n3_trace.argtypes = [ctypes.c_void_p]
for dbg, stg in [(int(x), 0) for x in os.environ.get('PROBE_TIMELINES', '16').split(',')]:
    os.environ['XRB_N3_STAGGER'] = str(stg)
    os.environ['XRB_NM_DBG'] = str(dbg)
    image, bias = packs[3]
    nerf_mlp_forward_tiles(image, bias, enc, rows, 63, 27, raw, version=3); torch.cuda.synchronize()
    buf = np.zeros((8, 16), np.int64)
    _C.lib.xrb_internal_n3_trace(buf.ctypes.data)
    t0 = buf[0, 0]
    print(f'--- timeline dbg={dbg}: per layer [P0 issuer start, P0 committed, P0 poller woke, P0 released, P1 issuer start, P0 epi done (w0), P0 epi done (w7), P1 committed] (cycles since layer-0 start)')
    for l in range(11):
        print(f'  L{l:2d} ' + ' '.join(f'{int(buf[e, l] - t0):7d}' for e in range(8)))
