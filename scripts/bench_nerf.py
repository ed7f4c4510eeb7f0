"""Developer probe: tensor-pipe throughput of the fused tcgen05 NerfMLP kernel vs the fp32 library-GEMM path (config-3 shape)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xrnerf_b200 import registry as R
from xrnerf_b200.nerf_mlp import nerf_mlp_forward

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
rows = 32768 * 64            # 32 768 rays x 64 samples
emb = torch.randn((rows, 90), device='cuda')
flop = rows * 593408 * 2
from xrnerf_b200.nerf_mlp import pack_nerf_mlp, pack_nerf_mlp_v2
for v, packer in (((2, pack_nerf_mlp_v2),) if os.environ.get('XRB_NM_DBG') else ((1, pack_nerf_mlp), (2, pack_nerf_mlp_v2))):
    image, bias = packer(mlp)
    t = timeit(lambda: nerf_mlp_forward(image, bias, emb, 63, 27, version=v))
    print(f'fused tcgen05 NerfMLP v{v}: {t:.3f} ms for {rows} rows -> {flop / t / 1e9:.1f} TFLOP/s ({rows / t / 1e3:.1f} Mrows/s)', flush=True)
    if v == 2:
        from xrnerf_b200 import _C
        from xrnerf_b200.nerf_mlp import nerf_mlp_forward_tiles
        enc = torch.empty(_C.lib.xrb_nerf_enc_image_bytes(rows, 63), dtype=torch.uint8, device='cuda')
        _C.check(_C.lib.xrb_nerf_pack_embedded(_C.ptr(emb), rows, 63, 27, _C.ptr(enc), _C.stream()))
        raw = torch.empty((rows, 4), device='cuda')
        t = timeit(lambda: nerf_mlp_forward_tiles(image, bias, enc, rows, 63, 27, raw))
        print(f'   v2 MLP kernel alone (pre-packed encodings): {t:.3f} ms -> {flop / t / 1e9:.1f} TFLOP/s', flush=True)
if os.environ.get('XRB_NM_DBG'):
    sys.exit(0)
with torch.enable_grad():
    x = emb[:32768 * 8].clone().requires_grad_(True)
    t2 = timeit(lambda: mlp.batchify_run_mlp(x), n=3, warm=1)
print(f'library fp32 GEMM path: {t2:.3f} ms for {x.shape[0]} rows -> {x.shape[0] * 593408 * 2 / t2 / 1e9:.1f} TFLOP/s')
torch.backends.cuda.matmul.allow_tf32 = True
with torch.enable_grad():
    t3 = timeit(lambda: mlp.batchify_run_mlp(x), n=3, warm=1)
print(f'library TF32 GEMM path: {t3:.3f} ms -> {x.shape[0] * 593408 * 2 / t3 / 1e9:.1f} TFLOP/s')
