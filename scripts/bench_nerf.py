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
image, bias = mlp._packed()
t = timeit(lambda: nerf_mlp_forward(image, bias, emb, 63, 27))
flop = rows * 593408 * 2
print(f'fused tcgen05 NerfMLP: {t:.3f} ms for {rows} rows -> {flop / t / 1e9:.1f} TFLOP/s ({rows / t / 1e3:.1f} Mrows/s)')
with torch.enable_grad():
    x = emb[:32768 * 8].clone().requires_grad_(True)
    t2 = timeit(lambda: mlp.batchify_run_mlp(x), n=3, warm=1)
print(f'library fp32 GEMM path: {t2:.3f} ms for {x.shape[0]} rows -> {x.shape[0] * 593408 * 2 / t2 / 1e9:.1f} TFLOP/s')
torch.backends.cuda.matmul.allow_tf32 = True
with torch.enable_grad():
    t3 = timeit(lambda: mlp.batchify_run_mlp(x), n=3, warm=1)
print(f'library TF32 GEMM path: {t3:.3f} ms -> {x.shape[0] * 593408 * 2 / t3 / 1e9:.1f} TFLOP/s')
