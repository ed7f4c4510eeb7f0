"""Host side of the fused tcgen05 NerfMLP (csrc/nerf_mlp_tc.cu): packs the nn.Linear weights of the reference-shaped NerfMLP
(/root/reference/xrnerf/models/mlps/nerf_mlp.py:29-49) into the kernel's weight image — a sequence of [N x 64] fp16 slabs, K-major,
128-byte swizzled, in the exact order the kernel's TMA producer streams them — and the fp32 bias vector."""
import numpy as np
import torch

from . import _C


def _slab(w_rows_by_64):
    """[N,64] float32 -> swizzled bytes (N*128). element (n,k) -> (n>>3)*1024 + (n&7)*128 + (((k>>3) ^ (n&7))<<4) + (k&7)*2"""
    n = w_rows_by_64.shape[0]
    assert w_rows_by_64.shape[1] == 64 and n % 8 == 0
    h = w_rows_by_64.astype(np.float16)
    out = np.zeros((n // 8, 8, 8, 8), np.float16)          # [row group][row in group][chunk position][elem]
    rows = np.arange(n)
    for c in range(8):
        pos = c ^ (rows & 7)
        out[rows >> 3, rows & 7, pos, :] = h[:, c * 8:(c + 1) * 8]
    return out.reshape(-1).view(np.uint8)


def _pad_cols(w, k0, k1):
    """columns [k0,k1) of w, zero-padded to 64"""
    blk = np.zeros((w.shape[0], 64), np.float32)
    k1 = min(k1, w.shape[1])
    if k1 > k0:
        blk[:, :k1 - k0] = w[:, k0:k1]
    return blk


def pack_nerf_mlp(mlp):
    """mlp: registry NerfMLP (netdepth 8, netwidth 256, skips [4], use_viewdirs). Returns (image u8 cuda, bias f32 cuda)."""
    assert len(mlp.pts_linears) == 8 and list(mlp.skips) == [4] and mlp.use_viewdirs and mlp.pts_linears[0].out_features == 256, \
        'fused NerfMLP kernel: netdepth=8, netwidth=256, skips=[4], use_viewdirs=True'
    ic, icd = mlp.input_ch, mlp.input_ch_dirs
    aux = (ic + 63) // 64
    g = lambda lin: (lin.weight.detach().float().cpu().numpy(), lin.bias.detach().float().cpu().numpy())
    slabs, biases = [], []
    W, b = g(mlp.pts_linears[0])
    for kb in range(aux):
        slabs.append(_slab(_pad_cols(W, kb * 64, (kb + 1) * 64)))
    biases.append(b)
    for l in range(1, 8):
        W, b = g(mlp.pts_linears[l])
        if l == 5:   # input = cat([input_pts(ic), h(256)]) -> image K order [h | pts]
            for kb in range(4):
                slabs.append(_slab(_pad_cols(W, ic + kb * 64, ic + (kb + 1) * 64)))
            for kb in range(aux):
                slabs.append(_slab(_pad_cols(W[:, :ic], kb * 64, (kb + 1) * 64)))
        else:
            for kb in range(4):
                slabs.append(_slab(_pad_cols(W, kb * 64, (kb + 1) * 64)))
        biases.append(b)
    Wf, bf = g(mlp.feature_linear)
    Wa, ba = g(mlp.alpha_linear)
    Wfa = np.zeros((272, 256), np.float32); Wfa[:256] = Wf; Wfa[256] = Wa[0]
    for kb in range(4):
        slabs.append(_slab(_pad_cols(Wfa, kb * 64, (kb + 1) * 64)))
    biases += [bf, np.concatenate([ba, np.zeros(15, np.float32)])]
    Wv, bv = g(mlp.views_linears[0])                     # [128, 256 + icd] on cat([feature, input_views])
    for kb in range(4):
        slabs.append(_slab(_pad_cols(Wv, kb * 64, (kb + 1) * 64)))
    slabs.append(_slab(_pad_cols(Wv[:, 256:], 0, 64)))
    biases.append(bv)
    Wr, br = g(mlp.rgb_linear)
    Wr16 = np.zeros((16, 128), np.float32); Wr16[:3] = Wr
    for kb in range(2):
        slabs.append(_slab(_pad_cols(Wr16, kb * 64, (kb + 1) * 64)))
    biases.append(np.concatenate([br, np.zeros(13, np.float32)]))
    dev = mlp.pts_linears[0].weight.device
    image = torch.from_numpy(np.concatenate(slabs)).to(dev)
    bias = torch.from_numpy(np.concatenate(biases).astype(np.float32)).to(dev)
    return image, bias


def nerf_mlp_forward(image, bias, embedded, input_ch, input_ch_dirs):
    _C.require_cuda(image, bias, embedded)
    embedded = embedded.contiguous().float()
    n = embedded.shape[0]
    raw = torch.empty((n, 4), dtype=torch.float32, device=embedded.device)
    _C.check(_C.lib.xrb_nerf_mlp_forward(_C.ptr(image), _C.ptr(bias), _C.ptr(embedded), n, int(input_ch), int(input_ch_dirs), _C.ptr(raw), _C.stream()), 'nerf_mlp_forward')
    return raw
