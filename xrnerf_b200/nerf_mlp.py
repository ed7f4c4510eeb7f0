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


def pack_nerf_mlp_v2(mlp):
    """Image for csrc/nerf_mlp_tc.cu v2: HALF slabs ([N/2 x 64]) in the kernel's MMA issue order
    S1=(half0; aux+lo blocks) S2=(half1; aux+lo) S3=(half0; hi blocks) S4=(half1; hi); bias vector = per-layer biases, then Wa[256], ba."""
    assert len(mlp.pts_linears) == 8 and list(mlp.skips) == [4] and mlp.use_viewdirs and mlp.pts_linears[0].out_features == 256
    ic, icd = mlp.input_ch, mlp.input_ch_dirs
    aux = (ic + 63) // 64
    g = lambda lin: (lin.weight.detach().float().cpu().numpy(), lin.bias.detach().float().cpu().numpy())
    PTS, DIR = 'pts', 'dir'

    def cols(W, kind, j, h_off=0, pts_w=None):
        """64 weight columns feeding A block (kind, j): hidden block j -> W[:, h_off+64j ..]; pts block j -> pts part; dir -> W[:, 256:]"""
        if kind == 'h':
            return _pad_cols(W, h_off + 64 * j, h_off + 64 * (j + 1))
        if kind == PTS:
            return _pad_cols(pts_w, 64 * j, 64 * (j + 1))
        return _pad_cols(W[:, 256:], 0, 64)

    layers = []   # (W_padded_rows [N, ...], first blocks, second blocks, N, n_halves, bias)
    W, b = g(mlp.pts_linears[0])
    layers.append((W, [(PTS, j, dict(pts_w=W)) for j in range(aux)], [], 256, 2, b))
    for l in range(1, 8):
        W, b = g(mlp.pts_linears[l])
        if l == 5:
            first = [(PTS, j, dict(pts_w=W[:, :ic])) for j in range(aux)] + [('h', 0, dict(h_off=ic)), ('h', 1, dict(h_off=ic))]
            second = [('h', 2, dict(h_off=ic)), ('h', 3, dict(h_off=ic))]
        else:
            first, second = [('h', 0, {}), ('h', 1, {})], [('h', 2, {}), ('h', 3, {})]
        layers.append((W, first, second, 256, 2, b))
    W, b = g(mlp.feature_linear)
    layers.append((W, [('h', 0, {}), ('h', 1, {})], [('h', 2, {}), ('h', 3, {})], 256, 2, b))
    W, b = g(mlp.views_linears[0])
    layers.append((W, [(DIR, 0, {}), ('h', 0, {}), ('h', 1, {})], [('h', 2, {}), ('h', 3, {})], 128, 2, b))
    Wr, br = g(mlp.rgb_linear)
    W16 = np.zeros((16, 128), np.float32); W16[:3] = Wr
    layers.append((W16, [('h', 0, {})], [('h', 1, {})], 16, 1, np.concatenate([br, np.zeros(13, np.float32)])))
    slabs, biases = [], []
    for (W, first, second, N, nh, b) in layers:
        hw = N // nh
        for (kind, j, kw) in first + second:        # stream order: per K-block, output half 0 then half 1 (one ring per issuer)
            for half in range(nh):
                slabs.append(_slab(cols(W, kind, j, **kw)[half * hw:(half + 1) * hw]))
        biases.append(b.astype(np.float32))
    Wa, ba = g(mlp.alpha_linear)
    biases += [Wa[0].astype(np.float32), ba.astype(np.float32)]
    dev = mlp.pts_linears[0].weight.device
    return torch.from_numpy(np.concatenate(slabs)).to(dev), torch.from_numpy(np.concatenate(biases)).to(dev)


def _slab64(w_rows_by_32):
    """[N,32] float32 -> K-major SWIZZLE_64B bytes (N*64): element (n,k) -> (n>>3)*512 + (n&7)*64 + (((k>>3) ^ ((n>>1)&3))<<4) + (k&7)*2"""
    n = w_rows_by_32.shape[0]
    assert w_rows_by_32.shape[1] == 32 and n % 8 == 0
    h = w_rows_by_32.astype(np.float16)
    out = np.zeros((n // 8, 8, 4, 8), np.float16)          # [row group][row in group][chunk position][elem]
    rows = np.arange(n)
    for c in range(4):
        pos = c ^ ((rows >> 1) & 3)
        out[rows >> 3, rows & 7, pos, :] = h[:, c * 8:(c + 1) * 8]
    return out.reshape(-1).view(np.uint8)


def pack_nerf_mlp_v3(mlp):
    """Image for csrc/nerf_mlp_tc3.cu: 16 KB slabs in the kernel's stream order — per layer, per K-block (AUX block first where a layer reads it):
    the 256-wide layers as two K-halves [256 x 32] (SWIZZLE_64B), views_linears.0 / rgb_linear as one [N x 64] slab (SWIZZLE_128B). Bias vector: fp32 per-layer biases, Wa[256], ba, padded to 8 floats, then an fp16 copy of the per-layer biases."""
    assert len(mlp.pts_linears) == 8 and list(mlp.skips) == [4] and mlp.use_viewdirs and mlp.pts_linears[0].out_features == 256
    ic, icd = mlp.input_ch, mlp.input_ch_dirs
    aux = (ic + 63) // 64
    assert aux in (1, 2)
    g = lambda lin: (lin.weight.detach().float().cpu().numpy(), lin.bias.detach().float().cpu().numpy())
    H = lambda W, j, off=0: _pad_cols(W, off + 64 * j, off + 64 * (j + 1))
    layers = []   # (column blocks in K order, N, n_halves, bias)
    W, b = g(mlp.pts_linears[0])
    layers.append(([_pad_cols(W, 64 * j, 64 * (j + 1)) for j in range(aux)], 256, 2, b))          # AUX (pts block 0) [, H3 = pts block 1]
    for l in range(1, 8):
        W, b = g(mlp.pts_linears[l])
        if l == 5:   # input = cat([pts(ic), h(256)])
            Wp = W[:, :ic]
            blocks = [_pad_cols(Wp, 0, 64)] + [H(W, j, ic) for j in range(4)] + ([_pad_cols(Wp, 64, 128)] if aux == 2 else [])
        else:
            blocks = [H(W, j) for j in range(4)]
        layers.append((blocks, 256, 2, b))
    W, b = g(mlp.feature_linear)
    layers.append(([H(W, j) for j in range(4)], 256, 2, b))
    W, b = g(mlp.views_linears[0])                                                                 # [128, 256 + icd] on cat([feature, dirs]); direction block first
    layers.append(([_pad_cols(W[:, 256:], 0, 64)] + [H(W, j) for j in range(4)], 128, 1, b))
    Wr, br = g(mlp.rgb_linear)
    W16 = np.zeros((16, 128), np.float32); W16[:3] = Wr
    layers.append(([H(W16, 0), H(W16, 1)], 16, 1, np.concatenate([br, np.zeros(13, np.float32)])))
    slabs, biases = [], []
    for blocks, N, nh, b in layers:
        hw = N // nh
        for blk in blocks:
            assert blk.shape == (N, 64)
            if nh == 2:     # N=256 layers: two K-halves [256 x 32] in the 64-byte-swizzle layout (one N=256 MMA pair each: A is read once)
                for kh in range(2):
                    slabs.append(_slab64(blk[:, 32 * kh:32 * (kh + 1)]))
            else:
                slabs.append(_slab(blk))
        biases.append(b.astype(np.float32))
    layer_bias = np.concatenate(biases)
    Wa, ba = g(mlp.alpha_linear)
    f32 = np.concatenate([layer_bias, Wa[0].astype(np.float32), ba.astype(np.float32)])
    f32 = np.concatenate([f32, np.zeros((-len(f32)) % 8, np.float32)])
    h16 = layer_bias.astype(np.float16)
    h16 = np.concatenate([h16, np.zeros((-len(h16)) % 8, np.float16)])
    dev = mlp.pts_linears[0].weight.device
    return torch.from_numpy(np.concatenate(slabs)).to(dev), torch.from_numpy(np.concatenate([f32, h16.view(np.float32)])).to(dev)


def nerf_mlp_forward(image, bias, embedded, input_ch, input_ch_dirs, version=1):
    _C.require_cuda(image, bias, embedded)
    embedded = embedded.contiguous().float()
    n = embedded.shape[0]
    raw = torch.empty((n, 4), dtype=torch.float32, device=embedded.device)
    if version >= 2:   # fp32 embedded -> fp16 tile image (one streaming kernel) -> MLP kernel that TMA-loads it
        enc = torch.empty(_C.lib.xrb_nerf_enc_image_bytes(n, int(input_ch)), dtype=torch.uint8, device=embedded.device)
        _C.check(_C.lib.xrb_nerf_pack_embedded(_C.ptr(embedded), n, int(input_ch), int(input_ch_dirs), _C.ptr(enc), _C.stream()), 'nerf_pack_embedded')
        return nerf_mlp_forward_tiles(image, bias, enc, n, input_ch, input_ch_dirs, raw, version=version)
    fn = _C.lib.xrb_nerf_mlp_forward
    _C.check(fn(_C.ptr(image), _C.ptr(bias), _C.ptr(embedded), n, int(input_ch), int(input_ch_dirs), _C.ptr(raw), _C.stream()), 'nerf_mlp_forward')
    return raw


def nerf_mlp_forward_tiles(image, bias, enc_image, n_rows, input_ch, input_ch_dirs, raw=None, version=2):
    """v2 / v3 kernel on an already packed encoding tile image (xrb_nerf_pack_embedded / xrb_nerf_posenc_tiles / xrb_mip_ipe_tiles_rays);
    image/bias must come from the matching packer (pack_nerf_mlp_v2 / pack_nerf_mlp_v3)."""
    if raw is None:
        raw = torch.empty((n_rows, 4), dtype=torch.float32, device=enc_image.device)
    if version == 3:
        _C.check(_C.lib.xrb_nerf_mlp_forward_v3(_C.ptr(image), _C.ptr(bias), _C.ptr(enc_image), n_rows, int(input_ch), int(input_ch_dirs), _C.ptr(raw), _C.stream()), 'nerf_mlp_forward_v3')
        return raw
    _C.check(_C.lib.xrb_nerf_mlp_forward_v2(_C.ptr(image), _C.ptr(bias), _C.ptr(enc_image), n_rows, int(input_ch), int(input_ch_dirs), _C.ptr(raw), _C.stream()), 'nerf_mlp_forward_v2')
    return raw
