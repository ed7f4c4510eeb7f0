"""GPU parity of the fused tcgen05 NerfMLP kernel (xrb_nerf_mlp_forward) against the fp32 path of the same module (library GEMMs, the
reference's arithmetic) and the numpy oracle. Tolerance for the fp16-operand / fp32-accumulate pipeline over 12 layers:
|err| <= 2e-2 * max|raw| (measured ~3e-3); composited rgb <= 5e-3."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NERF_MLP = dict(type='NerfMLP', skips=[4], netdepth=8, netwidth=256, netchunk=1024 * 32, output_ch=5, use_viewdirs=True,
                embedder=dict(type='BaseEmbedder', i_embed=0, multires=10, multires_dirs=4))
MIP_MLP = dict(type='NerfMLP', skips=[4], netdepth=8, netwidth=256, netchunk=1024 * 32, use_viewdirs=True,
               embedder=dict(type='MipNerfEmbedder', min_deg_point=0, max_deg_point=16, min_deg_view=0, max_deg_view=4, use_viewdirs=True, append_identity=True))


def _ref_fp32(mlp, emb):
    with torch.enable_grad():   # autograd on => library-GEMM fp32 path
        return mlp.batchify_run_mlp(emb.clone().requires_grad_(True)).detach()


@pytest.mark.parametrize('n_rays,s', [(1, 1), (1, 127), (3, 43), (40, 64), (500, 192)])
def test_fused_nerf_mlp_matches_fp32_path(n_rays, s):
    from xrnerf_b200 import registry as R
    torch.manual_seed(0)
    mlp = R.build_mlp(NERF_MLP).cuda()
    pts = torch.rand((n_rays, s, 3), device='cuda') * 4 - 2
    vd = torch.nn.functional.normalize(torch.randn((n_rays, 3), device='cuda'), dim=-1)
    with torch.no_grad():
        data = mlp({'pts': pts, 'viewdirs': vd})
        emb = mlp.embedder({'pts': pts, 'viewdirs': vd})['embedded']
    raw = data['raw']
    assert raw.shape == (n_rays, s, 4) and torch.isfinite(raw).all()
    ref = _ref_fp32(mlp, emb).reshape(n_rays, s, 4)
    err = (raw - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item() + 1e-3, err


def test_fused_nerf_mlp_vs_numpy_oracle_and_render():
    from oracle import nerf_oracle as O
    from xrnerf_b200 import registry as R
    torch.manual_seed(1)
    mlp = R.build_mlp(NERF_MLP).cuda()
    render = R.build_render(dict(type='NerfRender', white_bkgd=True, raw_noise_std=0))
    rng = np.random.default_rng(0)
    n, s = 64, 64
    o = (rng.random((n, 3)) * 0.2).astype(np.float32); d = rng.normal(0, 1, (n, 3)).astype(np.float32)
    vd = d / np.linalg.norm(d, axis=-1, keepdims=True)
    z = np.broadcast_to(np.linspace(2, 6, s, dtype=np.float32), (n, s)).copy()
    pts = o[:, None] + d[:, None] * z[..., None]
    sd = {k: v.detach().cpu().numpy() for k, v in mlp.state_dict().items()}
    raw_ref = O.nerf_mlp(sd, O.embed(pts, vd), 63, 27).reshape(n, s, 4)
    rgb_ref = O.nerf_render(raw_ref, z, d, white_bkgd=True)['rgb']
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    with torch.no_grad():
        data = mlp({'pts': t(pts), 'viewdirs': t(vd)})
        data.update(z_vals=t(z), rays_d=t(d))
        _, ret = render(data, is_test=True)
    assert np.abs(data['raw'].cpu().numpy() - raw_ref).max() <= 2e-2 * np.abs(raw_ref).max() + 1e-3
    assert np.abs(ret['rgb'].cpu().numpy() - rgb_ref).max() <= 5e-3


def test_fused_mip_mlp_matches_fp32_path():
    from xrnerf_b200 import registry as R
    torch.manual_seed(2)
    mlp = R.build_mlp(MIP_MLP).cuda()
    n, s1 = 300, 129
    d = torch.randn((n, 3), device='cuda')
    data = dict(rays_o=torch.rand((n, 3), device='cuda') * 0.2, rays_d=d, viewdirs=torch.nn.functional.normalize(d, dim=-1), radii=torch.full((n, 1), 1e-3, device='cuda'),
                z_vals=torch.linspace(2, 6, s1, device='cuda').expand(n, s1).contiguous())
    with torch.no_grad():
        emb = mlp.embedder(dict(data))['embedded']
        raw = mlp(dict(data))['raw']
    assert raw.shape == (n, s1 - 1, 4)
    ref = _ref_fp32(mlp, emb).reshape(n, s1 - 1, 4)
    assert (raw - ref).abs().max().item() <= 2e-2 * ref.abs().max().item() + 1e-3


def test_weight_repack_on_update():
    from xrnerf_b200 import registry as R
    torch.manual_seed(3)
    mlp = R.build_mlp(NERF_MLP).cuda()
    pts = torch.rand((8, 16, 3), device='cuda'); vd = torch.nn.functional.normalize(torch.randn((8, 3), device='cuda'), dim=-1)
    with torch.no_grad():
        a = mlp({'pts': pts, 'viewdirs': vd})['raw'].clone()
        mlp.rgb_linear.bias.add_(1.0)
        b = mlp({'pts': pts, 'viewdirs': vd})['raw']
    assert torch.allclose(b[..., :3], a[..., :3] + 1.0, atol=1e-5) and torch.equal(a[..., 3], b[..., 3])


MIP_MODEL = dict(type='MipNerfNetwork', cfg=dict(phase='test', ray_shape='cone', resample_padding=0.01, use_multiscale=False, coarse_loss_mult=0.1, num_levels=2, chunk=800, bs_data='rays_o'),
                 mlp=MIP_MLP, render=dict(type='MipNerfRender', white_bkgd=True, raw_noise_std=0, rgb_padding=0.001, density_bias=-1, density_activation='softplus'))


def _mip_rays(n, seed):
    g = torch.Generator(device='cuda').manual_seed(seed)
    d = torch.randn((n, 3), device='cuda', generator=g)
    o = torch.rand((n, 3), device='cuda', generator=g) * 0.2
    return o, d, torch.nn.functional.normalize(d, dim=-1), torch.full((n, 1), 1.2e-3, device='cuda')


def test_mip_ipe_tile_image_is_the_fp16_rounding_of_mip_embed(monkeypatch):
    """xrb_mip_ipe_tiles_rays == xrb_nerf_pack_embedded(xrb_mip_embed(...)) byte for byte (same fp32 expressions, one fp16 rounding); ragged last tile."""
    from xrnerf_b200 import _C
    n, s = 37, 19                         # 703 rows: 5 full tiles + a ragged one
    o, d, vd, radii = _mip_rays(n, 4)
    z = (2.0 + 4.0 * torch.rand((n, s + 1), device='cuda')).sort(dim=-1).values.contiguous()
    emb = torch.empty((n * s, 123), device='cuda')
    _C.check(_C.lib.xrb_mip_embed(_C.ptr(z), _C.ptr(o), _C.ptr(d), _C.ptr(radii.reshape(-1)), _C.ptr(vd), n, s, 0, 16, 0, 4, _C.ptr(emb), None, None, _C.stream()))
    nbytes = _C.lib.xrb_nerf_enc_image_bytes(n * s, 96)
    a = torch.zeros(nbytes, dtype=torch.uint8, device='cuda'); b = torch.full((nbytes,), 0xA5, dtype=torch.uint8, device='cuda')
    _C.check(_C.lib.xrb_nerf_pack_embedded(_C.ptr(emb), n * s, 96, 27, _C.ptr(a), _C.stream()))
    _C.check(_C.lib.xrb_mip_ipe_tiles_rays(_C.ptr(z), _C.ptr(o), _C.ptr(d), _C.ptr(radii.reshape(-1)), _C.ptr(vd), n, s, 0, 16, 0, 4, _C.ptr(b), _C.stream()))
    assert nbytes == 6 * 3 * 16384
    assert torch.equal(a.view(torch.float16), b.view(torch.float16))      # as fp16 VALUES: where exp(-y_var/2) < 2^-25 the specialised kernel writes +0 without evaluating sin (may be -0 above)
    monkeypatch.setenv('XRB_GENERIC_ENCODERS', '1')                        # the generic-degree kernel: byte for byte
    c = torch.full((nbytes,), 0x5A, dtype=torch.uint8, device='cuda')
    _C.check(_C.lib.xrb_mip_ipe_tiles_rays(_C.ptr(z), _C.ptr(o), _C.ptr(d), _C.ptr(radii.reshape(-1)), _C.ptr(vd), n, s, 0, 16, 0, 4, _C.ptr(c), _C.stream()))
    assert torch.equal(a, c)


def test_posenc_tile_images_match_fp32_posenc(monkeypatch):
    """xrb_nerf_posenc_tiles / _rays (specialised multires 10/4 kernel and the generic one) == fp16 rounding of xrb_nerf_posenc on pts = o + d*z formed
    by torch (GetPts). The specialised kernel uses sincosf where the fp32 kernel uses sinf / cosf: allow 1 fp16 ulp on < 0.1 % of the entries."""
    from xrnerf_b200 import _C
    n, s = 29, 21
    g = torch.Generator(device='cuda').manual_seed(9)
    o = torch.rand((n, 3), device='cuda', generator=g); d = torch.randn((n, 3), device='cuda', generator=g)
    vd = torch.nn.functional.normalize(d, dim=-1)
    z = (2.0 + 4.0 * torch.rand((n, s), device='cuda', generator=g)).sort(dim=-1).values.contiguous()
    pts = (o[:, None, :] + d[:, None, :] * z[:, :, None]).contiguous()
    emb = torch.empty((n * s, 90), device='cuda')
    _C.check(_C.lib.xrb_nerf_posenc(_C.ptr(pts), _C.ptr(vd), n * s, s, 10, 4, _C.ptr(emb), _C.stream()))
    nbytes = _C.lib.xrb_nerf_enc_image_bytes(n * s, 63)
    ref = torch.zeros(nbytes, dtype=torch.uint8, device='cuda')
    _C.check(_C.lib.xrb_nerf_pack_embedded(_C.ptr(emb), n * s, 63, 27, _C.ptr(ref), _C.stream()))
    ref16 = ref.view(torch.float16).float()
    for generic in (False, True):
        if generic:
            monkeypatch.setenv('XRB_GENERIC_ENCODERS', '1')
        for ray_mode in (False, True):
            img = torch.full((nbytes,), 0xA5, dtype=torch.uint8, device='cuda')
            if ray_mode:
                _C.check(_C.lib.xrb_nerf_posenc_tiles_rays(_C.ptr(o), _C.ptr(d), _C.ptr(z), _C.ptr(vd), n, s, 10, 4, _C.ptr(img), _C.stream()))
            else:
                _C.check(_C.lib.xrb_nerf_posenc_tiles(_C.ptr(pts), _C.ptr(vd), n * s, s, 10, 4, _C.ptr(img), _C.stream()))
            got = img.view(torch.float16).float()
            diff = (got - ref16).abs()
            assert diff.max().item() <= 1e-3, (generic, ray_mode)             # one fp16 ulp below 1.0 is 4.9e-4
            assert (diff > 0).float().mean().item() <= (0.0 if generic else 1e-3), (generic, ray_mode)


def test_fused_mip_renderer_matches_registry_network():
    """MipNerfRenderer (IPE tile images -> tcgen05 MLP -> composite, 2 levels) vs MipNerfNetwork.forward(is_test=True) with the fp32 library-GEMM MLP.
    Tolerance: fp16-operand MLP over 12 layers, then two composites and a resampling that moves with the coarse weights: rgb 1e-2."""
    from xrnerf_b200 import registry as R
    from xrnerf_b200.nerf import MipNerfRenderer
    torch.manual_seed(5)
    net = R.build_network(MIP_MODEL).cuda()
    n, S = 200, 128
    o, d, vd, radii = _mip_rays(n, 6)
    with torch.no_grad():
        got = MipNerfRenderer(net, near=2.0, far=6.0, n_samples=S).render(o, d, vd, radii)
        t = torch.linspace(0., 1., S + 1, device='cuda')
        data = dict(rays_o=o, rays_d=d, viewdirs=vd, radii=radii, z_vals=(2.0 * (1. - t) + 6.0 * t).expand(n, S + 1).contiguous())
        net.mlp.fused = False
        ref = net.forward(data, is_test=True)
    assert set(got) == set(ref) == {'rgb', 'disp', 'acc', 'coarse_rgb', 'coarse_disp', 'coarse_acc'}
    for k in ('rgb', 'coarse_rgb', 'acc', 'coarse_acc'):
        assert (got[k] - ref[k]).abs().max().item() <= 1e-2, k


@pytest.mark.parametrize('cfg,n_rows', [(NERF_MLP, 1), (NERF_MLP, 129), (NERF_MLP, 2 * 296 * 128 + 77), (MIP_MLP, 128), (MIP_MLP, 2 * 296 * 128 + 300)])
def test_nerf_mlp_v3_two_tiles_in_flight_matches_v2_and_fp32(cfg, n_rows, monkeypatch):
    """csrc/nerf_mlp_tc3.cu (two tile pipelines per SM, AUX block time-shared by point / direction encodings, several tiles per pipeline at the
    larger sizes) against v2 (same fp16 arithmetic in the same K order: <= 1e-3 of max|raw|, normally bit-identical) and the fp32 library path (2e-2)."""
    from xrnerf_b200 import registry as R
    from xrnerf_b200.nerf_mlp import nerf_mlp_forward
    torch.manual_seed(7)
    mlp = R.build_mlp(cfg).cuda()
    emb = torch.randn((n_rows, mlp.input_ch + mlp.input_ch_dirs), device='cuda').clamp_(-1, 1)
    out = {}
    for v in ('2', '3'):
        monkeypatch.setenv('XRB_NERF_MLP_V', v)
        image, bias = mlp._packed()
        out[v] = nerf_mlp_forward(image, bias, emb, mlp.input_ch, mlp.input_ch_dirs, version=int(v)).clone()
    torch.cuda.synchronize()
    scale = out['2'].abs().max().item()
    assert (out['3'] - out['2']).abs().max().item() <= 1e-3 * scale + 1e-6
    with torch.no_grad():
        ref = _ref_fp32(mlp, emb[:4096])
    assert (out['3'][:4096] - ref).abs().max().item() <= 2e-2 * ref.abs().max().item() + 1e-3


@pytest.mark.parametrize('cfg,n_rows', [(NERF_MLP, 1), (NERF_MLP, 5 * 128 + 3), (NERF_MLP, 2 * 296 * 128 + 77), (MIP_MLP, 2 * 296 * 128 + 300)])
def test_nerf_mlp_v3_cta_pair_multicast_matches_single_cta(cfg, n_rows, monkeypatch):
    """XRB_N3_CLUSTER=1: clusters of two CTAs, rank 0 loads every weight slab once and the TMA multicasts it into both CTAs' rings (phantom tiles keep the
    pair in lock step at the ragged end). Same arithmetic, same order: identical results."""
    from xrnerf_b200 import registry as R
    from xrnerf_b200.nerf_mlp import nerf_mlp_forward
    torch.manual_seed(11)
    mlp = R.build_mlp(cfg).cuda()
    emb = torch.randn((n_rows, mlp.input_ch + mlp.input_ch_dirs), device='cuda').clamp_(-1, 1)
    monkeypatch.setenv('XRB_NERF_MLP_V', '3')
    image, bias = mlp._packed()
    out = {}
    monkeypatch.setenv('XRB_N3_SHARED_RING', '0')             # private 2-slot rings (MD 0) / CTA pairs (MD 1)
    for cl in ('0', '1'):
        monkeypatch.setenv('XRB_N3_CLUSTER', cl)
        out[cl] = nerf_mlp_forward(image, bias, emb, mlp.input_ch, mlp.input_ch_dirs, version=3).clone()
    monkeypatch.setenv('XRB_N3_CLUSTER', '0')
    monkeypatch.setenv('XRB_N3_SHARED_RING', '1')             # one 4-slot ring consumed in the global order P0.layer, P1.layer, ... (csrc/nerf_mlp_tc3.cu, MD 2)
    out['sr'] = nerf_mlp_forward(image, bias, emb, mlp.input_ch, mlp.input_ch_dirs, version=3).clone()
    torch.cuda.synchronize()
    assert torch.equal(out['0'], out['1']) and torch.equal(out['0'], out['sr'])
