"""GPU parity of the fused NGP render (xrb_ngp_render: march -> tcgen05 field -> composite, no host sync) against the
oracle chain (reference march kernels on CPU -> tcnn restatement -> reference compositing kernel on CPU).
numsteps: bit-exact. rgb/alpha: 2e-3 abs (fp16 field, see test_gpu_field.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope='module')
def setup():
    from xrnerf_b200 import synth
    from xrnerf_b200.ngp import NgpField
    f = NgpField().cuda()
    table, dens, color = synth.ngp_weights(seed=3, hash_range=0.5, mlp_gain=2.0)
    with torch.no_grad():
        f.hash_params.copy_(dev(table)); f.density_params.copy_(dev(dens)); f.color_params.copy_(dev(color))
    return f, table, dens, color


def oracle_render(port, ref, scene, table, dens, color, o, d, bg, n_prior=0):
    c, _, ns, cnt = (ref or port).rays_sampler(o, d, scene['bitfield'], o.shape[0] * 64, n_prior_calls=n_prior)
    coords = c[:cnt[1]]
    raw = port.ngp_mlp_forward(table, dens, color, np.ascontiguousarray(coords[:, :3]), np.ascontiguousarray(coords[:, 4:]))
    rgb, alpha = (ref or port).calc_rgb_inference(raw, coords, ns, np.asarray(bg, np.float32))
    return rgb, alpha, ns, cnt


def test_render_matches_oracle_chain(port, scene, setup):
    from xrnerf_b200.ngp import NgpRenderer
    try:
        from oracle.oracle import Ref, have_ref
        ref = Ref(serial=True) if have_ref() else None
    except Exception:
        ref = None
    f, table, dens, color = setup
    o, d = scene['rays_o'], scene['rays_d']
    bg = (0.1, 0.5, 0.9)
    r = NgpRenderer(f, bg=bg)
    for call in range(2):  # second call exercises the advanced host RNG (n_prior_calls=1)
        rgb, alpha, ns, cnt = r.render(dev(o), dev(d), dev(scene['bitfield']))
        rgb_ref, alpha_ref, ns_ref, cnt_ref = oracle_render(port, ref, scene, table, dens, color, o, d, bg, n_prior=call)
        assert np.array_equal(ns.cpu().numpy(), ns_ref) and np.array_equal(cnt.cpu().numpy(), cnt_ref)
        assert np.abs(rgb.cpu().numpy() - rgb_ref).max() <= 2e-3
        assert np.abs(alpha.cpu().numpy() - alpha_ref).max() <= 2e-3
    assert alpha_ref.max() > 0.2  # the scene is actually visible with these weights


def test_render_full_image_properties(scene, setup):
    """BASELINE-size (800x800 = 640 000 rays): rays that miss get exactly the background and alpha 0; alpha in [0,1];
    rendering the image in 10 row-chunks with the same RNG call index gives a different jitter but the same miss set."""
    from xrnerf_b200 import synth
    from xrnerf_b200.ngp import NgpRenderer
    f = setup[0]
    o, d = synth.get_rays_ngp(scene['poses'][7])
    bg = (0.25, 0.5, 0.75)
    r = NgpRenderer(f, bg=bg, samples_per_ray_budget=48)
    rgb, alpha, ns, cnt = r.render(dev(o), dev(d), dev(scene['bitfield']))
    rgb, alpha, ns = rgb.cpu().numpy(), alpha.cpu().numpy().reshape(-1), ns.cpu().numpy()
    miss = ns[:, 0] == 0
    assert miss.any() and (~miss).any()
    assert np.array_equal(rgb[miss], np.tile(np.array(bg, np.float32), (miss.sum(), 1))) and (alpha[miss] == 0).all()
    assert (alpha >= 0).all() and (alpha <= 1 + 1e-6).all() and np.isfinite(rgb).all()
    assert cnt.cpu().numpy()[1] == ns[:, 0].sum()


def test_render_full_image_vs_oracle(port, scene, setup):
    """BASELINE-size parity (round-1 W2): one whole 800x800 view rendered by both CUDA paths (5-launch chain, single-launch kernel) from rays that reach the
    wrappers exactly as bench.py passes them; sample counts of ALL 640 000 rays bit-exact against the oracle's march, rgb/alpha of every 16th ray within 2e-3
    of the oracle chain (the oracle's tcnn restatement over 7 M samples would take minutes; its march over the full image takes seconds)."""
    from xrnerf_b200 import synth
    from xrnerf_b200.ngp import NgpRenderer
    f, table, dens, color = setup
    o, d = synth.get_rays_ngp(scene['poses'][21])
    bg = (0.3, 0.6, 0.1)
    n = o.shape[0]
    c, _, ns_ref, cnt_ref = port.rays_sampler(o, d, scene['bitfield'], n * 48)
    sel = np.arange(0, n, 16)
    rows = np.concatenate([np.arange(ns_ref[i, 1], ns_ref[i, 1] + ns_ref[i, 0]) for i in sel]) if len(sel) else np.zeros(0, np.int64)
    coords = np.ascontiguousarray(c[rows])
    raw = port.ngp_mlp_forward(table, dens, color, np.ascontiguousarray(coords[:, :3]), np.ascontiguousarray(coords[:, 4:]))
    ns_sel = np.stack([ns_ref[sel, 0], np.concatenate([[0], np.cumsum(ns_ref[sel, 0])[:-1]])], 1).astype(np.int32)
    rgb_ref, alpha_ref = port.calc_rgb_inference(raw, coords, ns_sel, np.asarray(bg, np.float32))
    r = NgpRenderer(f, bg=bg, samples_per_ray_budget=48)
    ot, dt_ = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    rgb, alpha, ns, cnt = r.render(ot, dt_, dev(scene['bitfield']))
    assert not bool(r.overflowed(cnt))
    assert np.array_equal(ns.cpu().numpy(), ns_ref) and int(cnt[1]) == int(cnt_ref[1])
    assert np.abs(rgb.cpu().numpy()[sel] - rgb_ref).max() <= 2e-3 and np.abs(alpha.cpu().numpy()[sel] - alpha_ref).max() <= 2e-3
    r.calls = 0
    rgb2, alpha2, ns2 = r.render_fused(ot, dt_, dev(scene['bitfield']))
    assert np.array_equal(ns2.cpu().numpy(), ns_ref[:, 0])
    assert np.abs(rgb2.cpu().numpy()[sel] - rgb_ref).max() <= 2e-3 and np.abs(alpha2.cpu().numpy()[sel] - alpha_ref).max() <= 2e-3
    assert alpha_ref.max() > 0.2


def test_strided_rays_render_the_same_image(scene, setup):
    """The renderers copy a strided ray tensor instead of reinterpreting its memory (round-1 W1: Fortran-ordered rays_o)."""
    from xrnerf_b200.ngp import NgpRenderer
    f = setup[0]
    o, d = dev(scene['rays_o']), dev(scene['rays_d'])
    o_f = o.t().contiguous().t()                       # same values, strides (1, n)
    assert not o_f.is_contiguous() and torch.equal(o_f, o)
    r = NgpRenderer(f)
    a = [x.clone() for x in r.render(o, d, dev(scene['bitfield']))]
    r.calls = 0
    b = r.render(o_f, d, dev(scene['bitfield']))
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    r.calls = 0
    c1 = [x.clone() for x in r.render_fused(o, d, dev(scene['bitfield']))]
    r.calls = 0
    c2 = r.render_fused(o_f, d, dev(scene['bitfield']))
    assert all(torch.equal(x, y) for x, y in zip(c1, c2))


def test_fused_single_launch_matches_oracle_and_unfused(port, scene, setup):
    """xrb_ngp_render_fused (one launch: march + encode + tcgen05 MLPs + composite) vs the oracle chain and vs the 5-launch path.
    samples per ray: bit-exact. rgb/alpha: 2e-3 vs the oracle (fp16 field), 2e-5 vs the unfused CUDA path (identical field arithmetic, the
    composite is a segmented scan instead of a sequential product)."""
    from xrnerf_b200.ngp import NgpRenderer
    try:
        from oracle.oracle import Ref, have_ref
        ref = Ref(serial=True) if have_ref() else None
    except Exception:
        ref = None
    f, table, dens, color = setup
    o, d = scene['rays_o'], scene['rays_d']
    bg = (0.1, 0.5, 0.9)
    r, r2 = NgpRenderer(f, bg=bg), NgpRenderer(f, bg=bg)
    for call in range(3):  # calls 1, 2 exercise the advanced host RNG and the kernel's self-resetting scheduler words
        rgb, alpha, ns = r.render_fused(dev(o), dev(d), dev(scene['bitfield']))
        rgb_u, alpha_u, ns_u, _ = r2.render(dev(o), dev(d), dev(scene['bitfield']))
        rgb_ref, alpha_ref, ns_ref, cnt_ref = oracle_render(port, ref, scene, table, dens, color, o, d, bg, n_prior=call)
        assert np.array_equal(ns.cpu().numpy(), ns_ref[:, 0])
        assert np.abs(rgb.cpu().numpy() - rgb_ref).max() <= 2e-3 and np.abs(alpha.cpu().numpy() - alpha_ref).max() <= 2e-3
        assert (rgb - rgb_u).abs().max().item() <= 2e-5 and (alpha - alpha_u).abs().max().item() <= 2e-5
    assert int(r._ws_fused[:8].view(torch.int32).abs().sum()) == 0   # scheduler words back to zero


@pytest.mark.parametrize('n', [1, 31, 33, 1000])
def test_fused_ragged_sizes(scene, setup, n):
    from xrnerf_b200.ngp import NgpRenderer
    f = setup[0]
    o, d, bf = dev(scene['rays_o'][:n]), dev(scene['rays_d'][:n]), dev(scene['bitfield'])
    a, b = NgpRenderer(f, bg=(1.0, 1.0, 1.0)), NgpRenderer(f, bg=(1.0, 1.0, 1.0))
    rgb, alpha, ns = a.render_fused(o, d, bf)
    rgb_u, alpha_u, ns_u, _ = b.render(o, d, bf)
    assert torch.equal(ns, ns_u[:, 0]) and (rgb - rgb_u).abs().max().item() <= 2e-5 and (alpha - alpha_u).abs().max().item() <= 2e-5


def test_fused_full_image_and_dense_grid(scene, setup):
    """BASELINE-size image (640 000 rays) through the single-launch kernel == the 5-launch path; and an ALL-ONES occupancy grid (the first
    256 training iterations' worst case: every ray takes hundreds of samples, several 64-sample march rounds per ray)."""
    from xrnerf_b200 import synth
    from xrnerf_b200.ngp import NgpRenderer
    f = setup[0]
    o, d = synth.get_rays_ngp(scene['poses'][7])
    a, b = NgpRenderer(f, bg=(0.25, 0.5, 0.75), samples_per_ray_budget=48), NgpRenderer(f, bg=(0.25, 0.5, 0.75), samples_per_ray_budget=48)
    rgb, alpha, ns = a.render_fused(dev(o), dev(d), dev(scene['bitfield']))
    rgb_u, alpha_u, ns_u, _ = b.render(dev(o), dev(d), dev(scene['bitfield']))
    assert torch.equal(ns, ns_u[:, 0]) and (rgb - rgb_u).abs().max().item() <= 2e-5 and (alpha - alpha_u).abs().max().item() <= 2e-5
    miss = ns == 0
    assert miss.any() and (alpha[miss] == 0).all()
    n = 2048
    ones = torch.full_like(dev(scene['bitfield']), 255)
    a2, b2 = NgpRenderer(f, bg=(0.0, 0.0, 0.0)), NgpRenderer(f, bg=(0.0, 0.0, 0.0), samples_per_ray_budget=1024)
    rgb, alpha, ns = a2.render_fused(dev(o[:n * 300:300]), dev(d[:n * 300:300]), ones)
    rgb_u, alpha_u, ns_u, _ = b2.render(dev(o[:n * 300:300]), dev(d[:n * 300:300]), ones)
    assert int(ns.max()) > 128 and torch.equal(ns, ns_u[:, 0])
    assert (rgb - rgb_u).abs().max().item() <= 1e-4 and (alpha - alpha_u).abs().max().item() <= 1e-4


@pytest.mark.parametrize('T,bwd_impl', [(1 << 16, 1), (1 << 16, 0), (20000, 1)])
def test_trainer_step_matches_autograd_path(scene, setup, T, bwd_impl):
    """One fused training step (xrnerf_b200.train.NgpTrainer: march + compaction on the aux stream, device-side sample counts, tcgen05 field backward) against
    the registry/autograd path the reference's call sites use (rays_sampler_api -> compacted_coord_api -> NgpField autograd -> _CalcRgbBp -> HuberLoss x5):
    same loss, same parameter gradients (T = 2^16: nothing truncated; T = 20000: ray-order truncation, padding rows beyond the compacted count unread), and the
    trained-ray count the trainer reports is what the compaction kept."""
    from xrnerf_b200.ngp import NgpField
    from xrnerf_b200.train import NgpTrainer, huber5_grad
    from xrnerf_b200.registry.renders import _CalcRgbBp
    import xrnerf_b200.raymarch_cuda as rm
    torch.manual_seed(0)
    f = NgpField(seed=5).cuda()
    with torch.no_grad():
        f.hash_params.mul_(3000.0)                          # table values ~0.3 so that the encoding matters
    n = 4096
    o, d, bf = dev(scene['rays_o']), dev(scene['rays_d']), dev(scene['bitfield'])
    target = torch.rand((n, 3), device='cuda'); bg = torch.rand((n, 3), device='cuda')
    # ---- autograd path
    rm.reset_rng()
    cap = n * 64
    coords = torch.zeros((cap, 7), device='cuda'); ridx = torch.zeros((n, 1), dtype=torch.int32, device='cuda'); ns = torch.zeros((n, 2), dtype=torch.int32, device='cuda')
    cnt = torch.zeros(2, dtype=torch.int32, device='cuda')
    rm.rays_sampler_api(o, d, bf, None, None, None, 0.0, 1.0, 0.05, 1.0 / 256, coords, ridx, ns, cnt)
    cc = torch.zeros((T, 7), device='cuda'); nsc = torch.zeros((n, 2), dtype=torch.int32, device='cuda'); rc = torch.zeros(1, dtype=torch.int32, device='cuda'); sc = torch.zeros(1, dtype=torch.int32, device='cuda')
    rm.compacted_coord_api(None, coords, ns, None, 2, 3, 0.0, 1.0, cc, nsc, rc, sc)
    n_c = min(int(sc.item()), T)
    for p_ in (f.hash_params, f.density_params, f.color_params):
        p_.grad = None
    f.impl = 1
    raw = f(cc[:n_c, :3], cc[:n_c, 4:])
    rgb = _CalcRgbBp.apply(raw, cc, ns, nsc, bg, torch.ones(1, device='cuda'), 2, 3, (0.0, 1.0))
    loss_ref, _ = huber5_grad(rgb, target)
    loss_ref.backward()
    g_ref = [p_.grad.clone() for p_ in (f.hash_params, f.density_params, f.color_params)]
    # ---- fused step (weights untouched so far)
    tr = NgpTrainer(f, bf, n, target_batch_size=T, bwd_impl=bwd_impl, lr=0.0)
    loss = float(tr.step(o, d, target, bg))
    torch.cuda.synchronize()
    assert abs(loss - float(loss_ref)) <= 1e-4 * abs(float(loss_ref))
    assert int(tr.compacted_samples()) == int(sc.item()) and int(tr.trained_rays()) == int((nsc[:, 0] == ns[:, 0]).sum())
    if T < int(cnt[1]):
        assert int(tr.trained_rays()) < n                   # truncated rays are reported, not counted as trained
    for name, a, b in zip(('table', 'density', 'color'), tr.grads, g_ref):
        scale = float(b.abs().max())
        err = (a - b).abs()
        assert scale > 0
        # both sides run the same forward kernel; the backward differs only in the path compared (tcgen05 vs itself through autograd: identical; CUDA cores: fp16 dZ staging)
        assert float(err.max()) <= (2e-2 if bwd_impl == 0 else 1e-5) * scale, (name, float(err.max()), scale)
        assert float(torch.sqrt((err.double() ** 2).sum() / (b.double() ** 2).sum())) <= 3e-3, name


def test_trainer_reduces_loss_and_keeps_shadows_current(scene, setup):
    """several fused steps with next-batch prefetch: the loss goes down and the fp16 table / cell image / UMMA weight image the optimiser refreshed are exactly
    what a forced refresh() would build"""
    from xrnerf_b200.ngp import NgpField
    from xrnerf_b200.train import NgpTrainer
    torch.manual_seed(0)
    f = NgpField(seed=5).cuda()
    n = 4096
    o, d, bf = dev(scene['rays_o']), dev(scene['rays_d']), dev(scene['bitfield'])
    target = torch.full((n, 3), 0.5, device='cuda'); bg = torch.zeros((n, 3), device='cuda')
    tr = NgpTrainer(f, bf, n, target_batch_size=1 << 16, ema_momentum=0.05)
    p0 = f.density_params.detach().clone()
    losses = [float(tr.step(o, d, target, bg, next_rays=(o, d))) for _ in range(8)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
    assert not torch.equal(p0, f.density_params.detach())
    pts = torch.rand((1000, 3), device='cuda'); dirs = torch.rand((1000, 3), device='cuda')
    a = f.run_mlp(pts, dirs).clone()
    f.mark_dirty(); f.refresh()
    assert torch.equal(a, f.run_mlp(pts, dirs))
