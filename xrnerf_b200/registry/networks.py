"""NerfNetwork / MipNerfNetwork / HashNerfNetwork: the control flow of the reference's networks
(/root/reference/xrnerf/models/networks/nerf.py:15-180, mipnerf.py:14-74, hashnerf.py:16-52) over this package's kernels.
The mmcv-runner-facing API is kept: train_step(data, optimizer) -> {loss, log_vars, num_samples}, val_step, batchify_forward,
set_val_pipeline. Dataset/runner/hook plumbing is out of scope."""
import torch
from torch import nn

from .. import _C
from . import builder
from .builder import NETWORKS

# ---- networks/utils: metrics.py:3-16, batching.py, transforms.py (merge_ret / recover_shape)
img2mse = lambda x, y: torch.mean((x - y) ** 2)
mse2psnr = lambda x: -10. * torch.log(x) / torch.log(torch.tensor([10.], device=x.device))


def HuberLoss(x, y, delta=0.1, reduction='sum'):
    rel = (x - y).abs()
    loss = torch.where(rel > delta, rel - 0.5 * delta, 0.5 / delta * rel * rel)
    return loss.mean() if reduction == 'mean' else loss.sum()


def unfold_batching(x):
    """networks/utils/batching.py:5-12: (bs, N_per_sampler, ...) -> (bs * N_per_sampler, ...); 1-D tensors and non-tensors pass through."""
    if torch.is_tensor(x) and x.dim() > 1:
        return x.reshape((-1,) + tuple(x.shape[2:]))     # == torch.cat([x[b] for b in range(bs)], 0)
    return x


def merge_ret(ret, fine_ret):
    out = {'coarse_' + k: v for k, v in ret.items()}
    out.update(fine_ret)
    return out


def recover_shape(x, src_shape):
    return x.reshape([int(v) for v in src_shape[:2]] + [-1])


def sample_pdf(data, N_samples, is_perturb=False, is_test=False):
    """hierarchical_sample.py:6-53 as one kernel (inverse CDF + merge sort + pts)."""
    z, w, o, d = (data[k].contiguous().float() for k in ('z_vals', 'weights', 'rays_o', 'rays_d'))
    _C.require_cuda(z, w, o, d)
    n, s = z.shape
    det = is_test or not is_perturb
    u = None if det else torch.rand((n, N_samples), device=z.device)
    z_out = torch.empty((n, s + N_samples), dtype=torch.float32, device=z.device)
    want_pts = not ('pts' in data and data['pts'] is None)   # the fused renderer forms positions inside the encoding kernel and asks for no pts
    pts = torch.empty((n, s + N_samples, 3), dtype=torch.float32, device=z.device) if want_pts else None
    _C.check(_C.lib.xrb_nerf_sample_pdf(_C.ptr(z), _C.ptr(w.detach()), _C.ptr(o), _C.ptr(d), _C.ptr(u), n, s, N_samples, _C.ptr(z_out), _C.ptr(pts), _C.stream()), 'sample_pdf')
    data['pts'], data['z_vals'] = pts, z_out
    return data


def resample_along_rays(data, randomized, ray_shape, resample_padding, rand=None):
    """mip.py:146-176 as one kernel (blur + piecewise-constant inverse CDF). `rand`: the [n, S+1] uniforms the reference draws with torch.rand when
    `randomized` (mip.py:31-33); default: drawn on the device."""
    if ray_shape != 'cone':
        raise NotImplementedError('ray_shape cone (the reference configs)')
    z, w = data['z_vals'].contiguous().float(), data['weights'].detach().contiguous().float()
    n, s1 = z.shape
    u = None
    if randomized:
        sdt = 1.0 / s1
        r = torch.rand((n, s1), device=z.device) if rand is None else rand.to(z.device).float()
        u = torch.arange(s1, device=z.device) * sdt + r * (sdt - torch.finfo(torch.float32).eps)
        u = torch.minimum(u, torch.tensor(1. - torch.finfo(torch.float32).eps, device=z.device)).contiguous()
    z_new = torch.empty_like(z)
    _C.check(_C.lib.xrb_mip_resample(_C.ptr(z), _C.ptr(w), _C.ptr(u), n, s1 - 1, float(resample_padding), _C.ptr(z_new), _C.stream()), 'mip_resample')
    data['z_vals'] = z_new
    return data


@NETWORKS.register_module()
class NerfNetwork(nn.Module):
    def __init__(self, cfg, mlp=None, mlp_fine=None, render=None):
        super().__init__()
        self.phase = cfg.get('phase', 'train')
        for k in ('chunk', 'bs_data', 'is_perturb', 'N_importance'):
            if k in cfg:
                setattr(self, k, cfg[k])
        if mlp is not None:
            self.mlp = builder.build_mlp(mlp)
        if mlp_fine is not None:
            self.mlp_fine = builder.build_mlp(mlp_fine)
        if render is not None:
            self.render = builder.build_render(render)
        self.val_pipeline = None

    def forward(self, data, is_test=False):
        data, ret = self.render(self.mlp(data), is_test)
        if self.N_importance > 0:
            data = sample_pdf(data, self.N_importance, self.is_perturb, is_test)
            _, fine_ret = self.render(self.mlp_fine(data), is_test)
            ret = merge_ret(ret, fine_ret)
        return ret

    def batchify_forward(self, data, is_test=False):
        N = data[self.bs_data].shape[0]
        all_ret = {}
        for i in range(0, N, self.chunk):
            chunk = {k: (v[i:i + self.chunk] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == N else v) for k, v in data.items()}
            ret = self.forward(chunk, is_test)
            for k, v in ret.items():
                all_ret.setdefault(k, []).append(v)
        return {k: torch.cat(v, 0) for k, v in all_ret.items()}

    def train_step(self, data, optimizer, **kwargs):
        data = {k: unfold_batching(v) for k, v in data.items()}
        ret = self.forward(data, is_test=False)
        img_loss = img2mse(ret['rgb'], data['target_s'])
        psnr = mse2psnr(img_loss)
        loss = img_loss
        if 'coarse_rgb' in ret:
            loss = loss + img2mse(ret['coarse_rgb'], data['target_s'])
        return {'loss': loss, 'log_vars': {'loss': loss.item(), 'psnr': psnr.item()}, 'num_samples': ret['rgb'].shape[0]}

    def val_step(self, data, optimizer=None, **kwargs):
        """Renders data['poses'] through the installed val pipeline (rank 0 only in the reference; here every rank renders what it is given)."""
        data = {k: unfold_batching(v) for k, v in data.items()}
        for m in (getattr(self, 'mlp', None), getattr(self, 'mlp_fine', None)):
            if m is not None and hasattr(m, 'mark_dirty'):
                m.mark_dirty()                      # weights may have been swapped through `.data` (EMAHook) since the last pack
        rgbs = []
        with torch.no_grad():
            for i in range(data['poses'].shape[0]):
                d = self.val_pipeline({'pose': data['poses'][i], 'idx': i})
                ret = self.batchify_forward(d, is_test=True)
                rgbs.append(recover_shape(ret['rgb'], d['src_shape']))
        return {'rgbs': rgbs}

    def set_val_pipeline(self, func):
        self.val_pipeline = func


@NETWORKS.register_module()
class MipNerfNetwork(NerfNetwork):
    def __init__(self, cfg, mlp=None, render=None):
        super().__init__(cfg, mlp=mlp, render=render)
        self.num_levels, self.resample_padding, self.ray_shape = cfg.num_levels if hasattr(cfg, 'num_levels') else cfg['num_levels'], cfg['resample_padding'], cfg['ray_shape']
        self.use_multiscale, self.coarse_loss_mult = cfg['use_multiscale'], cfg['coarse_loss_mult']

    def forward(self, data, is_test):
        randomized = not is_test
        ret = {}
        for i_level in range(self.num_levels):
            if i_level > 0:
                data = resample_along_rays(data, randomized, self.ray_shape, self.resample_padding)
            # level 0: sample_along_rays == cast_rays on the given z_vals; the Gaussians are formed inside the embedder kernel
            data, temp_ret = self.render(self.mlp(data), is_test)
            ret = temp_ret if not ret else merge_ret(ret, temp_ret)
        return ret

    def train_step(self, data, optimizer, **kwargs):
        data = {k: unfold_batching(v) for k, v in data.items()}
        ret = self.forward(data, is_test=False)
        mask = torch.broadcast_to(data['lossmult'], ret['rgb'].shape) if 'lossmult' in data else torch.ones_like(ret['rgb'])
        loss_fine = (mask * (ret['rgb'] - data['target_s']) ** 2).sum() / mask.sum()
        loss_coarse = (mask * (ret['coarse_rgb'] - data['target_s']) ** 2).sum() / mask.sum()
        loss = loss_fine + self.coarse_loss_mult * loss_coarse
        psnr = mse2psnr(loss_fine)
        return {'loss': loss, 'log_vars': {'loss': loss.item(), 'loss_fine': loss_fine.item(), 'loss_coarse': loss_coarse.item(), 'psnr': psnr.item()}, 'num_samples': ret['rgb'].shape[0]}


@NETWORKS.register_module()
class HashNerfNetwork(NerfNetwork):
    def __init__(self, cfg, sampler=None, mlp=None, render=None):
        super().__init__(cfg)
        self.sampler = builder.build_sampler(sampler)
        self.mlp = builder.build_mlp(mlp)
        self.render = builder.build_render(render)

    def forward(self, data, is_test=False):
        data = self.sampler.sample(data, self.mlp, is_test)
        data = self.mlp(data)
        data, ret = self.render(data, self.sampler, is_test)
        return ret

    def batchify_forward(self, data, is_test=False):
        """Inference over a whole ray set (hashnerf.py:69-71 via nerf.py:50-69: 4096-ray chunks, each = sample (host sync) -> 3 tcnn launches -> composite).
        Here the test path is ONE fused launch per call for any number of rays (xrb_ngp_render_fused; no chunk loop, no host sync, no [S,7]/[S,4] buffers);
        `fused_inference = False` or training falls back to the reference-shaped chunk loop over forward()."""
        smp = self.sampler
        if not (is_test and getattr(self, 'fused_inference', True) and data['rays_o'].is_cuda and hasattr(smp, 'aabb_range')):
            return super().batchify_forward(data, is_test)
        from ..ngp import NgpRenderer
        smp.check_device(data)
        r = getattr(self, '_renderer', None)
        if r is None:
            r = self._renderer = NgpRenderer(self.mlp.field, aabb=tuple(float(a) for a in smp.aabb_range), near=smp.near_distance, cone=smp.cone_angle_constant,
                                             rgb_act=int(smp.rgb_activation), dens_act=int(smp.density_activation), bg=tuple(float(c) for c in self.render.bg_color))
        import xrnerf_b200.raymarch_cuda as rm
        r.calls = rm.rng_calls['ray_sampler']               # the same jitter stream the unfused path would use next (hidden static pcg32 of the reference, Q9)
        rgb, alpha, _ = r.render_fused(data['rays_o'].contiguous().float(), data['rays_d'].contiguous().float(), smp.density_grid_bitfield)
        rm.rng_calls['ray_sampler'] += 1
        return {'rgb': rgb.clone(), 'alpha': alpha.clone()}

    def val_step(self, data, optimizer=None, **kwargs):
        """hashnerf.py:54-93 (rank-0 loop over validation poses) with the image rendered by one launch and the PSNR taken on the device; every rank renders the
        poses it is given. Returns the reference's keys plus 'psnr'."""
        if self.phase == 'test':
            return self.test_step(data, **kwargs)
        # validation may run right after mmcv's EMAHook swapped the parameters through `.data` (Parameter._version unchanged): rebuild the fp16 shadows, the cell
        # image and the UMMA weight image once per call instead of trusting the version key
        self.mlp.mark_dirty()
        data = {k: unfold_batching(v) for k, v in data.items()}
        poses, images = data['poses'], data.get('images')
        rgbs, gt_imgs, psnrs, elapsed = [], [], [], []
        with torch.no_grad():
            for i in range(poses.shape[0]):
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                d = self.val_pipeline({'pose': poses[i], 'idx': i})
                ret = self.batchify_forward(d, is_test=True)
                rgb = recover_shape(ret['rgb'], d['src_shape'])
                if images is not None:
                    alpha = images[i][..., 3:].to(rgb.device)
                    gt = images[i][..., :3].to(rgb.device) * alpha
                    rgb = rgb * alpha
                    gt_imgs.append(gt)
                    psnrs.append(mse2psnr(img2mse(rgb, gt)))
                t1.record()
                rgbs.append(rgb); elapsed.append((t0, t1))
        torch.cuda.synchronize()
        return {'rgbs': rgbs, 'disps': [], 'gt_imgs': gt_imgs, 'elapsed_time': [a.elapsed_time(b) * 1e-3 for a, b in elapsed], 'psnr': [float(p.item()) for p in psnrs]}

    def test_step(self, data, **kwargs):
        """hashnerf.py:95-111: one spiral pose per call; `data` already holds the pose's rays (rays_o, rays_d, src_shape, idx) from the test pipeline.
        A dict with only `poses` is run through the installed val pipeline first."""
        data = {k: unfold_batching(v) for k, v in data.items()}
        idx = data.get('idx', 0)
        idx = int(idx.item()) if torch.is_tensor(idx) else int(idx)
        if idx == 0 or not getattr(self, '_test_refreshed', False):       # first pose of a test run: the weights may have been swapped / loaded through `.data`
            self.mlp.mark_dirty()
            self._test_refreshed = True
        with torch.no_grad():
            d = data if 'rays_o' in data else self.val_pipeline({'pose': data['poses'], 'idx': idx})
            ret = self.batchify_forward(d, is_test=True)
        rgb, alpha = ret['rgb'], ret['alpha']
        if 'src_shape' in d:
            rgb, alpha = recover_shape(rgb, d['src_shape']), recover_shape(alpha, d['src_shape'])
        return {'spiral_rgb': rgb, 'spiral_alpha': alpha, 'idx': idx}

    def train_step(self, data, optimizer, **kwargs):
        data = {k: unfold_batching(v) for k, v in data.items()}
        ret = self.forward(data, is_test=False)
        bs = ret['rgb'].shape[0]
        alpha = data['alpha'].detach()
        huber = HuberLoss(ret['rgb'], data['target_s'], 0.1, 'sum')
        psnr = mse2psnr(img2mse(ret['rgb'] * alpha, data['target_s'] * alpha))
        loss = huber * 5
        return {'loss': loss, 'log_vars': {'loss': loss.item(), 'psnr': psnr.item()}, 'num_samples': bs}
