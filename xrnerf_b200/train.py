"""Fused Instant-NGP training step (HashNerfNetwork.train_step, /root/reference/xrnerf/models/networks/hashnerf.py:32-52 +
ngp_grid_sampler.py:189-266 + the mmcv OptimizerHook/Adam it runs under) without host synchronisation:

  march -> compaction -> field forward (tcgen05) -> composite forward -> Huber x5 loss -> composite backward ->
  field backward (one kernel) -> [all-reduce of ONE flat fp32 gradient buffer] -> fused Adam (+ fp16 / UMMA-image refresh)

Data parallel: every rank holds identical weights and takes a disjoint ray batch; the only collective is the gradient
all-reduce (what MMDistributedDataParallel does implicitly in the reference, core/apis/train.py:28-36); the sum is divided by
world size inside the Adam kernel. The occupancy grid is updated identically on every rank (same weights + same RNG call
index => same grid), so it needs no communication.
"""
import torch

from . import _C
from . import raymarch_cuda as rm


class FlatGradBuffer:
    """One contiguous fp32 buffer viewed as the per-parameter gradients (host logic; works on CPU tensors for the gloo test)."""

    def __init__(self, params):
        self.sizes = [int(p.numel()) for p in params]
        self.flat = torch.zeros(sum(self.sizes), dtype=torch.float32, device=params[0].device)
        self.views, o = [], 0
        for n in self.sizes:
            self.views.append(self.flat[o:o + n])
            o += n

    def zero_(self):
        self.flat.zero_()

    def allreduce(self, group=None):
        """SUM over ranks (the division by world size happens in the optimiser kernel). Returns the divisor to apply."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            return float(dist.get_world_size(group))
        return 1.0


def huber5_grad(rgb, target, delta=0.1):
    """d/d rgb of 5 * HuberLoss(rgb, target, 0.1, 'sum') (networks/utils/metrics.py:8-16, hashnerf.py:39-44) and the loss value."""
    diff = rgb - target
    rel = diff.abs()
    loss = torch.where(rel > delta, rel - 0.5 * delta, 0.5 / delta * rel * rel).sum() * 5
    grad = torch.where(rel > delta, torch.sign(diff), diff / delta) * 5
    return loss, grad


class NgpTrainer:
    def __init__(self, field, bitfield, n_rays, target_batch_size=1 << 18, aabb=(0.0, 1.0), near=0.05, cone=1.0 / 256, rgb_act=2, dens_act=3,
                 lr=1e-2, betas=(0.9, 0.99), eps=1e-15, weight_decay=1e-6, samples_per_ray_budget=64, group=None, ema_momentum=None, ema_warm_up=100):
        self.f, self.bitfield, self.n_rays, self.T = field, bitfield, n_rays, target_batch_size
        self.aabb, self.near, self.cone, self.rgb_act, self.dens_act = aabb, near, cone, rgb_act, dens_act
        self.lr, self.betas, self.eps, self.wd, self.group = lr, betas, eps, weight_decay, group
        dev = field.hash_params.device
        self.params = [field.hash_params, field.density_params, field.color_params]
        self.grads = FlatGradBuffer(self.params)
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.step_n = 0
        # EMAHook(momentum=0.05) of the reference config (nerf_blender_local01.py:24), folded into the Adam pass; buffers start as copies (EMAHook.before_run)
        self.ema_momentum, self.ema_warm_up = ema_momentum, ema_warm_up
        self.ema = [p.detach().clone() for p in self.params] if ema_momentum is not None else [None] * len(self.params)
        cap = n_rays * samples_per_ray_budget
        self.coords = torch.empty((cap, 7), dtype=torch.float32, device=dev)
        self.coords_c = torch.zeros((self.T, 7), dtype=torch.float32, device=dev)
        self.raw = torch.empty((self.T, 4), dtype=torch.float32, device=dev)
        self.draw = torch.zeros((self.T, 4), dtype=torch.float32, device=dev)
        self.rays_index = torch.zeros((n_rays, 1), dtype=torch.int32, device=dev)
        self.numsteps = torch.zeros((n_rays, 2), dtype=torch.int32, device=dev)
        self.numsteps_c = torch.zeros((n_rays, 2), dtype=torch.int32, device=dev)
        self.counter = torch.zeros(2, dtype=torch.int32, device=dev)
        self.cnt_c = torch.zeros(2, dtype=torch.int32, device=dev)
        self.rgb = torch.zeros((n_rays, 3), dtype=torch.float32, device=dev)
        self.grid_mean = torch.ones(1, dtype=torch.float32, device=dev)
        field.refresh()

    def step(self, rays_o, rays_d, target, bg):
        f = self.f
        self.counter.zero_(); self.cnt_c.zero_()
        rm.rays_sampler_api(rays_o, rays_d, self.bitfield, None, None, None, self.aabb[0], self.aabb[1], self.near, self.cone, self.coords, self.rays_index, self.numsteps, self.counter)
        # the reference's no-grad pre-pass over ALL samples (ngp_grid_sampler.py:229-230) feeds only compacted_coord's dead transmittance loop
        # (compacted_coord.cu:41-44): skipping it changes no result
        self.coords_c.zero_()
        rm.compacted_coord_api(None, self.coords, self.numsteps, None, self.rgb_act, self.dens_act, self.aabb[0], self.aabb[1], self.coords_c, self.numsteps_c, self.cnt_c[0:1], self.cnt_c[1:2])
        n_rows = self.T
        pp, dp = _C.rows(self.coords_c[:, :3])[0], _C.rows(self.coords_c[:, 4:])[0]
        _C.check(_C.lib.xrb_ngp_mlp_forward(f.cfg, f.tab, _C.ptr(f._dens16), _C.ptr(f._color16), _C.ptr(f._image), pp, 7, dp, 7, n_rows, _C.ptr(self.raw), 1, _C.stream()), 'field fwd')
        rm.calc_rgb_forward_api(self.raw, self.coords_c, self.numsteps, self.numsteps_c, bg, self.rgb_act, self.dens_act, 0.0, 1.0, self.rgb)
        loss, g = huber5_grad(self.rgb, target)
        self.draw.zero_()
        rm.calc_rgb_backward_api(self.raw, self.numsteps_c, self.coords_c, g, self.rgb, self.grid_mean, self.rgb_act, self.dens_act, 0.0, 1.0, self.draw)
        self.grads.zero_()
        gv = self.grads.views
        _C.check(_C.lib.xrb_ngp_mlp_backward(f.cfg, f.tab, _C.ptr(f._dens16), _C.ptr(f._color16), pp, 7, dp, 7,
                                             _C.ptr(self.draw), n_rows, _C.ptr(gv[0]), _C.ptr(gv[1]), _C.ptr(gv[2]), _C.stream()), 'field bwd')
        div = self.grads.allreduce(self.group)
        self.step_n += 1
        shadows = [f._table16, f._dens16, f._color16]
        it = self.step_n - 1                                                            # runner.iter inside after_train_iter
        mom = 0.0 if self.ema_momentum is None else min(self.ema_momentum, (1 + it) / (self.ema_warm_up + it))
        for p, p16, g_, m, v, e in zip(self.params, shadows, gv, self.m, self.v, self.ema):
            _C.check(_C.lib.xrb_adam_ema_step(_C.ptr(p.data), _C.ptr(p16), _C.ptr(g_), _C.ptr(m), _C.ptr(v), p.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
                                              self.step_n, div, _C.ptr(e), mom, _C.stream()), 'adam')
        _C.check(_C.lib.xrb_ngp_pack_weights(f.cfg, _C.ptr(f.density_params.data), _C.ptr(f.color_params.data), _C.ptr(f._image), _C.stream()), 'pack')
        f.rebuild_cells()                                                               # the fp16 table changed: refresh its cell image
        f._ver = (f.hash_params._version, f.density_params._version, f.color_params._version, f.hash_params.device)  # shadows are current
        return loss
