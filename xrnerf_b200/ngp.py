"""Host-side objects of the fused Instant-NGP path: parameter storage (tcnn layouts), fp16 working copies, the packed
tensor-core weight image, and thin wrappers over the C ABI (include/xrnerf_b200.h: xrb_ngp_*).
"""
import math

import torch
from torch import nn

from . import _C

PER_LEVEL_SCALE = float(2.0 ** (math.log2(2048 * 1 / 16) / (16 - 1)))  # hashnerf_mlp.py:17-20 with bound=1 (Q6)


class NgpField(nn.Module):
    """HashNerfMLP's arithmetic (hash grid + SH + density/colour FullyFusedMLPs) as ONE kernel.

    Parameters keep tcnn's flat layouts and the reference's state_dict keys are produced by the registry class
    (xrnerf_b200.api.HashNerfMLP) that owns one of these. `impl`: 1 = tcgen05 tensor-core tiles, 0 = CUDA cores.
    """

    def __init__(self, n_levels=16, n_features=2, log2_hashmap_size=19, base_resolution=16, per_level_scale=PER_LEVEL_SCALE, width=64,
                 density_hidden=1, color_hidden=2, seed=1337, impl=1):
        super().__init__()
        self.cfg = _C.NgpConfig(n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale, width, density_hidden, color_hidden)
        n_hash = _C.lib.xrb_tcnn_hashgrid_num_params(self.cfg)
        if n_hash < 0:
            raise _C.XrbError(_C.lib.xrb_last_error().decode())
        g = torch.Generator().manual_seed(seed)
        self.hash_params = nn.Parameter((torch.rand(n_hash, generator=g) * 2 - 1) * 1e-4)

        def xavier(shapes):
            return torch.cat([((torch.rand(o, i, generator=g) * 2 - 1) * math.sqrt(6.0 / (i + o))).reshape(-1) for (o, i) in shapes])
        self.density_params = nn.Parameter(xavier([(width, 32)] + [(width, width)] * (density_hidden - 1) + [(16, width)]))
        self.color_params = nn.Parameter(xavier([(width, 32)] + [(width, width)] * (color_hidden - 1) + [(16, width)]))
        self.impl = impl
        self._ver = None
        self._table16 = self._dens16 = self._color16 = self._image = None

    # ---- fp16 working copies + packed UMMA weight image, refreshed when a master parameter changed
    def refresh(self, force=False):
        ver = (self.hash_params._version, self.density_params._version, self.color_params._version, self.hash_params.device)
        if not force and ver == self._ver:
            return
        dev = self.hash_params.device
        if self._table16 is None or self._table16.device != dev:
            self._table16 = torch.empty(self.hash_params.numel(), dtype=torch.float16, device=dev)
            self._dens16 = torch.empty(self.density_params.numel(), dtype=torch.float16, device=dev)
            self._color16 = torch.empty(self.color_params.numel(), dtype=torch.float16, device=dev)
            self._image = torch.empty(_C.lib.xrb_ngp_weight_image_bytes(self.cfg), dtype=torch.uint8, device=dev)
        s = _C.stream()
        _C.check(_C.lib.xrb_tcnn_cast_params(_C.ptr(self.hash_params.detach()), _C.ptr(self._table16), self.hash_params.numel(), s), 'cast hash')
        _C.check(_C.lib.xrb_tcnn_cast_params(_C.ptr(self.density_params.detach()), _C.ptr(self._dens16), self.density_params.numel(), s), 'cast density')
        _C.check(_C.lib.xrb_tcnn_cast_params(_C.ptr(self.color_params.detach()), _C.ptr(self._color16), self.color_params.numel(), s), 'cast color')
        _C.check(_C.lib.xrb_ngp_pack_weights(self.cfg, _C.ptr(self.density_params.detach()), _C.ptr(self.color_params.detach()), _C.ptr(self._image), s), 'pack')
        self._ver = ver

    def run_mlp(self, pts, dirs, impl=None):
        """pts, dirs: [S,3] float32 views (row stride in floats may be > 3, e.g. coords[:, :3] / coords[:, 4:]). -> raw [S,4] f32."""
        _C.require_cuda(pts, dirs)
        self.refresh()
        assert pts.dtype == torch.float32 and dirs.dtype == torch.float32 and pts.stride(1) == 1 and dirs.stride(1) == 1
        n = pts.shape[0]
        raw = torch.empty((n, 4), dtype=torch.float32, device=pts.device)
        impl = self.impl if impl is None else impl
        _C.check(_C.lib.xrb_ngp_mlp_forward(self.cfg, _C.ptr(self._table16), _C.ptr(self._dens16), _C.ptr(self._color16), _C.ptr(self._image), _C.ptr(pts),
                                            pts.stride(0), _C.ptr(dirs), dirs.stride(0), n, _C.ptr(raw), impl, _C.stream()), 'ngp_mlp_forward')
        return raw

    def forward(self, pts, dirs):
        """differentiable (w.r.t. the parameters) HashNerfMLP.run_mlp"""
        if torch.is_grad_enabled() and any(p.requires_grad for p in (self.hash_params, self.density_params, self.color_params)):
            return _FieldFn.apply(self, pts, dirs, self.hash_params, self.density_params, self.color_params)
        return self.run_mlp(pts, dirs)

    def backward_params(self, pts, dirs, grad_raw, out=None):
        """dL/draw [S,4] -> (d_hash, d_density, d_color) fp32, accumulated into `out` if given (zero-initialised otherwise)."""
        _C.require_cuda(pts, dirs, grad_raw)
        self.refresh()
        if out is None:
            out = (torch.zeros_like(self.hash_params), torch.zeros_like(self.density_params), torch.zeros_like(self.color_params))
        n = pts.shape[0]
        _C.check(_C.lib.xrb_ngp_mlp_backward(self.cfg, _C.ptr(self._table16), _C.ptr(self._dens16), _C.ptr(self._color16), _C.ptr(pts), pts.stride(0), _C.ptr(dirs),
                                             dirs.stride(0), _C.ptr(grad_raw), n, _C.ptr(out[0]), _C.ptr(out[1]), _C.ptr(out[2]), _C.stream()), 'ngp_mlp_backward')
        return out

    def run_density(self, pts, impl=None):
        _C.require_cuda(pts)
        self.refresh()
        assert pts.dtype == torch.float32 and pts.stride(1) == 1
        n = pts.shape[0]
        out = torch.empty((n, 1), dtype=torch.float32, device=pts.device)
        impl = self.impl if impl is None else impl
        _C.check(_C.lib.xrb_ngp_density_forward(self.cfg, _C.ptr(self._table16), _C.ptr(self._dens16), _C.ptr(self._image), _C.ptr(pts), pts.stride(0), n,
                                                _C.ptr(out), impl, _C.stream()), 'ngp_density_forward')
        return out


class _FieldFn(torch.autograd.Function):
    """autograd bridge for NgpField.run_mlp: forward = xrb_ngp_mlp_forward, backward = xrb_ngp_mlp_backward (one kernel
    producing the gradients of all three parameter vectors). pts/dirs are not differentiated (the reference detaches them,
    hashnerf_mlp.py:58-68)."""

    @staticmethod
    def forward(ctx, field, pts, dirs, hash_params, density_params, color_params):
        ctx.field = field
        ctx.save_for_backward(pts, dirs)
        return field.run_mlp(pts, dirs)

    @staticmethod
    def backward(ctx, grad_raw):
        f = ctx.field
        pts, dirs = ctx.saved_tensors
        grad_raw = grad_raw.contiguous().to(torch.float32)
        dt, dd, dc = f.backward_params(pts, dirs, grad_raw)
        return None, None, None, dt, dd, dc


class NgpRenderer:
    """Fused inference render of a ray batch (xrb_ngp_render): march -> field -> composite, no host sync, persistent workspace."""

    def __init__(self, field, aabb=(0.0, 1.0), near=0.05, cone=1.0 / 256, rgb_act=2, dens_act=3, bg=(0.0, 0.0, 0.0), samples_per_ray_budget=64):
        self.field, self.aabb, self.near, self.cone = field, aabb, near, cone
        self.rgb_act, self.dens_act, self.bg = rgb_act, dens_act, tuple(bg)
        self.budget = samples_per_ray_budget
        self.calls = 0
        self._ws = None
        self._out = {}

    def render(self, rays_o, rays_d, bitfield, out=None):
        _C.require_cuda(rays_o, rays_d, bitfield)
        f = self.field
        f.refresh()
        n = rays_o.shape[0]
        max_samples = n * self.budget
        need = _C.lib.xrb_ngp_render_workspace(n, max_samples)
        if self._ws is None or self._ws.numel() < need or self._ws.device != rays_o.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=rays_o.device)
        if out is None:
            key = (n, rays_o.device)
            if key not in self._out:
                self._out[key] = (torch.empty((n, 3), dtype=torch.float32, device=rays_o.device), torch.empty((n, 1), dtype=torch.float32, device=rays_o.device),
                                  torch.empty((n, 2), dtype=torch.int32, device=rays_o.device), torch.empty(2, dtype=torch.int32, device=rays_o.device))
            out = self._out[key]
        rgb, alpha, numsteps, counters = out
        _C.check(_C.lib.xrb_ngp_render(f.cfg, _C.ptr(f._table16), _C.ptr(f._image), _C.ptr(bitfield), _C.ptr(rays_o), _C.ptr(rays_d), n, max_samples, self.aabb[0],
                                       self.aabb[1], self.near, self.cone, 9121, self.calls, _C.float3(self.bg), self.rgb_act, self.dens_act, _C.ptr(rgb), _C.ptr(alpha),
                                       _C.ptr(numsteps), _C.ptr(counters), _C.ptr(self._ws), _C.stream()), 'ngp_render')
        self.calls += 1
        return rgb, alpha, numsteps, counters

    def render_fused(self, rays_o, rays_d, bitfield, out=None, ws=None):
        """Single-launch render (xrb_ngp_render_fused). Returns (rgb [n,3], alpha [n,1], n_samples i32[n]). `ws`: a zero-initialised workspace
        owned by the caller (one per stream when several batches are in flight); default: one per renderer."""
        _C.require_cuda(rays_o, rays_d, bitfield)
        f = self.field
        f.refresh()
        n = rays_o.shape[0]
        if ws is None:
            if getattr(self, '_ws_fused', None) is None or self._ws_fused.device != rays_o.device:
                self._ws_fused = torch.zeros(_C.lib.xrb_ngp_render_fused_workspace(), dtype=torch.uint8, device=rays_o.device)
            ws = self._ws_fused
        if out is None:
            key = ('fused', n, rays_o.device)
            if key not in self._out:
                self._out[key] = (torch.empty((n, 3), dtype=torch.float32, device=rays_o.device), torch.empty((n, 1), dtype=torch.float32, device=rays_o.device),
                                  torch.empty(n, dtype=torch.int32, device=rays_o.device))
            out = self._out[key]
        rgb, alpha, ns = out
        _C.check(_C.lib.xrb_ngp_render_fused(f.cfg, _C.ptr(f._table16), _C.ptr(f._image), _C.ptr(bitfield), _C.ptr(rays_o), _C.ptr(rays_d), n, self.aabb[0], self.aabb[1],
                                             self.near, self.cone, 9121, self.calls, _C.float3(self.bg), self.rgb_act, self.dens_act, _C.ptr(rgb), _C.ptr(alpha), _C.ptr(ns),
                                             _C.ptr(ws), _C.stream()), 'ngp_render_fused')
        self.calls += 1
        return rgb, alpha, ns
