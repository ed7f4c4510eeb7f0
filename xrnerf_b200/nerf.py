"""Fused inference path of the hierarchical NeRF (NerfNetwork.forward with N_importance > 0,
/root/reference/xrnerf/models/networks/nerf.py:39-48 + the val pipeline's GetZvals/GetPts, datasets/pipelines/create.py:486-597) for a batch of rays:

  z_vals (linspace) -> [posenc straight into fp16 UMMA tile images, positions formed in registers] -> coarse NerfMLP (tcgen05) -> composite ->
  sample_pdf (one kernel) -> posenc tiles -> fine NerfMLP (tcgen05) -> composite

No [N,S,3] `pts`, no fp32 `embedded`, no per-chunk Python loop: 7 kernel launches per ray batch of any size.
"""
import torch

from . import _C
from .nerf_mlp import nerf_mlp_forward_tiles


class NerfRenderer:
    def __init__(self, network, near=2.0, far=6.0, n_samples=64):
        """network: registry NerfNetwork (mlp, mlp_fine, render); BaseEmbedder encodings."""
        self.net, self.near, self.far, self.S = network, near, far, n_samples
        e = network.mlp.embedder
        self.multires, self.multires_dirs = e.multires, e.multires_dirs

    def _mlp(self, mlp, rays_o, rays_d, viewdirs, z):
        n, s = z.shape
        rows = n * s
        enc = torch.empty(_C.lib.xrb_nerf_enc_image_bytes(rows, mlp.input_ch), dtype=torch.uint8, device=z.device)
        _C.check(_C.lib.xrb_nerf_posenc_tiles_rays(_C.ptr(rays_o), _C.ptr(rays_d), _C.ptr(z), _C.ptr(viewdirs), n, s, self.multires, self.multires_dirs, _C.ptr(enc), _C.stream()),
                 'posenc_tiles_rays')
        image, bias = mlp._packed()
        assert mlp.kernel_version == 2
        return nerf_mlp_forward_tiles(image, bias, enc, rows, mlp.input_ch, mlp.input_ch_dirs).view(n, s, 4)

    @torch.no_grad()
    def render(self, rays_o, rays_d, viewdirs):
        from .registry.networks import sample_pdf
        net = self.net
        n = rays_o.shape[0]
        t = torch.linspace(0., 1., self.S, device=rays_o.device)
        z = (self.near * (1. - t) + self.far * t).expand(n, self.S).contiguous()           # GetZvals (create.py:502-531), lindisp=False
        data = {'rays_o': rays_o, 'rays_d': rays_d, 'viewdirs': viewdirs, 'z_vals': z}
        data['raw'] = self._mlp(net.mlp, rays_o, rays_d, viewdirs, z)
        data, ret = net.render(data, is_test=True)
        if getattr(net, 'N_importance', 0) > 0:
            data['pts'] = None
            data = sample_pdf(data, net.N_importance, False, True)
            data['raw'] = self._mlp(net.mlp_fine, rays_o, rays_d, viewdirs, data['z_vals'])
            _, fine = net.render(data, is_test=True)
            ret = {'coarse_rgb': ret['rgb'], 'rgb': fine['rgb'], 'disp': fine['disp'], 'acc': fine['acc']}
        return ret
