"""BaseEmbedder / MipNerfEmbedder with the reference's constructor kwargs and data-dict protocol
(/root/reference/xrnerf/models/embedders/base.py:8-77, mipnerf_embedder.py:12-99), computed by one CUDA kernel each."""
import torch
from torch import nn

from .. import _C
from .builder import EMBEDDERS


@EMBEDDERS.register_module()
class BaseEmbedder(nn.Module):
    def __init__(self, i_embed=0, multires=10, multires_dirs=4, input_ch=3, **kwargs):
        super().__init__()
        if input_ch != 3:
            raise NotImplementedError('BaseEmbedder: input_ch=3 (the reference configs)')
        self.i_embed = i_embed
        self.multires = 0 if i_embed == -1 else multires
        self.multires_dirs = 0 if i_embed == -1 else multires_dirs
        self.embed_ch = 3 + 6 * self.multires
        self.embed_ch_dirs = 3 + 6 * self.multires_dirs

    def get_embed_ch(self):
        return self.embed_ch, self.embed_ch_dirs

    def forward(self, data):
        data['unflatten_shape'] = data['pts'].shape[:-1]
        pts, viewdirs = data['pts'], data['viewdirs']
        _C.require_cuda(pts, viewdirs)
        flat = pts.reshape(-1, 3).contiguous().float()
        samples_per_ray = 1 if pts.dim() == viewdirs.dim() else pts.shape[-2]
        vd = viewdirs.reshape(-1, 3).contiguous().float()
        out = torch.empty((flat.shape[0], self.embed_ch + self.embed_ch_dirs), dtype=torch.float32, device=flat.device)
        _C.check(_C.lib.xrb_nerf_posenc(_C.ptr(flat), _C.ptr(vd), flat.shape[0], samples_per_ray, self.multires, self.multires_dirs, _C.ptr(out), _C.stream()), 'posenc')
        data['embedded'] = out
        return data


@EMBEDDERS.register_module()
class MipNerfEmbedder(BaseEmbedder):
    def __init__(self, min_deg_point, max_deg_point, min_deg_view, max_deg_view, input_ch=3, use_viewdirs=False, diag=True, append_identity=True):
        nn.Module.__init__(self)
        if not diag or not append_identity or input_ch != 3:
            raise NotImplementedError('MipNerfEmbedder: diag=True, append_identity=True, input_ch=3 (the reference configs)')
        self.min_deg, self.max_deg = min_deg_point, max_deg_point
        self.min_deg_view, self.max_deg_view = min_deg_view, max_deg_view
        self.use_viewdirs, self.diag, self.append_identity, self.input_ch = use_viewdirs, diag, append_identity, input_ch

    def get_embed_ch(self):
        return 2 * 3 * (self.max_deg - self.min_deg), 2 * 3 * (self.max_deg_view - self.min_deg_view) + 3

    def forward(self, data):
        """Needs data['z_vals'] [N,S+1], rays_o, rays_d, radii, viewdirs (the fused cast_rays+IPE kernel recomputes the Gaussians
        in registers); data['samples'] is accepted for protocol parity but not read."""
        z, o, d, radii, vd = (data[k].contiguous().float() for k in ('z_vals', 'rays_o', 'rays_d', 'radii', 'viewdirs'))
        _C.require_cuda(z, o, d, radii, vd)
        n, s = z.shape[0], z.shape[1] - 1
        c_ipe, c_dir = self.get_embed_ch()
        out = torch.empty((n * s, c_ipe + c_dir), dtype=torch.float32, device=z.device)
        _C.check(_C.lib.xrb_mip_embed(_C.ptr(z), _C.ptr(o), _C.ptr(d), _C.ptr(radii.reshape(-1)), _C.ptr(vd), n, s, self.min_deg, self.max_deg, self.min_deg_view, self.max_deg_view,
                                      _C.ptr(out), None, None, _C.stream()), 'mip_embed')
        data['unflatten_shape'] = torch.Size((n, s))
        data['embedded'] = out
        return data
