"""NerfMLP / HashNerfMLP with the reference's kwargs, state_dict keys and data-dict protocol
(/root/reference/xrnerf/models/mlps/nerf_mlp.py:11-94, hashnerf_mlp.py:23-111)."""
import torch
import torch.nn.functional as F
from torch import nn

from . import builder
from .builder import MLPS
from ..ngp import NgpField, PER_LEVEL_SCALE


@MLPS.register_module()
class BaseMLP(nn.Module):
    def __init__(self, **kwarg):
        super().__init__()


@MLPS.register_module()
class NerfMLP(BaseMLP):
    """8x256 ReLU MLP with skip at layer 4 and a 128-wide view branch; parameters are nn.Linear modules with the reference's
    names (pts_linears.i, views_linears.0, feature_linear, alpha_linear, rgb_linear) so reference checkpoints load.

    Inference (no_grad) runs the whole chain in ONE persistent tcgen05 kernel (csrc/nerf_mlp_tc3.cu; fp16 operands, fp32 accumulate).
    With autograd enabled (training) run_mlp is ONE autograd node whose forward and backward are UMMA kernels over fp16 tile images
    (xrnerf_b200/nerf_train.py, csrc/nerf_train.cu); `fused_train = False` (or an unsupported shape) falls back to nn.Linear under autograd."""

    def __init__(self, skips=[4], netdepth=8, netwidth=256, output_ch=4, use_viewdirs=True, netchunk=1024 * 32, embedder=None, fused=True, **kwarg):
        super().__init__()
        self.skips, self.chunk, self.use_viewdirs, self.fused = skips, netchunk, use_viewdirs, fused
        self.embedder = builder.build_embedder(embedder)
        D, W = netdepth, netwidth
        self.input_ch, self.input_ch_dirs = self.embedder.get_embed_ch()
        self.pts_linears = nn.ModuleList([nn.Linear(self.input_ch, W)] + [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + self.input_ch, W) for i in range(D - 1)])
        if self.use_viewdirs:
            self.views_linears = nn.ModuleList([nn.Linear(self.input_ch_dirs + W, W // 2)])
            self.feature_linear = nn.Linear(W, W)
            self.alpha_linear = nn.Linear(W, 1)
            self.rgb_linear = nn.Linear(W // 2, 3)
        else:
            self.output_linear = nn.Linear(W, output_ch)

    def forward(self, data):
        data = self.embedder(data)
        out = self.batchify_run_mlp(data['embedded'])
        data['raw'] = torch.reshape(out, list(data['unflatten_shape']) + [out.shape[-1]])
        del data['unflatten_shape']
        return data

    # the packed tcgen05 weight image is keyed on Parameter._version, which writes through `.data` (EMAHook swap, load_state_dict, .to()) do not bump
    def mark_dirty(self):
        self._pack_ver = None

    def _apply(self, fn, *a, **kw):
        self._pack_ver = None
        return super()._apply(fn, *a, **kw)

    def _load_from_state_dict(self, *a, **kw):
        self._pack_ver = None
        return super()._load_from_state_dict(*a, **kw)

    def _fused_ok(self, x):
        return (self.fused and not torch.is_grad_enabled() and x.is_cuda and self.use_viewdirs and len(self.pts_linears) == 8 and list(self.skips) == [4]
                and self.pts_linears[0].out_features == 256 and (self.input_ch, self.input_ch_dirs) in ((63, 27), (96, 27)))

    def _packed(self):
        import os
        self.kernel_version = int(os.environ.get('XRB_NERF_MLP_V', '3'))
        ver = tuple(p._version for p in self.parameters()) + (self.pts_linears[0].weight.device, self.kernel_version)
        if getattr(self, '_pack_ver', None) != ver:
            from ..nerf_mlp import pack_nerf_mlp, pack_nerf_mlp_v2, pack_nerf_mlp_v3
            self._pack = {1: pack_nerf_mlp, 2: pack_nerf_mlp_v2, 3: pack_nerf_mlp_v3}[self.kernel_version](self)
            self._pack_ver = ver
        return self._pack

    def batchify_run_mlp(self, x):
        if self._fused_ok(x):   # inference: the whole 12-GEMM chain in one tcgen05 kernel, no chunking needed (activations never leave the SM)
            from ..nerf_mlp import nerf_mlp_forward
            image, bias = self._packed()
            return nerf_mlp_forward(image, bias, x, self.input_ch, self.input_ch_dirs, version=self.kernel_version)
        if self.fused and getattr(self, 'fused_train', True) and torch.is_grad_enabled() and x.is_cuda and any(p.requires_grad for p in self.parameters()):
            from .. import nerf_train
            if nerf_train.supported(self):   # training: every dense layer forward + backward as UMMA kernels over fp16 tile images (csrc/nerf_train.cu), one autograd node
                return nerf_train.run_mlp_train(self, x)
        if self.chunk is None:
            return self.run_mlp(x)
        return torch.cat([self.run_mlp(x[i:i + self.chunk]) for i in range(0, x.shape[0], self.chunk)], 0)

    def run_mlp(self, x):
        input_pts, input_views = torch.split(x, [self.input_ch, self.input_ch_dirs], dim=-1)
        h = input_pts
        for i, l in enumerate(self.pts_linears):
            h = F.relu(l(h))
            if i in self.skips:
                h = torch.cat([input_pts, h], -1)
        if self.use_viewdirs:
            alpha = self.alpha_linear(h)
            h = torch.cat([self.feature_linear(h), input_views], -1)
            for l in self.views_linears:
                h = F.relu(l(h))
            return torch.cat([self.rgb_linear(h), alpha], -1)
        return self.output_linear(h)


class _ParamHolder(nn.Module):
    """gives a flat parameter the reference's state_dict key `<name>.params` (tcnn modules expose exactly one `params`)"""

    def __init__(self, p):
        super().__init__()
        self.params = p


def _hidden(network_config):
    if 'n_hidden_layers' in network_config:
        return int(network_config['n_hidden_layers'])
    return int(network_config.get('num_layers', 5))  # the reference's key is num_layers; tcnn's default would be 5


@MLPS.register_module()
class HashNerfMLP(BaseMLP):
    """state_dict keys: embedder_pos.params, embedder_dir.params (empty), density_net.params, color_net.params."""

    def __init__(self, bound=1, embedder_pos=None, embedder_dir=None, density_net=None, color_net=None, impl=1, **kwarg):
        super().__init__()
        enc = dict(embedder_pos['encoding_config'])
        if enc.get('otype') != 'HashGrid' or embedder_dir['encoding_config'].get('otype') != 'SphericalHarmonics' or int(embedder_dir['encoding_config'].get('degree', 4)) != 4:
            raise NotImplementedError('HashNerfMLP: HashGrid position encoding + SphericalHarmonics degree 4 (the reference config)')
        if int(density_net['n_output_dims']) != 16 or int(color_net['n_output_dims']) != 3:
            raise NotImplementedError('HashNerfMLP: density_net 16 outputs, color_net 3 outputs (the reference config)')
        field = NgpField(n_levels=int(enc.get('n_levels', 16)), n_features=int(enc.get('n_features_per_level', 2)), log2_hashmap_size=int(enc.get('log2_hashmap_size', 19)),
                              base_resolution=int(enc.get('base_resolution', 16)), per_level_scale=PER_LEVEL_SCALE,  # get_per_level_scale(1): ignores `bound` (Q6)
                              width=int(density_net['network_config'].get('n_neurons', 64)), density_hidden=_hidden(density_net['network_config']),
                              color_hidden=_hidden(color_net['network_config']), impl=impl)
        # not registered as a sub-module: its Parameters are the very objects held by the four tcnn-named holders below, so
        # state_dict()/load_state_dict() see exactly the reference's keys and .to()/.cuda() (in-place on .data) still moves them
        self.__dict__['field'] = field
        f = field
        self.embedder_pos = _ParamHolder(f.hash_params)
        self.embedder_dir = _ParamHolder(nn.Parameter(torch.zeros(0)))
        self.density_net = _ParamHolder(f.density_params)
        self.color_net = _ParamHolder(f.color_params)

    # the fused field caches fp16 shadows, the cell image and the UMMA weight image keyed on Parameter._version, which writes through `.data` (mmcv's EMAHook,
    # load_state_dict's copy_, .to()) do not bump: every such entry point marks the caches dirty
    def _apply(self, fn, *a, **kw):
        self.field.mark_dirty()
        return super()._apply(fn, *a, **kw)

    def _load_from_state_dict(self, *a, **kw):
        self.field.mark_dirty()
        return super()._load_from_state_dict(*a, **kw)

    def load_state_dict(self, *a, **kw):
        self.field.mark_dirty()
        return super().load_state_dict(*a, **kw)

    def mark_dirty(self):
        self.field.mark_dirty()

    def forward(self, data):
        shape = data['pts'].shape[:-1]
        out = self.run_mlp(data)
        data['raw'] = torch.reshape(out, list(shape) + [out.shape[-1]])
        return data

    def run_mlp(self, data):
        pts, viewdirs = data['pts'], data['viewdirs']
        if pts.dim() > viewdirs.dim():
            viewdirs = viewdirs[:, None].expand(pts.shape)
        pts, viewdirs = pts.reshape(-1, 3).detach(), viewdirs.reshape(-1, 3).detach()
        if pts.stride(-1) != 1:
            pts = pts.contiguous()
        if viewdirs.stride(-1) != 1:
            viewdirs = viewdirs.contiguous()
        return self.field(pts.float(), viewdirs.float())

    def run_density(self, pts_flat):
        return self.field.run_density(pts_flat.float() if pts_flat.stride(-1) == 1 else pts_flat.contiguous().float())
