"""TEST INFRASTRUCTURE ONLY — imports the reference's own pure-PyTorch hot-path modules UNMODIFIED from /root/reference
(this container only; the GPU box has no /root/reference) so that golden vectors can be generated from them
(tests/golden/make_golden.py) and the restatement in oracle/nerf_oracle.py can be checked against them.

`import xrnerf.models` does not work headless (needs mmcv, tkinter via a stray `from turtle import forward`, lpips, imageio;
SURVEY §8c), so: fake minimal `mmcv` / `turtle` modules + namespace packages whose __path__ points into /root/reference,
then importlib the individual files. Nothing from the reference is copied.
"""
import importlib
import os
import sys
import types

REF = '/root/reference'


def available():
    return os.path.isdir(os.path.join(REF, 'xrnerf', 'models'))


class _Registry:
    def __init__(self, name, parent=None, **kw):
        self.name, self._m = name, {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self._m[name or cls.__name__] = cls
            return cls
        return deco if module is None else deco(module)

    def get(self, k):
        return self._m.get(k)

    def build(self, cfg, **kw):
        cfg = dict(cfg)
        t = cfg.pop('type')
        return self._m[t](**cfg)


class ConfigDict(dict):
    __getattr__ = dict.get


def install():
    if 'xrnerf' in sys.modules and getattr(sys.modules['xrnerf'], '_oracle_shim', False):
        return
    assert available(), '/root/reference is not present (golden vectors are generated in the build container only)'
    mmcv = types.ModuleType('mmcv')
    mmcv.ConfigDict = ConfigDict
    mmcv.Config = ConfigDict
    mmcv.is_str = lambda x: isinstance(x, str)
    utils = types.ModuleType('mmcv.utils'); utils.Registry = _Registry; utils.build_from_cfg = lambda cfg, reg, default_args=None: reg.build(cfg)
    cnn = types.ModuleType('mmcv.cnn'); cnn.MODELS = _Registry('model')
    runner = types.ModuleType('mmcv.runner'); runner.get_dist_info = lambda: (0, 1); runner.load_checkpoint = lambda *a, **k: None
    mmcv.utils, mmcv.cnn, mmcv.runner = utils, cnn, runner
    sys.modules.update({'mmcv': mmcv, 'mmcv.utils': utils, 'mmcv.cnn': cnn, 'mmcv.runner': runner})
    turtle = types.ModuleType('turtle'); turtle.forward = None; turtle.pd = None
    sys.modules.setdefault('turtle', turtle)
    for name in ['xrnerf', 'xrnerf.models', 'xrnerf.models.embedders', 'xrnerf.models.mlps', 'xrnerf.models.renders', 'xrnerf.models.networks',
                 'xrnerf.models.networks.utils', 'xrnerf.models.samplers', 'xrnerf.models.samplers.utils']:
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, *name.split('.'))]
        m._oracle_shim = True
        sys.modules[name] = m
    sys.modules['xrnerf.models'].builder = importlib.import_module('xrnerf.models.builder')


def load(name):
    """e.g. load('embedders.base') -> the reference module xrnerf/models/embedders/base.py"""
    install()
    return importlib.import_module('xrnerf.models.' + name)


def load_pipelines():
    """the reference's dataset pipeline transforms that generate samples on the hot path (GetRays/GetViewdirs/GetBounds/GetZvals/GetPts,
    datasets/pipelines/create.py; PerturbZvals, augment.py) and get_rays_np_hash (datasets/load_data/get_rays.py), imported unmodified."""
    install()
    import mmcv
    if 'mmcv.parallel' not in sys.modules:
        par = types.ModuleType('mmcv.parallel'); par.collate = lambda *a, **k: None
        sys.modules['mmcv.parallel'] = par; mmcv.parallel = par
        mmcv.utils.digit_version = lambda v: (0,)
    sys.modules.setdefault('imageio', types.ModuleType('imageio'))
    for name in ['xrnerf.datasets', 'xrnerf.datasets.pipelines', 'xrnerf.datasets.load_data', 'xrnerf.datasets.utils']:
        if name not in sys.modules:
            m = types.ModuleType(name); m.__path__ = [os.path.join(REF, *name.split('.'))]; sys.modules[name] = m
    create = importlib.import_module('xrnerf.datasets.pipelines.create')
    augment = importlib.import_module('xrnerf.datasets.pipelines.augment')
    get_rays = importlib.import_module('xrnerf.datasets.load_data.get_rays')
    return create, augment, get_rays
