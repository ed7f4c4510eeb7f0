"""Registry boundary of the reference (/root/reference/xrnerf/models/builder.py:7-36): one registry aliased as
MLPS/RENDERS/EMBEDDERS/NETWORKS/SAMPLERS and build_{mlp,render,embedder,network,sampler}(cfg) = pop `type`, pass the
rest as kwargs. mmcv is not a dependency here; Registry reimplements the two calls the reference uses."""


class Registry:
    def __init__(self, name):
        self.name, self._modules = name, {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            key = name or cls.__name__
            if key in self._modules and not force:
                raise KeyError(f'{key} is already registered in {self.name}')
            self._modules[key] = cls
            return cls
        return deco(module) if module is not None else deco

    def get(self, key):
        return self._modules.get(key)

    def __contains__(self, key):
        return key in self._modules

    def build(self, cfg, **default_args):
        if cfg is None:
            return None
        cfg = dict(cfg)
        if 'type' not in cfg:
            raise KeyError(f'cfg must contain "type", got {list(cfg)}')
        t = cfg.pop('type')
        cls = self._modules.get(t) if isinstance(t, str) else t
        if cls is None:
            raise KeyError(f'{t} is not in the {self.name} registry')
        for k, v in default_args.items():
            cfg.setdefault(k, v)
        return cls(**cfg)


MODELS = Registry('models')
MLPS = RENDERS = EMBEDDERS = NETWORKS = SAMPLERS = MODELS


def build_mlp(cfg):
    return MLPS.build(cfg)


def build_render(cfg):
    return RENDERS.build(cfg)


def build_embedder(cfg):
    return EMBEDDERS.build(cfg)


def build_network(cfg):
    return NETWORKS.build(cfg)


def build_sampler(cfg):
    return SAMPLERS.build(cfg)
