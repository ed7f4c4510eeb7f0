"""Host-side mirror of the reference's model API (xrnerf/models): same registry names, constructor kwargs, data-dict keys and
state_dict keys, so the reference's config files for NeRF / Mip-NeRF / Instant-NGP build unchanged (`build_network(cfg.model)`)."""
from .builder import (EMBEDDERS, MLPS, MODELS, NETWORKS, RENDERS, SAMPLERS, Registry, build_embedder, build_mlp, build_network, build_render, build_sampler)
from .config import ConfigDict, load_config
from . import embedders, mlps, renders, samplers, networks  # noqa: F401  (registration side effects)
from .embedders import BaseEmbedder, MipNerfEmbedder
from .mlps import HashNerfMLP, NerfMLP
from .renders import HashNerfRender, MipNerfRender, NerfRender
from .samplers import NGPGridSampler
from .networks import HashNerfNetwork, MipNerfNetwork, NerfNetwork

__all__ = ['MODELS', 'MLPS', 'RENDERS', 'EMBEDDERS', 'NETWORKS', 'SAMPLERS', 'Registry', 'build_mlp', 'build_render', 'build_embedder', 'build_network', 'build_sampler',
           'ConfigDict', 'load_config', 'BaseEmbedder', 'MipNerfEmbedder', 'NerfMLP', 'HashNerfMLP', 'NerfRender', 'MipNerfRender', 'HashNerfRender', 'NGPGridSampler',
           'NerfNetwork', 'MipNerfNetwork', 'HashNerfNetwork']
