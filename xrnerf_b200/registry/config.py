"""Loads the reference's Python-dict config files unchanged (what `mmcv.Config.fromfile` does for them,
/root/reference/xrnerf/core/apis/api.py:11): the file is executed and its public names become a ConfigDict."""
import os
import runpy


class ConfigDict(dict):
    """dict with attribute access, recursively (mmcv.ConfigDict behaviour the reference relies on: cfg.get, cfg.chunk ...)"""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        for key, v in list(self.items()):
            self[key] = self._wrap(v)

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = self._wrap(v)


def load_config(path, dataname=None):
    ns = runpy.run_path(os.path.abspath(path))
    cfg = {k: v for k, v in ns.items() if not k.startswith('_') and not callable(v) and not isinstance(v, type(os))}
    if dataname is not None:  # '#DATANAME#' substitution (xrnerf/core/apis/helper.py:41-49)
        def sub(v):
            if isinstance(v, str):
                return v.replace('#DATANAME#', dataname)
            if isinstance(v, dict):
                return {k: sub(x) for k, x in v.items()}
            if isinstance(v, (list, tuple)):
                return type(v)(sub(x) for x in v)
            return v
        cfg = sub(cfg)
    return ConfigDict(cfg)
