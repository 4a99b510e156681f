"""Tiny CfgNode: attribute-access dict with clone/defrost/freeze/merge_from_file (MYCONFIG.py:12,218-314)."""
import copy
import yaml


class CfgNode(dict):
    def __init__(self, init=None):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def defrost(self):
        pass

    def freeze(self):
        pass

    def setdefault(self, k, default=None):
        if k not in self:
            self[k] = CfgNode(default) if isinstance(default, dict) else default
        return self[k]

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict):
                if k not in self or not isinstance(self[k], dict):
                    self[k] = CfgNode()
                self[k]._merge(v)
            else:
                self[k] = v

    def merge_from_file(self, path):
        with open(path, "r") as f:
            self._merge(yaml.safe_load(f) or {})
