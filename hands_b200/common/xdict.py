"""Minimal `xdict`: the dict container the reference's hot path returns (common/xdict.py:26-288).

Only the behaviour the path and its callers rely on: duplicate keys are an error on assignment
(`xdict.py:50-55`), `postfix/prefix/merge/detach/to/search/subset/overwrite`.
"""
import torch


class xdict(dict):
    def __init__(self, mydict=None):
        super().__init__()
        if mydict is not None:
            for k, v in mydict.items():
                super().__setitem__(k, v)

    def __setitem__(self, key, val):
        assert key not in self.keys(), f"Key already exists {key}"
        super().__setitem__(key, val)

    def overwrite(self, k, v):
        super().__setitem__(k, v)

    def subset(self, keys):
        return xdict({k: self[k] for k in keys})

    def search(self, keyword, replace_to=None):
        out = xdict()
        for k, v in self.items():
            if keyword in k:
                out[k if replace_to is None else k.replace(keyword, replace_to)] = v
        return out

    def merge(self, dict2):
        mine, theirs = set(self.keys()), set(dict2.keys())
        assert not (mine & theirs), f"Merge failed: duplicate keys ({mine & theirs})"
        for k, v in dict2.items():
            self[k] = v

    def prefix(self, text):
        return xdict({text + k: v for k, v in self.items()})

    def postfix(self, text):
        return xdict({k + text: v for k, v in self.items()})

    def replace_keys(self, str_src, str_tar):
        return xdict({k.replace(str_src, str_tar): v for k, v in self.items()})

    def sorted_keys(self):
        return sorted(self.keys())

    def to(self, dev):
        return xdict({k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in self.items()})

    def detach(self):
        return xdict({k: (v.cpu().detach() if isinstance(v, torch.Tensor) else v) for k, v in self.items()})

    def has_invalid(self):
        for k, v in self.items():
            if isinstance(v, torch.Tensor) and (torch.isnan(v).any() or torch.isinf(v).any()):
                return True
        return False
