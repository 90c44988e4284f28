"""dotdict, as vae/utils.py:3-7 (missing keys read as None)."""


class dotdict(dict):
    __getattr__ = dict.get
    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__
