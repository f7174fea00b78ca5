import functools


class CfgNode(dict):
    """Attribute dict (lenient stand-in for yacs CfgNode)."""

    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _called_with_cfg(*args, **kwargs):
    if len(args) and isinstance(args[0], CfgNode):
        return True
    return isinstance(kwargs.get("cfg", None), CfgNode)


def configurable(init_func=None, *, from_config=None):
    """Same call contract as detectron2.config.configurable for __init__."""
    assert init_func is not None and from_config is None

    @functools.wraps(init_func)
    def wrapped(self, *args, **kwargs):
        fc = type(self).from_config
        if _called_with_cfg(*args, **kwargs):
            init_func(self, **fc(*args, **kwargs))
        else:
            init_func(self, *args, **kwargs)

    return wrapped


def get_cfg():
    return CfgNode()
