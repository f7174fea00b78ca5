import collections.abc
from itertools import repeat

from torch import nn
from torch.nn.init import trunc_normal_ as _tn


class DropPath(nn.Module):
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        assert not self.training, "oracle shim: eval only"
        return x


def to_2tuple(x):
    if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
        return tuple(x)
    return tuple(repeat(x, 2))


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return _tn(tensor, mean=mean, std=std, a=a, b=b)
