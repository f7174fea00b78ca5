"""Stand-in for detectron2's `build_resnet_backbone` (detectron2/modeling/backbone/resnet.py), which the reference selects
with MODEL.BACKBONE.NAME: "build_resnet_backbone" (configs/cityscapes/semantic-segmentation/
Base-Cityscapes-SemanticSegmentation.yaml:4) but does NOT vendor, and whose version it does not pin (INSTALL.md:13-19).

TEST INFRASTRUCTURE, PARITY UNPINNED: this is a restatement of the published architecture (BasicStem: 7x7/2 conv + norm +
ReLU + 3x3/2 max-pool; BottleneckBlock with the stride in the 3x3 conv, RESNETS.STRIDE_IN_1X1: False; eval-mode batch
norm), with detectron2's module / state_dict names (stem.conv1[.norm], res{2..5}.{i}.conv{1,2,3}[.norm], .shortcut[.norm]),
i.e. the names tools/convert-torchvision-to-d2.py:33-44 produces.  tests/test_resnet_oracle.py checks it against
torchvision.models.resnet50/101 under exactly that key mapping."""
import torch.nn.functional as F
from torch import nn

from detectron2.layers import Conv2d, ShapeSpec
from detectron2.modeling import BACKBONE_REGISTRY, Backbone

BLOCKS = {50: [3, 4, 6, 3], 101: [3, 4, 23, 3]}


def _bn(c):
    return nn.BatchNorm2d(c)            # "SyncBN" / "BN" / "FrozenBN" are the same affine map in eval mode


class BottleneckBlock(nn.Module):
    def __init__(self, cin, cout, width, stride):
        super().__init__()
        self.shortcut = Conv2d(cin, cout, kernel_size=1, stride=stride, bias=False, norm=_bn(cout)) if cin != cout else None
        self.conv1 = Conv2d(cin, width, kernel_size=1, stride=1, bias=False, norm=_bn(width))
        self.conv2 = Conv2d(width, width, kernel_size=3, stride=stride, padding=1, bias=False, norm=_bn(width))
        self.conv3 = Conv2d(width, cout, kernel_size=1, bias=False, norm=_bn(cout))

    def forward(self, x):
        out = F.relu(self.conv1(x))
        out = F.relu(self.conv2(out))
        out = self.conv3(out)
        sc = self.shortcut(x) if self.shortcut is not None else x
        return F.relu(out + sc)


class BasicStem(nn.Module):
    def __init__(self, cin=3, cout=64):
        super().__init__()
        self.conv1 = Conv2d(cin, cout, kernel_size=7, stride=2, padding=3, bias=False, norm=_bn(cout))

    def forward(self, x):
        return F.max_pool2d(F.relu(self.conv1(x)), kernel_size=3, stride=2, padding=1)


class ResNet(Backbone):
    def __init__(self, depth, out_features):
        super().__init__()
        self.stem = BasicStem()
        self._out_features = list(out_features)
        self._out_feature_channels, self._out_feature_strides = {}, {}
        cin, width, stride = 64, 64, 4
        for i, n in enumerate(BLOCKS[depth]):
            cout = width * 4
            blocks = []
            for j in range(n):
                blocks.append(BottleneckBlock(cin, cout, width, 2 if (j == 0 and i > 0) else 1))
                cin = cout
            if i > 0:
                stride *= 2
            name = f"res{i + 2}"
            self.add_module(name, nn.Sequential(*blocks))
            self._out_feature_channels[name], self._out_feature_strides[name] = cout, stride
            width *= 2

    def forward(self, x):
        outs = {}
        x = self.stem(x)
        for name in ("res2", "res3", "res4", "res5"):
            x = getattr(self, name)(x)
            if name in self._out_features:
                outs[name] = x
        return outs

    def output_shape(self):
        return {n: ShapeSpec(channels=self._out_feature_channels[n], stride=self._out_feature_strides[n]) for n in self._out_features}


@BACKBONE_REGISTRY.register()
def build_resnet_backbone(cfg, input_shape):
    r = cfg.MODEL.RESNETS
    assert not r.STRIDE_IN_1X1, "the Cityscapes configs use STRIDE_IN_1X1: False"
    return ResNet(int(r.DEPTH), list(r.OUT_FEATURES))
