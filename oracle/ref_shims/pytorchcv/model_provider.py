"""``get_model("resnet50", pretrained=...)`` of pytorchcv==0.0.65, restated as a plain ``nn.Module`` whose attribute names
are pytorchcv's (``features.init_block.conv.{conv,bn,activ}``, ``features.init_block.pool``,
``features.stageK.unitJ.body.conv{1,2,3}.{conv,bn}``, ``features.stageK.unitJ.identity_conv.{conv,bn}``) so that the
UNMODIFIED reference ``Encoder`` (model_training/model/encoders.py:9-59) builds on it and ``state_dict()`` carries the key
names of the released TorchScript checkpoint.  ResNet-50 v1 "a" variant: bottleneck factor 4, units (3,4,6,3), the stride of
a down-sampling unit sits on its FIRST 1x1 (``conv1_stride=True``; pytorchcv's ``resnet50b`` moves it to the 3x3).
ConvBlock = Conv2d(bias=False) -> BatchNorm2d(eps=1e-5) -> ReLU(inplace).  ``pretrained`` cannot be honoured offline and is
ignored (weights are loaded by the harness).
"""
import torch.nn as nn


class ConvBlock(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, activation=True):
        super().__init__()
        self.activate = activation
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=False)
        self.bn = nn.BatchNorm2d(out_channels, eps=1e-5)
        if activation:
            self.activ = nn.ReLU(inplace=True)

    def forward(self, x):
        x = self.bn(self.conv(x))
        return self.activ(x) if self.activate else x


class ResBottleneck(nn.Module):
    def __init__(self, in_channels, out_channels, stride, conv1_stride=True, bottleneck_factor=4):
        super().__init__()
        mid = out_channels // bottleneck_factor
        self.conv1 = ConvBlock(in_channels, mid, 1, stride if conv1_stride else 1, 0)
        self.conv2 = ConvBlock(mid, mid, 3, 1 if conv1_stride else stride, 1)
        self.conv3 = ConvBlock(mid, out_channels, 1, 1, 0, activation=False)

    def forward(self, x):
        return self.conv3(self.conv2(self.conv1(x)))


class ResUnit(nn.Module):
    def __init__(self, in_channels, out_channels, stride, conv1_stride=True):
        super().__init__()
        self.resize_identity = (in_channels != out_channels) or (stride != 1)
        self.body = ResBottleneck(in_channels, out_channels, stride, conv1_stride)
        if self.resize_identity:
            self.identity_conv = ConvBlock(in_channels, out_channels, 1, stride, 0, activation=False)
        self.activ = nn.ReLU(inplace=True)

    def forward(self, x):
        identity = self.identity_conv(x) if self.resize_identity else x
        x = self.body(x)
        x = x + identity
        return self.activ(x)


class ResInitBlock(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = ConvBlock(in_channels, out_channels, 7, 2, 3)
        self.pool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)

    def forward(self, x):
        return self.pool(self.conv(x))


class ResNet(nn.Module):
    def __init__(self, channels, init_block_channels=64, conv1_stride=True, in_channels=3, num_classes=1000):
        super().__init__()
        self.features = nn.Sequential()
        self.features.add_module("init_block", ResInitBlock(in_channels, init_block_channels))
        cin = init_block_channels
        for i, per_stage in enumerate(channels):
            stage = nn.Sequential()
            for j, cout in enumerate(per_stage):
                stride = 2 if (j == 0 and i != 0) else 1
                stage.add_module("unit{}".format(j + 1), ResUnit(cin, cout, stride, conv1_stride))
                cin = cout
            self.features.add_module("stage{}".format(i + 1), stage)
        self.features.add_module("final_pool", nn.AvgPool2d(kernel_size=7, stride=1))
        self.output = nn.Linear(cin, num_classes)

    def forward(self, x):
        x = self.features(x)
        return self.output(x.view(x.size(0), -1))


def get_model(name, **kwargs):
    kwargs.pop("pretrained", None)
    if name != "resnet50":
        raise ValueError("shim: only resnet50 is restated (the reference's backbone, backbone.yaml:9-15)")
    layers = (3, 4, 6, 3)
    widths = (256, 512, 1024, 2048)
    return ResNet([[w] * n for w, n in zip(widths, layers)], conv1_stride=True, **kwargs)
