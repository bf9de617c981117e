"""ResNet-18 encoder without the final FC, B200-native.

Drop-in for the reference's models/resnet.py (ResNet 124-216, BasicBlock 39-77, resnet18 228-236): same
constructor arguments, module hierarchy and state_dict keys (OIHW fp32 conv weights, BatchNorm
weight/bias/running_mean/running_var/num_batches_tracked), same initialisation (kaiming-normal fan_out convs,
unit BatchNorm, reference lines 160-165).  The nn.Conv2d / nn.BatchNorm2d children are parameter CONTAINERS
only: forward() hands the whole encoder to libstraps_b200 (tcgen05 implicit-GEMM convolutions with folded
BatchNorm, residual add and ReLU in the epilogue; see csrc/conv_tc.cu, csrc/regressor.cu).  Only the ResNet-18
configuration of the reference is in scope.
"""
import torch.nn as nn

from straps_b200.engine import RegressorEngine, needs_grad

__all__ = ['ResNet', 'BasicBlock', 'resnet18']


class BasicBlock(nn.Module):
    """Parameter container for conv3x3-BN-ReLU-conv3x3-BN (+ 1x1/2 downsample) -- reference resnet.py:39-77."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super(BasicBlock, self).__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        raise RuntimeError('BasicBlock is a parameter container on the B200 path; call the enclosing ResNet')


class ResNet(nn.Module):
    def __init__(self, block, layers, in_channels, **kwargs):
        super(ResNet, self).__init__()
        if block is not BasicBlock or list(layers) != [2, 2, 2, 2] or kwargs:
            raise NotImplementedError('only the ResNet-18 configuration of the reference is built on the B200 path')
        self.inplanes = 64
        self.conv1 = nn.Conv2d(in_channels, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._stage(64, 1)
        self.layer2 = self._stage(128, 2)
        self.layer3 = self._stage(256, 2)
        self.layer4 = self._stage(512, 2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self._engine = RegressorEngine(encoder=self)

    def _stage(self, planes, stride):
        down = None
        if stride != 1 or self.inplanes != planes:
            down = nn.Sequential(nn.Conv2d(self.inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
        blocks = [BasicBlock(self.inplanes, planes, stride, down), BasicBlock(planes, planes)]
        self.inplanes = planes
        return nn.Sequential(*blocks)

    def forward(self, x):
        """[B, C, 256, 256] -> [B, 512] (reference resnet.py:201-216)."""
        if needs_grad(self):
            return self._engine.forward_train(x, 0, want='feat')
        if self.training:
            return self._engine.forward_batch_stats(x, 0, want='feat')
        return self._engine.encoder_forward(x)


def resnet18(in_channels, pretrained=False, progress=True, **kwargs):
    if pretrained:
        raise NotImplementedError('no network access: pretrained ImageNet weights cannot be downloaded')
    return ResNet(BasicBlock, [2, 2, 2, 2], in_channels, **kwargs)
