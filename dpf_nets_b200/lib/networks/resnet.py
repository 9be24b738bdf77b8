"""ResNet-18 image encoder with the reference's deviations from torchvision (4-channel stem,
fc -> fc_bn -> relu tail; lib/networks/resnet.py:109-224) and the same state_dict keys.  Dense 2-D
convolutions stay on cuDNN (out of the hand-kernel scope)."""
import torch
import torch.nn as nn


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        out = self.bn2(self.conv2(self.relu(self.bn1(self.conv1(x)))))
        return self.relu(out + idt)


class ResNet(nn.Module):
    def __init__(self, layers=(2, 2, 2, 2), num_classes=1000):
        super().__init__()
        self.conv1 = nn.Conv2d(4, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        inplanes = 64
        for li, (planes, n) in enumerate(zip((64, 128, 256, 512), layers)):
            stride = 1 if li == 0 else 2
            blocks = []
            for b in range(n):
                ds = None
                if b == 0 and (stride != 1 or inplanes != planes):
                    ds = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
                blocks.append(BasicBlock(inplanes, planes, stride if b == 0 else 1, ds))
                inplanes = planes
            setattr(self, 'layer%d' % (li + 1), nn.Sequential(*blocks))
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512, num_classes)
        self.fc_bn = nn.BatchNorm1d(num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        x = torch.flatten(self.avgpool(x), 1)
        return self.relu(self.fc_bn(self.fc(x)))


def resnet18(pretrained=False, progress=True, **kwargs):
    return ResNet((2, 2, 2, 2), **kwargs)
