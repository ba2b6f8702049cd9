"""Plain-PyTorch ResNet-18 / VGG-11-BN (CIFAR variants) for the REFERENCE arm only.

The reference's ``models.py`` has no ResNet/VGG although BASELINE.json's headline config names
ResNet-18.  To run the reference's own training/aggregation loop on that architecture, ``get_model`` is pointed at
these stock ``torch.nn`` definitions (cuDNN/cuBLAS, fp32 -- the reference has no AMP).  They are written by us, contain
none of the product's kernels/engine, and every number obtained with them is labelled as such.
"""
import torch.nn as nn
import torch.nn.functional as F


class BasicBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        out = F.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return F.relu(out + idt)


class ResNet18(nn.Module):
    BLOCKS = (2, 2, 2, 2)

    def __init__(self, num_classes=10):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 3, 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        layers, cin = [], 64
        for (cout, stride), nb in zip([(64, 1), (128, 2), (256, 2), (512, 2)], self.BLOCKS):
            layers += [BasicBlock(cin, cout, stride)] + [BasicBlock(cout, cout, 1) for _ in range(nb - 1)]
            cin = cout
        self.layers = nn.Sequential(*layers)
        self.fc = nn.Linear(512, num_classes)

    def forward(self, x):
        x = F.relu(self.bn1(self.conv1(x)))
        x = self.layers(x)
        x = x.mean(dim=(2, 3))
        return self.fc(x)


class ResNet34(ResNet18):
    BLOCKS = (3, 4, 6, 3)


class VGG11(nn.Module):
    CFG = [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"]

    def __init__(self, num_classes=10):
        super().__init__()
        mods, cin = [], 3
        for v in self.CFG:
            if v == "M":
                mods.append(nn.MaxPool2d(2, 2))
            else:
                mods += [nn.Conv2d(cin, v, 3, padding=1), nn.BatchNorm2d(v), nn.ReLU(inplace=True)]
                cin = v
        self.features = nn.Sequential(*mods)
        self.classifier = nn.Linear(512, num_classes)

    def forward(self, x):
        return self.classifier(self.features(x).flatten(1))


class VGG16(VGG11):
    CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"]
