"""VGG16-BN-cifar harness used by bench.py, smoke() and the trajectory tests.

Mirrors the *module structure* of the reference's ``custom_vgg_cifar100``
(models/vgg.py:33-122, cfg at CPG_cifar100_main_normal.py:188) so that the mask
keys (``module.features.<idx>``), ``datasets`` / ``classifiers`` / ``add_dataset`` /
``set_dataset`` and therefore SparsePruner / Manager work unchanged.  The layer
classes are injected (``nl``): ``cpg_b200.layers`` for the product, the oracle's
``OracleSharable*`` for the CPU arm, or the reference's own ``models.layers``.

This is harness code, not part of the accelerated path.
"""
import torch
import torch.nn as nn

VGG16_CIFAR_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M']


class View(nn.Module):
    def __init__(self, *shape):
        super().__init__()
        self.shape = shape

    def forward(self, x):
        return x.reshape(*self.shape)


def make_features(conv_cls, linear_cls, cfg=VGG16_CIFAR_CFG, width=1.0, batch_norm=True):
    layers, in_ch = [], 3
    for v in cfg:
        if v == 'M':
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
            continue
        out_ch = int(v * width)
        layers.append(conv_cls(in_ch, out_ch, kernel_size=3, padding=1, bias=False))
        if batch_norm:
            layers.append(nn.BatchNorm2d(out_ch))
        layers.append(nn.ReLU(inplace=True))
        in_ch = out_ch
    layers += [View(-1, int(512 * width)),
               linear_cls(int(512 * width), int(4096 * width)), nn.ReLU(True),
               linear_cls(int(4096 * width), int(4096 * width)), nn.ReLU(True)]
    return nn.Sequential(*layers)


class VGGCifar(nn.Module):
    """Same attribute surface as models/vgg.py:33-93."""

    def __init__(self, conv_cls, linear_cls, width=1.0, cfg=VGG16_CIFAR_CFG):
        super().__init__()
        self._conv_cls, self._linear_cls = conv_cls, linear_cls
        self.features = make_features(conv_cls, linear_cls, cfg, width)
        self.network_width_multiplier = width
        self.datasets, self.classifiers = [], nn.ModuleList()
        self.dataset2num_classes = {}
        self.classifier = None
        self._initialize_weights()

    def _initialize_weights(self):
        # models/vgg.py:59-70
        for m in self.modules():
            if isinstance(m, self._conv_cls):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, self._linear_cls):
                nn.init.normal_(m.weight, 0, 0.01)
                nn.init.constant_(m.bias, 0)

    def add_dataset(self, dataset, num_classes):
        if dataset not in self.datasets:
            self.datasets.append(dataset)
            self.dataset2num_classes[dataset] = num_classes
            head = nn.Linear(int(4096 * self.network_width_multiplier), num_classes)
            nn.init.normal_(head.weight, 0, 0.01)
            nn.init.constant_(head.bias, 0)
            self.classifiers.append(head)

    def set_dataset(self, dataset):
        assert dataset in self.datasets
        self.classifier = self.classifiers[self.datasets.index(dataset)]

    def forward(self, x):
        return self.classifier(self.features(x))


def sharable_layers(model, conv_cls, linear_cls):
    return [(n, m) for n, m in model.named_modules() if isinstance(m, (conv_cls, linear_cls))]


def fill_params_deterministic(model, seed):
    """Platform-independent parameter fill (numpy RandomState), in named_parameters order.
    Used so golden trajectories do not depend on torch's RNG stream."""
    import numpy as np
    rng = np.random.RandomState(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if 'piggymask' in name:
                continue
            if p.dim() >= 2:
                fan = p[0].numel()
                a = rng.standard_normal(p.shape).astype('float32') * (2.0 / fan) ** 0.5
            elif name.endswith('weight'):
                a = (1.0 + 0.1 * rng.standard_normal(p.shape)).astype('float32')
            else:
                a = (0.05 * rng.standard_normal(p.shape)).astype('float32')
            p.copy_(torch.from_numpy(a))
