"""cpg_b200 -- B200-native (sm_100a) implementation of CPG's masked-convolution train/prune
hot path, behind the reference's own module API:

    cpg_b200.layers  <->  models/layers.py   (Binarizer, SharableConv2d, SharableLinear)
    cpg_b200.prune   <->  utils/prune.py     (SparsePruner)

``install()`` aliases them into ``sys.modules`` so an unmodified CPG checkout picks them up.
"""
import sys

__version__ = '0.1.0'


def torch_compat_shims(device='cuda'):
    """Keep the reference's host-side helpers running on current torch releases without editing them.

    utils/__init__.py:37-38 (``Metric.update``) accumulates ``self.sum += val * num`` into a CPU scalar tensor while
    ``val`` (the loss) lives on the GPU; torch releases that refuse the mixed-device in-place add get the value moved
    to the host first -- the same arithmetic.  Returns the list of shims applied (empty when none was needed)."""
    import torch
    applied = []
    try:
        import utils
    except Exception:
        return applied
    try:
        t = torch.tensor(0.)
        t += torch.tensor(1., device=device) * 2
    except RuntimeError:
        def update(self, val, num):
            if torch.is_tensor(val):
                val = val.detach().cpu()
            self.sum += val * num
            self.n += num
        utils.Metric.update = update
        applied.append('utils.Metric.update')
    return applied


def _view_forward(self, input):
    return input.reshape(*self.shape)


def install():
    """Make ``import models.layers`` / ``from utils.prune import SparsePruner`` resolve to this
    package inside a CPG checkout (call before importing ``utils.manager``)."""
    from . import layers, prune
    sys.modules['models.layers'] = layers
    try:
        import models  # the checkout's package
        models.layers = layers
    except Exception:
        pass
    # Conv outputs of cpg_b200.layers are physically NHWC (torch.channels_last): the reference's flatten module
    # `input.view(*self.shape)` (models/vgg.py:30-31, models/spherenet.py:20-21) then raises for feature maps
    # larger than 1x1.  reshape() keeps the logical (C, H, W) flatten order and is a view whenever view() is.
    for modname in ('models.vgg', 'models.spherenet'):
        try:
            mod = __import__(modname, fromlist=['View'])
            if hasattr(mod, 'View'):
                mod.View.forward = _view_forward
        except Exception:
            pass
    try:
        import utils.prune as ref_prune
        ref_prune.SparsePruner = prune.SparsePruner
        ref_prune.nl = layers
        import utils.manager as ref_manager
        ref_manager.SparsePruner = prune.SparsePruner
        ref_manager.nl = layers
    except Exception:
        pass
    return layers, prune
