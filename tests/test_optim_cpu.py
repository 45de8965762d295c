"""Host-side behaviour of cpg_b200.optim that needs no GPU: the constructor accepts exactly the configurations the
reference builds (CPG_cifar100_main_normal.py:339-346) and there is no CPU path."""
import pytest
import torch


def test_constructor_contract():
    from cpg_b200.optim import SGD, Adam
    p = torch.nn.Parameter(torch.zeros(4))
    o = SGD([p], lr=1e-2, weight_decay=0.0, momentum=0.9, nesterov=True)         # the reference's call, verbatim
    assert o.param_groups[0]['lr'] == 1e-2 and o.param_groups[0]['momentum'] == 0.9 and o.param_groups[0]['nesterov']
    a = Adam([p], lr=5e-4)
    assert a.param_groups[0]['betas'] == (0.9, 0.999) and a.param_groups[0]['eps'] == 1e-8
    for bad in (dict(momentum=0.0), dict(nesterov=False), dict(weight_decay=1e-4), dict(dampening=0.1)):
        with pytest.raises(ValueError):
            SGD([p], lr=0.1, **bad)
    for bad in (dict(amsgrad=True), dict(weight_decay=1e-4), dict(lr=-1.0)):
        with pytest.raises(ValueError):
            Adam([p], **bad)
    # the learning-rate schedule of the reference writes param_group['lr'] (utils/__init__.py Optimizers)
    o.param_groups[0]['lr'] = 1e-3
    assert o.state_dict()['param_groups'][0]['lr'] == 1e-3


def test_no_cpu_path():
    from cpg_b200 import _lib
    from cpg_b200.optim import SGD, Adam
    for cls in (SGD, Adam):
        p = torch.nn.Parameter(torch.zeros(4))
        p.grad = torch.zeros(4)
        with pytest.raises(_lib.CpgbError):
            cls([p], lr=0.1).step()
        q = torch.nn.Parameter(torch.zeros(4))          # no gradient: nothing to do, nothing raised
        cls([q], lr=0.1).step()
