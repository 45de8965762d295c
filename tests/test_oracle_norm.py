"""Pin the oracle's restatement of nn.BatchNorm2d -> nn.ReLU (-> nn.MaxPool2d(2, 2)) (SURVEY 8f N4,
models/vgg.py:95-122) against the stock torch modules on the CPU, float64."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import cpg_oracle as O


@pytest.mark.parametrize('train', [True, False])
@pytest.mark.parametrize('relu,pool', [(True, True), (True, False), (False, False)])
def test_bn_relu_pool_oracle_vs_torch(train, relu, pool):
    rng = np.random.RandomState(3 + int(train) + 2 * int(relu) + 4 * int(pool))
    N, C, H, W = 5, 6, 8, 6
    x = rng.standard_normal((N, C, H, W)) * 1.7 + 0.4
    gamma, beta = rng.uniform(0.5, 1.5, C), rng.standard_normal(C) * 0.3
    rm, rv = rng.standard_normal(C) * 0.2, rng.uniform(0.5, 2.0, C)
    bn = nn.BatchNorm2d(C).double()
    with torch.no_grad():
        bn.weight.copy_(torch.from_numpy(gamma)); bn.bias.copy_(torch.from_numpy(beta))
        bn.running_mean.copy_(torch.from_numpy(rm)); bn.running_var.copy_(torch.from_numpy(rv))
    bn.train(train)
    xt = torch.from_numpy(x).requires_grad_(True)
    yt = bn(xt)
    if relu:
        yt = torch.relu(yt)
    if pool:
        yt = torch.nn.functional.max_pool2d(yt, 2, 2)
    dy = rng.standard_normal(tuple(yt.shape))
    yt.backward(torch.from_numpy(dy))
    y, nrm, nrv, mean, rstd = O.bn_relu_pool_forward(x, gamma, beta, rm, rv, train, 0.1, 1e-5, relu, pool)
    dx, dg, db = O.bn_relu_pool_backward(x, dy, gamma, beta, mean, rstd, train, relu, pool)
    close = lambda a, b: np.allclose(a, b, rtol=1e-10, atol=1e-11)
    assert close(y, yt.detach().numpy())
    assert close(nrm, bn.running_mean.numpy()) and close(nrv, bn.running_var.numpy())
    assert close(dx, xt.grad.numpy()) and close(dg, bn.weight.grad.numpy()) and close(db, bn.bias.grad.numpy())


@pytest.mark.parametrize('train', [True, False])
def test_bn_add_relu_oracle_vs_torch(train):
    """The tail of a residual block (models/resnet.py:50-55, 92-98): out = bn(out); out += identity; out = relu(out)."""
    rng = np.random.RandomState(11 + int(train))
    N, C, H, W = 4, 5, 6, 7
    x = rng.standard_normal((N, C, H, W)) * 1.3 + 0.2
    r = rng.standard_normal((N, C, H, W))
    gamma, beta = rng.uniform(0.5, 1.5, C), rng.standard_normal(C) * 0.3
    rm, rv = rng.standard_normal(C) * 0.2, rng.uniform(0.5, 2.0, C)
    bn = nn.BatchNorm2d(C).double()
    with torch.no_grad():
        bn.weight.copy_(torch.from_numpy(gamma)); bn.bias.copy_(torch.from_numpy(beta))
        bn.running_mean.copy_(torch.from_numpy(rm)); bn.running_var.copy_(torch.from_numpy(rv))
    bn.train(train)
    relu = nn.ReLU(inplace=True)
    xt, rt = torch.from_numpy(x).requires_grad_(True), torch.from_numpy(r).requires_grad_(True)
    out = bn(xt)
    out += rt
    yt = relu(out)
    dy = rng.standard_normal(tuple(yt.shape))
    yt.backward(torch.from_numpy(dy))
    y, nrm, nrv, mean, rstd = O.bn_add_relu_forward(x, r, gamma, beta, rm, rv, train, 0.1, 1e-5)
    dx, dres, dg, db = O.bn_add_relu_backward(x, y, dy, gamma, beta, mean, rstd, train)
    close = lambda a, b: np.allclose(a, b, rtol=1e-10, atol=1e-11)
    assert close(y, yt.detach().numpy())
    assert close(nrm, bn.running_mean.numpy()) and close(nrv, bn.running_var.numpy())
    assert close(dx, xt.grad.numpy()) and close(dres, rt.grad.numpy())
    assert close(dg, bn.weight.grad.numpy()) and close(db, bn.bias.grad.numpy())
