"""ResNet forward against an implementation that shares NOTHING with the executor's front end ("a tracer
error -- SAME-padding arithmetic, jnp.mean lowering, ResNet block wiring in nets.py -- is invisible to every test").

`torch_resnet` below is written from the published architecture (He et al. 2015 with the v1.5 stride placement, as Flax's
`ResNet` / Elegy's `elegy.nets.ResNet50` define it: 7x7/2 stem with padding 3, BatchNorm, ReLU, 3x3/2 max-pool with XLA's SAME
padding, bottleneck blocks 1x1 -> 3x3 (stride) -> 1x1 with a projection shortcut where the shape changes, global average pool,
dense layer) in plain torch.nn.functional on float64: no vkjax_b200.frontend tracer, no oracle/eval_jaxpr.py, no nets.apply.
The only thing it takes from the repo is the dictionary of weights.  The oracle (CPU) and the executor (GPU, fp32-exact mode,
the reference's tolerance rtol 1e-4 / atol 1e-5 of tests/test_elegy_resnet.py:32) must both reproduce it."""
import os

import numpy as np
import pytest

from vkjax_b200 import nets
from vkjax_b200.frontend import make_jaxpr


def _same_pad(size, k, stride):
    """XLA / TensorFlow SAME padding of one spatial dimension: (lo, hi)"""
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return total // 2, total - total // 2


def torch_resnet(s, x, stage_sizes, bottleneck, t=None, as_numpy=True):
    """`t` converts a weight to a torch tensor (default: float64, no gradient); tests/test_autodiff_independent.py passes leaves that
    require gradients and as_numpy=False to differentiate the logits"""
    import torch
    import torch.nn.functional as F
    t = t or (lambda a: torch.from_numpy(np.asarray(a, np.float64)))

    def conv(x, w, stride, pad):                       # x NCHW, w HWIO -> OIHW; pad = ((top, bottom), (left, right))
        x = F.pad(x, (pad[1][0], pad[1][1], pad[0][0], pad[0][1]))
        return F.conv2d(x, t(w).permute(3, 2, 0, 1), stride=stride)

    def bn(x, p):
        c = lambda a: t(a).reshape(1, -1, 1, 1)
        return (x - c(p['mean'])) * (c(p['scale']) / torch.sqrt(c(p['var']) + 1e-5)) + c(p['offset'])

    x = torch.from_numpy(np.asarray(x, np.float64)).permute(0, 3, 1, 2)
    x = torch.relu(bn(conv(x, s['stem']['conv'], 2, ((3, 3), (3, 3))), s['stem']['bn']))
    ph, pw = _same_pad(x.shape[2], 3, 2), _same_pad(x.shape[3], 3, 2)
    x = F.max_pool2d(F.pad(x, (pw[0], pw[1], ph[0], ph[1]), value=float('-inf')), 3, 2)
    blocks = iter(s['blocks'])
    cin = x.shape[1]
    for stage, n_blocks in enumerate(stage_sizes):
        for j in range(n_blocks):
            b = next(blocks)
            stride = 2 if (stage > 0 and j == 0) else 1
            if bottleneck:
                y = torch.relu(bn(conv(x, b['conv1'], 1, ((0, 0), (0, 0))), b['bn1']))
                y = torch.relu(bn(conv(y, b['conv2'], stride, ((1, 1), (1, 1))), b['bn2']))
                y = bn(conv(y, b['conv3'], 1, ((0, 0), (0, 0))), b['bn3'])
            else:
                y = torch.relu(bn(conv(x, b['conv1'], stride, ((1, 1), (1, 1))), b['bn1']))
                y = bn(conv(y, b['conv2'], 1, ((1, 1), (1, 1))), b['bn2'])
            shortcut = x
            if stride != 1 or cin != y.shape[1]:
                assert 'proj' in b, 'the weights lack a projection where the published architecture has one'
                # XLA SAME padding of a 1x1 stride-s convolution is zero: the projection samples every s-th pixel
                shortcut = bn(conv(x, b['proj'], stride, ((0, 0), (0, 0))), b['bn_proj'])
            else:
                assert 'proj' not in b
            x = torch.relu(shortcut + y)
            cin = x.shape[1]
    assert next(blocks, None) is None
    x = x.mean(dim=(2, 3))
    y = x @ t(s['fc']['w']) + t(s['fc']['b'])
    return y.numpy() if as_numpy else y


CASES = [('resnet18', nets.ResNet18, (2, 2, 2, 2), False, (2, 64, 64, 3)),
         ('resnet50', nets.ResNet50, (3, 4, 6, 3), True, (2, 64, 64, 3)),
         ('resnet50 odd size', nets.ResNet50, (3, 4, 6, 3), True, (1, 75, 53, 3))]      # odd extents: every SAME padding is asymmetric


@pytest.mark.parametrize('name,ctor,stages,bottleneck,shape', CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_independent_torch_resnet(name, ctor, stages, bottleneck, shape, monkeypatch):
    from oracle.eval_jaxpr import eval_jaxpr
    monkeypatch.setenv('ORACLE_CONV_BACKEND', 'torch')
    m = ctor()
    s = m.init(5)
    x = np.random.default_rng(6).random(shape, np.float32)
    jaxpr = make_jaxpr(lambda x, s: m.apply(s, x))(x, s)
    from vkjax_b200 import tree_util
    y = eval_jaxpr(jaxpr, *tree_util.tree_leaves([x, s]))[0]
    ytrue = torch_resnet(s, x, stages, bottleneck)
    assert y.shape == ytrue.shape == (shape[0], 1000)
    assert np.abs(ytrue).max() > 0.1                                   # the logits carry signal
    assert np.allclose(y, ytrue, rtol=1e-4, atol=1e-5), float(np.abs(y - ytrue).max())


@pytest.mark.gpu
@pytest.mark.parametrize('name,ctor,stages,bottleneck,shape', CASES, ids=[c[0] for c in CASES])
def test_executor_matches_independent_torch_resnet(name, ctor, stages, bottleneck, shape):
    import vkjax_b200 as vkjax
    m = ctor()
    s = m.init(5)
    x = np.random.default_rng(6).random(shape, np.float32)
    ytrue = torch_resnet(s, x, stages, bottleneck)
    y = vkjax.wrap(lambda x, s: m.apply(s, x), precision='fp32')(x, s)
    assert np.allclose(y, ytrue, rtol=1e-4, atol=1e-5), float(np.abs(y - ytrue).max())      # reference tests/test_elegy_resnet.py:32
    y32 = vkjax.wrap(lambda x, s: m.apply(s, x), precision='tf32')(x, s)
    assert np.linalg.norm(y32 - ytrue) / np.linalg.norm(ytrue) < 5e-3                        # single pass: relative L2
