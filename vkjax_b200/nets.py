"""Self-contained model definitions that trace to the primitive sequences Elegy 0.7.1 models emit
 -- Elegy itself is not installable in this image.

  ResNet (v1.5 bottleneck / basic)  ≙ elegy.nets.ResNet50 / ResNet18 (reference README.md:34-43,
                                       tests/test_elegy_resnet.py)
  MLP  (LeNet-300-100)              ≙ reference tests/test_elegy_mlp.py:14-33
  ConvNet                           ≙ reference tests/test_elegy_conv.py:14-23

A module is a pair of functions: `init(rng) -> states` (numpy, synthetic random init as specified in
SURVEY.md §8d) and `apply(states, x) -> y` (written against vkjax_b200.frontend, i.e. traceable by
`vkjax.wrap`).  NHWC activations, HWIO kernels, inference-mode BatchNorm on running statistics:
  inv = scale * rsqrt(var + eps);  y = (x - mean) * inv + offset     (Haiku/Elegy BatchNormalization)
"""
import numpy as np

from .core import ConvDimensionNumbers
from .frontend import lax, jnp, nn

NHWC_HWIO = ConvDimensionNumbers((0, 3, 1, 2), (3, 2, 0, 1), (0, 3, 1, 2))
BN_EPS = 1e-5


def conv2d(x, w, stride=1, padding='SAME'):
    return lax.conv_general_dilated(x, w, (stride, stride), padding, dimension_numbers=NHWC_HWIO)


def batch_norm(x, p):
    inv = p['scale'] * lax.rsqrt(p['var'] + BN_EPS)
    return (x - p['mean']) * inv + p['offset']


def linear(x, p):
    y = jnp.dot(x, p['w'])
    return y + jnp.broadcast_to(p['b'], y.shape)


def _conv_init(rng, kh, kw, cin, cout):
    return rng.normal(0.0, np.sqrt(2.0 / (kh * kw * cin)), (kh, kw, cin, cout)).astype(np.float32)   # He


def _bn_init(rng, c, last=False):
    """Synthetic BatchNorm state.  `last` = the BatchNorm that closes a residual branch: Elegy / Flax initialise its scale
    to ZERO so that a fresh block is the identity; to keep the branch in the computation (a zero scale would make the
    parity tests blind to conv3) it gets a small scale U(0.1, 0.4) instead.  With scale 1 there (round 1) the residual
    sums doubled the variance per block and random-init logits reached +-6.7e3, where the reference's
    rtol 1e-4 / atol 1e-5 (tests/test_elegy_resnet.py:32) says nothing; now they are O(1)."""
    scale = rng.uniform(0.1, 0.4, (1, 1, 1, c)).astype(np.float32) if last else np.ones((1, 1, 1, c), np.float32)
    return {'scale': scale, 'offset': np.zeros((1, 1, 1, c), np.float32),
            'mean': rng.normal(0.0, 0.1, (1, 1, 1, c)).astype(np.float32),
            'var': rng.uniform(0.5, 1.5, (1, 1, 1, c)).astype(np.float32)}


# ---- traced initialisers (vkModel.call_init_step: run on the device through frontend.random) ----------------------------
# Haiku / Elegy defaults: Linear and Conv2D kernels ~ TruncatedNormal(stddev = 1 / sqrt(fan_in)) (truncated at +-2 sigma),
# biases 0; reference tests/test_elegy_mlp.py:30 gives the last Linear a RandomNormal bias.
class _KeyStream:
    """key, sub = split(key) per request (how Haiku's hk.next_rng_key() walks its stream)"""
    def __init__(self, key):
        self.key = key

    def next(self):
        from .frontend import random, lax
        ks = random.split(self.key, 2)
        self.key = lax.reshape(lax.slice(ks, (0, 0), (1, 2)), (2,))
        return lax.reshape(lax.slice(ks, (1, 0), (2, 2)), (2,))


def _trunc_normal(key, shape, fan_in):
    from .frontend import random, lax
    return lax.mul(random.truncated_normal(key, -2.0, 2.0, shape), np.float32(1.0 / np.sqrt(fan_in)))


def _zeros_traced(shape):
    return jnp.broadcast_to(np.float32(0.0), tuple(shape))


class ResNet:
    """ResNet v1.5: stride on the 3x3 of each bottleneck; projection shortcut when the shape changes."""
    def __init__(self, stage_sizes, bottleneck=True, num_classes=1000, width=64, name='resnet'):
        self.stage_sizes, self.bottleneck, self.num_classes, self.width, self.name = \
            tuple(stage_sizes), bottleneck, num_classes, width, name

    def init(self, rng, in_channels=3):
        rng = np.random.default_rng(rng) if not isinstance(rng, np.random.Generator) else rng
        w = self.width
        s = {'stem': {'conv': _conv_init(rng, 7, 7, in_channels, w), 'bn': _bn_init(rng, w)}, 'blocks': []}
        cin = w
        for i, n_blocks in enumerate(self.stage_sizes):
            f = w * 2 ** i
            cout = f * 4 if self.bottleneck else f
            for j in range(n_blocks):
                stride = 2 if (i > 0 and j == 0) else 1
                b = {}
                if self.bottleneck:
                    b['conv1'], b['bn1'] = _conv_init(rng, 1, 1, cin, f), _bn_init(rng, f)
                    b['conv2'], b['bn2'] = _conv_init(rng, 3, 3, f, f), _bn_init(rng, f)
                    b['conv3'], b['bn3'] = _conv_init(rng, 1, 1, f, cout), _bn_init(rng, cout, last=True)
                else:
                    b['conv1'], b['bn1'] = _conv_init(rng, 3, 3, cin, f), _bn_init(rng, f)
                    b['conv2'], b['bn2'] = _conv_init(rng, 3, 3, f, f), _bn_init(rng, f, last=not self.bottleneck)
                if stride != 1 or cin != cout:
                    b['proj'], b['bn_proj'] = _conv_init(rng, 1, 1, cin, cout), _bn_init(rng, cout)
                s['blocks'].append(b)
                cin = cout
        s['fc'] = {'w': rng.normal(0.0, np.sqrt(1.0 / cin), (cin, self.num_classes)).astype(np.float32),
                   'b': np.zeros((self.num_classes,), np.float32)}
        return s

    def init_traced(self, key, x):
        """The same state tree as init(), drawn on the device: conv kernels He-normal via threefry / erf_inv chains,
        BatchNorm scale 1 / offset 0, running mean ~ N(0, 0.1), running var ~ U(0.5, 1.5) (synthetic statistics)."""
        from .frontend import random, lax
        ks = _KeyStream(key)
        conv = lambda kh, kw, cin, cout: lax.mul(random.normal(ks.next(), (kh, kw, cin, cout)), np.float32(np.sqrt(2.0 / (kh * kw * cin))))

        def bn(c, last=False):
            scale = random.uniform(ks.next(), (1, 1, 1, c), np.float32, 0.1, 0.4) if last else jnp.broadcast_to(np.float32(1.0), (1, 1, 1, c))
            return {'scale': scale, 'offset': _zeros_traced((1, 1, 1, c)),
                    'mean': lax.mul(random.normal(ks.next(), (1, 1, 1, c)), np.float32(0.1)),
                    'var': random.uniform(ks.next(), (1, 1, 1, c), np.float32, 0.5, 1.5)}
        w = self.width
        s = {'stem': {'conv': conv(7, 7, x.shape[-1], w), 'bn': bn(w)}, 'blocks': []}
        cin = w
        for i, n_blocks in enumerate(self.stage_sizes):
            f = w * 2 ** i
            cout = f * 4 if self.bottleneck else f
            for j in range(n_blocks):
                stride = 2 if (i > 0 and j == 0) else 1
                b = {}
                if self.bottleneck:
                    b['conv1'], b['bn1'] = conv(1, 1, cin, f), bn(f)
                    b['conv2'], b['bn2'] = conv(3, 3, f, f), bn(f)
                    b['conv3'], b['bn3'] = conv(1, 1, f, cout), bn(cout, last=True)
                else:
                    b['conv1'], b['bn1'] = conv(3, 3, cin, f), bn(f)
                    b['conv2'], b['bn2'] = conv(3, 3, f, f), bn(f, last=not self.bottleneck)
                if stride != 1 or cin != cout:
                    b['proj'], b['bn_proj'] = conv(1, 1, cin, cout), bn(cout)
                s['blocks'].append(b)
                cin = cout
        s['fc'] = {'w': _trunc_normal(ks.next(), (cin, self.num_classes), cin), 'b': _zeros_traced((self.num_classes,))}
        return s

    def block_strides(self):
        return [2 if (i > 0 and j == 0) else 1 for i, n in enumerate(self.stage_sizes) for j in range(n)]

    def apply(self, s, x):
        x = conv2d(x, s['stem']['conv'], 2, [(3, 3), (3, 3)])
        x = nn.relu(batch_norm(x, s['stem']['bn']))
        x = lax.reduce_window(x, -jnp.inf, lax.max, (1, 3, 3, 1), (1, 2, 2, 1), 'SAME')
        for b, stride in zip(s['blocks'], self.block_strides()):
            residual = x
            if self.bottleneck:
                y = nn.relu(batch_norm(conv2d(x, b['conv1'], 1), b['bn1']))
                y = nn.relu(batch_norm(conv2d(y, b['conv2'], stride, [(1, 1), (1, 1)]), b['bn2']))
                y = batch_norm(conv2d(y, b['conv3'], 1), b['bn3'])
            else:
                y = nn.relu(batch_norm(conv2d(x, b['conv1'], stride, [(1, 1), (1, 1)]), b['bn1']))
                y = batch_norm(conv2d(y, b['conv2'], 1, [(1, 1), (1, 1)]), b['bn2'])
            if 'proj' in b:
                residual = batch_norm(conv2d(x, b['proj'], stride), b['bn_proj'])
            x = nn.relu(residual + y)
        x = jnp.mean(x, axis=(1, 2))
        return linear(x, s['fc'])

    def conv_flops(self, batch, hw=224):
        """2*M*N*K summed over every conv (+ the FC), for the roofline."""
        total, rows = 0, []
        def add(name, h, k, cin, cout, stride):
            nonlocal total
            oh = -(-h // stride)
            fl = 2 * batch * oh * oh * cout * k * k * cin
            by = 4 * (batch * h * h * cin + k * k * cin * cout + batch * oh * oh * cout)
            rows.append((name, batch * oh * oh, cout, k * k * cin, fl, by))
            total += fl
            return oh
        h = add('stem', hw, 7, 3, self.width, 2)
        h = -(-h // 2)
        cin = self.width
        for i, n_blocks in enumerate(self.stage_sizes):
            f = self.width * 2 ** i
            cout = f * 4 if self.bottleneck else f
            for j in range(n_blocks):
                stride = 2 if (i > 0 and j == 0) else 1
                if self.bottleneck:
                    add(f's{i}b{j}.conv1', h, 1, cin, f, 1)
                    h2 = add(f's{i}b{j}.conv2', h, 3, f, f, stride)
                    add(f's{i}b{j}.conv3', h2, 1, f, cout, 1)
                else:
                    h2 = add(f's{i}b{j}.conv1', h, 3, cin, f, stride)
                    add(f's{i}b{j}.conv2', h2, 3, f, f, 1)
                if stride != 1 or cin != cout:
                    add(f's{i}b{j}.proj', h, 1, cin, cout, stride)
                h, cin = h2, cout
        rows.append(('fc', batch, self.num_classes, cin, 2 * batch * cin * self.num_classes,
                     4 * (batch * cin + cin * self.num_classes + batch * self.num_classes)))
        total += rows[-1][4]
        return total, rows


def ResNet50(**kw):
    return ResNet([3, 4, 6, 3], bottleneck=True, name='resnet50', **kw)


def ResNet18(**kw):
    return ResNet([2, 2, 2, 2], bottleneck=False, name='resnet18', **kw)


class MLP:
    """≙ reference tests/test_elegy_mlp.py:14-33: image/255 -> Flatten -> 300 -> relu -> 100 -> relu -> 10."""
    def __init__(self, n1=300, n2=100, n_out=10, name='mlp'):
        self.sizes, self.name = (n1, n2, n_out), name

    def init(self, rng, in_features=32 * 32 * 3):
        rng = np.random.default_rng(rng) if not isinstance(rng, np.random.Generator) else rng
        s, fan_in = [], in_features
        for i, n in enumerate(self.sizes):
            last = i == len(self.sizes) - 1
            s.append({'w': np.clip(rng.normal(0, np.sqrt(1.0 / fan_in), (fan_in, n)), -2 / np.sqrt(fan_in), 2 / np.sqrt(fan_in)).astype(np.float32),
                      'b': rng.normal(0, 1, (n,)).astype(np.float32) if last else np.zeros((n,), np.float32)})
            fan_in = n
        return s

    def init_traced(self, key, x):
        """≙ the Elegy MLP of reference tests/test_elegy_mlp.py:14-33 initialised by call_init_step: truncated-normal
        kernels, zero biases, RandomNormal bias on the last layer."""
        from .frontend import random
        ks = _KeyStream(key)
        fan_in = int(np.prod(x.shape[1:], dtype=np.int64))
        s = []
        for i, n in enumerate(self.sizes):
            last = i == len(self.sizes) - 1
            s.append({'w': _trunc_normal(ks.next(), (fan_in, n), fan_in),
                      'b': random.normal(ks.next(), (n,)) if last else _zeros_traced((n,))})
            fan_in = n
        return s

    def apply(self, s, image):
        x = image.astype(jnp.float32) / 255.0
        x = x.reshape(x.shape[0], -1)
        for i, layer in enumerate(s):
            x = linear(x, layer)
            if i < len(s) - 1:
                x = nn.relu(x)
        return x


class ConvNet:
    """≙ reference tests/test_elegy_conv.py:14-23: 2 x (Conv2D 32, 3x3, stride 2 + bias, ReLU) -> Flatten -> Linear(10)."""
    name = 'convnet'

    def init(self, rng, in_shape=(32, 32, 3)):
        rng = np.random.default_rng(rng) if not isinstance(rng, np.random.Generator) else rng
        h, w, c = in_shape
        s = {'c1': {'w': _conv_init(rng, 3, 3, c, 32), 'b': np.zeros((32,), np.float32)},
             'c2': {'w': _conv_init(rng, 3, 3, 32, 32), 'b': np.zeros((32,), np.float32)}}
        oh, ow = -(-(-(-h // 2)) // 2), -(-(-(-w // 2)) // 2)
        n = oh * ow * 32
        s['fc'] = {'w': rng.normal(0, np.sqrt(1.0 / n), (n, 10)).astype(np.float32), 'b': np.zeros((10,), np.float32)}
        return s

    def init_traced(self, key, x):
        ks = _KeyStream(key)
        h, w, c = x.shape[1:]
        s = {'c1': {'w': _trunc_normal(ks.next(), (3, 3, c, 32), 9 * c), 'b': _zeros_traced((32,))},
             'c2': {'w': _trunc_normal(ks.next(), (3, 3, 32, 32), 9 * 32), 'b': _zeros_traced((32,))}}
        n = -(-(-(-h // 2)) // 2) * -(-(-(-w // 2)) // 2) * 32
        s['fc'] = {'w': _trunc_normal(ks.next(), (n, 10), n), 'b': _zeros_traced((10,))}
        return s

    def apply(self, s, x):
        for k in ('c1', 'c2'):
            y = conv2d(x, s[k]['w'], 2, 'SAME')
            x = nn.relu(y + jnp.broadcast_to(s[k]['b'], y.shape))
        x = x.reshape(x.shape[0], -1)
        return linear(x, s['fc'])
