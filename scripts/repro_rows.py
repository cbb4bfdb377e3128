"""Repro / unit check of the B2J_CT_ROWS stem path: python scripts/repro_rows.py [case ...]"""
import sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import vkjax_b200 as vkjax
from vkjax_b200.frontend import lax, jnp
from vkjax_b200.core import ConvDimensionNumbers
from common import oracle

dn = ConvDimensionNumbers((0, 3, 1, 2), (3, 2, 0, 1), (0, 3, 1, 2))


def stem(x, w):
    return lax.conv_general_dilated(x, w, (2, 2), 'SAME', dimension_numbers=dn)


def stem_u8(x, w):
    return lax.conv_general_dilated(x.astype(jnp.float32) / 255.0, w, (2, 2), 'SAME', dimension_numbers=dn)


cases = sys.argv[1:] or ['f32_small', 'f32_224', 'u8_small', 'u8_224']
rng = np.random.default_rng(0)
w = rng.normal(0, 0.1, (7, 7, 3, 64)).astype(np.float32)
for case in cases:
    kind, size = case.split('_')
    hw = 64 if size == 'small' else 224
    nb = 4 if size == 'small' else int(size) // 224 * 8 if size.isdigit() else 4
    if size == '224':
        nb = 8
    if size == 'big':
        hw, nb = 224, 64
    for prec in ('tf32', 'fp32'):
        if kind == 'f32':
            x = rng.random((nb, hw, hw, 3), np.float32)
            f, fn = vkjax.wrap(stem, precision=prec), stem
        else:
            x = rng.integers(0, 256, (nb, hw, hw, 3), dtype=np.uint8)
            f, fn = vkjax.wrap(stem_u8, precision=prec), stem_u8
        y = f(x, w)
        yt = oracle(fn, [x, w])
        yt = yt[0] if isinstance(yt, (tuple, list)) else yt
        err = float(np.abs(y - np.asarray(yt).reshape(y.shape)).max())
        interp = list(f._jaxpr_interpreters.values())[0]
        print(case, prec, 'ops', [type(o).__name__ + ':' + getattr(o, 'path', '') for o in interp.all_ops], 'rows', [o.attrs.get('rows') for o in interp.all_ops if hasattr(o, 'attrs')], 'max abs err', err, flush=True)
