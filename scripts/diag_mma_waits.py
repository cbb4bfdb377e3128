"""Developer diagnostic: run single conv layers of ResNet-50 b256 under a -DB2J_DIAG_MMA_WAITS build (B2J_LIB=...) so that the
MMA thread of CTA 0 prints where it waited.  python scripts/diag_mma_waits.py [fp32|tf32]"""
import sys
import numpy as np
sys.path.insert(0, '.')
import vkjax_b200 as vkjax
from vkjax_b200.frontend import lax
from vkjax_b200.core import ConvDimensionNumbers
dn = ConvDimensionNumbers((0, 3, 1, 2), (3, 2, 0, 1), (0, 3, 1, 2))
prec = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
rng = np.random.default_rng(0)
cases = [('stage0 3x3 C64->64', (256, 56, 56, 64), (3, 3, 64, 64)), ('stage1 3x3 C128->128', (256, 28, 28, 128), (3, 3, 128, 128)),
         ('stage2 3x3 C256->256', (256, 14, 14, 256), (3, 3, 256, 256)), ('stage2 1x1 C1024->256', (256, 14, 14, 1024), (1, 1, 1024, 256)),
         ('stage2 1x1 C256->1024', (256, 14, 14, 256), (1, 1, 256, 1024))]
for name, xs, ws in cases:
    x = rng.random(xs, np.float32); w = rng.normal(0, 0.05, ws).astype(np.float32)
    f = vkjax.wrap(lambda x, w: lax.conv_general_dilated(x, w, (1, 1), 'SAME', dimension_numbers=dn), precision=prec)
    print('==', name, prec, flush=True)
    f(x, w)
    f(x, w)
