"""Host logic (no GPU): the JAX-free tracer emits the dialect the handlers expect, and the fusion pass produces
the op lists the design promises."""
import numpy as np
import pytest

from vkjax_b200 import JaxprInterpreter, nets, ops, fusion, runtime as rt
from vkjax_b200.frontend import make_jaxpr, jit, lax, jnp, nn
from vkjax_b200.ops import ChainOp, ContractionOp, KernelOp


def names(jaxpr):
    return [e.primitive.name for e in jaxpr.jaxpr.eqns]


def test_readme_example_jaxpr():
    j, shapes = make_jaxpr(lambda x, W, b: jnp.dot(x, W) + b, return_shape=True)(np.zeros((8, 128)), np.zeros((128, 16)), np.zeros(16))
    assert names(j) == ['dot_general', 'broadcast_in_dim', 'add']
    assert shapes.shape == (8, 16) and shapes.dtype == np.float32          # float64 inputs are traced as float32
    assert j.jaxpr.eqns[0].params['dimension_numbers'] == (((1,), (0,)), ((), ()))


def test_nested_jit_and_relu_dialect():
    def add4(x, y):
        return jit(lambda a: a + a)(x) + jit(lambda a, b: a + b)(y, 1.0)
    j = make_jaxpr(add4)(5.0, 7.1)
    assert names(j) == ['xla_call', 'xla_call', 'add']
    assert j.jaxpr.eqns[0].params['device'] is None and j.jaxpr.eqns[0].params['backend'] is None
    j = make_jaxpr(nn.relu)(np.zeros(3, np.float32))
    assert names(j) == ['custom_jvp_call_jaxpr'] and j.jaxpr.eqns[0].params['num_consts'] == 0


def test_everything_is_recorded_nothing_evaluated():
    j = make_jaxpr(lambda: lax.add(5, 100))()
    assert names(j) == ['add'] and len(j.jaxpr.invars) == 0
    with pytest.raises(RuntimeError):
        lax.add(np.float32(1), np.float32(2))           # outside a trace there is no host evaluation path


def test_static_argnums():
    f = lambda x, training: x * 2.0 if training else x + 1.0
    assert names(make_jaxpr(f, static_argnums=[1])(np.zeros(3, np.float32), True)) == ['mul']
    assert names(make_jaxpr(f, static_argnums=[1])(np.zeros(3, np.float32), False)) == ['add']


def test_resnet50_fusion_shape():
    model = nets.ResNet50()
    s = model.init(0)
    j = make_jaxpr(lambda x, s: model.apply(s, x))(np.zeros((4, 224, 224, 3), np.float32), s)
    it = JaxprInterpreter(j, dry_run=True, precision='tf32')
    assert it.unfused_ops == len([e for e in j.jaxpr.eqns]) - 0 or it.unfused_ops > 400
    convs = [o for o in it.all_ops if isinstance(o, ContractionOp)]
    assert len(convs) == 54 and all(o.path == 'tc' for o in convs)
    assert sum(1 for o in convs if len(o.epilogue) == 5) == 16          # BN + residual + ReLU
    assert sum(1 for o in convs if len(o.epilogue) == 4) == 33          # BN + ReLU
    assert sum(1 for o in convs if len(o.epilogue) == 3) == 4           # projection shortcut: BN only
    assert len(it.all_ops) == 110
    flops, rows = model.conv_flops(256)
    assert abs(flops / 1e9 - 2093.66) < 0.1                              # 2092.61 conv + 1.05 FC


def test_operand_classification():
    assert fusion.classify_operand((1, 1, 1, 64), (8, 7, 7, 64))[:2] == (rt.OPK_MOD, 64)
    assert fusion.classify_operand((8, 1), (8, 10))[:2] == (rt.OPK_DIV, 10)
    assert fusion.classify_operand((), (3, 3))[0] == rt.OPK_SCALAR
    assert fusion.classify_operand((4, 4), (4, 4))[0] == rt.OPK_FULL
    kind, _, strides = fusion.classify_operand((2, 1, 5), (2, 3, 5))
    assert kind == rt.OPK_STRIDED and strides == [5, 0, 1]


def test_broadcast_elision_and_keep_outputs():
    def f(x, b):
        return x + lax.broadcast_in_dim(b, x.shape, (1,))
    j = make_jaxpr(f)(np.zeros((4, 8), np.float32), np.zeros((8,), np.float32))
    assert [type(o).__name__ for o in JaxprInterpreter(j, dry_run=True, fuse=False).all_ops] == ['KernelOp', 'ChainOp']
    assert [type(o).__name__ for o in JaxprInterpreter(j, dry_run=True, fuse=True).all_ops] == ['ChainOp']
    # an intermediate that is also an output must stay materialised
    def g(x):
        y = x * 2.0
        return y, y + 1.0
    j = make_jaxpr(g)(np.zeros((4,), np.float32))
    assert len(JaxprInterpreter(j, dry_run=True, fuse=True).all_ops) == 2


def test_error_conventions_host_side():
    from vkjax_b200 import core
    from vkjax_b200.frontend import tracing
    def bad(x):
        return tracing.bind(core.Primitive('sort'), x, out_avals=[core.ShapedArray(x.shape, x.dtype)])
    with pytest.raises(NotImplementedError):
        JaxprInterpreter(make_jaxpr(bad)(np.zeros(4, np.float32)), dry_run=True)
    with pytest.raises(NotImplementedError):
        make_jaxpr(lambda x: x + 1)(np.zeros(3, np.float16))            # unsupported dtype (reference ops.py:19-25)
    with pytest.raises(ValueError):
        JaxprInterpreter(make_jaxpr(lambda x: x)(np.zeros(3, np.float32)), dry_run=True, precision='bf16')


def test_resident_inputs_hoist_weight_only_work():
    """Ops fed only by device-resident inputs (BN parameter folding) and the filter re-layout of tensor-core convs go
    to the prologue; nothing is hoisted when the weights arrive as host arrays."""
    model = nets.ResNet50()
    s = model.init(0)
    from vkjax_b200 import tree_util
    x = np.zeros((4, 224, 224, 3), np.float32)
    j = make_jaxpr(lambda x, s: model.apply(s, x))(x, s)
    n_leaves = len(tree_util.tree_leaves((x, s)))
    resident = (False,) + (True,) * (n_leaves - 1)
    it = JaxprInterpreter(j, dry_run=True, precision='tf32', resident_inputs=resident)
    chains = [o for o in it.all_ops if isinstance(o, ChainOp) and getattr(o, 'hoisted', False)]
    preps = [o for o in it.all_ops if isinstance(o, ContractionOp) and getattr(o, 'prep_hoisted', False)]
    assert len(chains) == 53 and len(preps) == 54 and it.n_hoisted == 107
    assert all(o.temps[0].is_constant() for o in preps)                    # prepared weights persist across calls
    assert not any(getattr(o, 'hoisted', False) for o in it.all_ops if isinstance(o, (ContractionOp, KernelOp)))
    it0 = JaxprInterpreter(j, dry_run=True, precision='tf32')
    assert it0.n_hoisted == 0
    # the image is an input of every conv: nothing that depends on it may be hoisted
    it1 = JaxprInterpreter(j, dry_run=True, precision='tf32', resident_inputs=(True,) + (False,) * (n_leaves - 1))
    assert it1.n_hoisted == 0
