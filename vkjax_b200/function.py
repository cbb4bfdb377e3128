"""`vkjax.wrap(fun)` / `vkjax.Function` (≙ reference vkjax/function.py:7-47).

Traces `fun` to a jaxpr on first call with a new input signature, builds a JaxprInterpreter for
it and caches it; later calls replay the recorded CUDA graph.  The tracer is `jax.make_jaxpr`
when JAX is importable, otherwise the built-in JAX-free front end (vkjax_b200/frontend), selected
with `frontend='auto'|'jax'|'builtin'`.
"""
import typing as tp

import numpy as np

from . import tree_util
from .interpreter import JaxprInterpreter, DeviceArray, leaf_shape_dtype


def _pick_frontend(frontend: str):
    if frontend in ('auto', 'jax'):
        try:
            import jax                                    # noqa: F401
            return 'jax'
        except ImportError:
            if frontend == 'jax':
                raise
    return 'builtin'


class Function:
    def __init__(self, function: tp.Callable, static_argnums: tp.Tuple[int] = (), profiling: bool = False,
                 *args, frontend: str = 'auto', **kwargs):
        self.frontend = _pick_frontend(frontend)
        if self.frontend == 'jax':
            import jax
            self.jaxpr_function = jax.make_jaxpr(function, static_argnums, return_shape=True)
            self._tree = jax.tree_util
        else:
            from .frontend import make_jaxpr
            self.jaxpr_function = make_jaxpr(function, static_argnums, return_shape=True)
            self._tree = tree_util
        self._static_argnums = (static_argnums,) if isinstance(static_argnums, int) else tuple(static_argnums)
        self._jaxpr_interpreters = dict()
        self._output_shapes = dict()
        self._profiling = profiling
        kwargs['profiling'] = profiling
        kwargs['static_argnums'] = self._static_argnums
        self._interpreter_args = (args, kwargs)

    def __call__(self, *args: tp.Any, **kwargs: tp.Any) -> tp.Any:
        jaxpr_interpreter, output_shapes, leaves = self._get_or_create_jaxpr_interpreter(args)
        output = jaxpr_interpreter.run_leaves(leaves, **kwargs)
        output = self._restore_shapes(output, output_shapes, getattr(jaxpr_interpreter, 'gather_buffers', None) is not None)
        if not self._profiling:
            return output
        return output, jaxpr_interpreter.get_profiling_info()

    def map(self, arg_batches: tp.Sequence[tp.Tuple], lanes: int = 2) -> tp.List[tp.Any]:
        """[self(*args) for args in arg_batches] for same-signature argument tuples, with the host->device copy of the
        next batch overlapping the replay of the current one (JaxprInterpreter.run_many)."""
        arg_batches = list(arg_batches)
        if not arg_batches:
            return []
        memo = {}                      # arguments shared between batches (the model state) are flattened once
        interp, output_shapes, leaves0 = self._get_or_create_jaxpr_interpreter(arg_batches[0], memo)
        all_leaves = [leaves0]
        for a in arg_batches[1:]:
            it, _, leaves = self._get_or_create_jaxpr_interpreter(a, memo)
            if it is not interp:
                raise TypeError('Function.map needs argument tuples of one signature')
            all_leaves.append(leaves)
        gathered = getattr(interp, 'gather_buffers', None) is not None
        return [self._restore_shapes(o, output_shapes, gathered) for o in interp.run_many_leaves(all_leaves, lanes=lanes)]

    def _flatten_arg(self, a, memo):
        """(leaves, (treedef, shapes, dtypes, resident flags)) of one top-level argument."""
        hit = memo.get(id(a)) if memo is not None else None
        if hit is not None and hit[0] is a:
            return hit[1], hit[2]
        leaves, structure = self._tree.tree_flatten(a)
        sd = [leaf_shape_dtype(x) for x in leaves]
        sig = (structure, tuple(s for s, _ in sd), tuple(d for _, d in sd), tuple(isinstance(x, DeviceArray) for x in leaves))
        if memo is not None and not isinstance(a, np.ndarray):
            memo[id(a)] = (a, leaves, sig)
        return leaves, sig

    def _get_or_create_jaxpr_interpreter(self, args: tp.Tuple[tp.Any], memo=None):
        leaves, sigs = [], []
        for i, a in enumerate(args):
            if i in self._static_argnums:
                # static argument *values* are part of the key: the reference keys on their shape/dtype only,
                # so e.g. training=True/False would share one trace (quirk Q1, reference function.py:27-30)
                sigs.append(('static', _hashable(a)))
                continue
            l, sig = self._flatten_arg(a, memo)
            leaves += l
            sigs.append(sig)
        shape_structure = tuple(sigs)
        if shape_structure not in self._jaxpr_interpreters:
            # new input shapes or structure, need to re-trace
            resident = tuple(isinstance(x, DeviceArray) for x in leaves)
            trace_args = self._tree.tree_map(_abstract_leaf, args) if any(resident) else args
            jaxpr, output_shapes = self.jaxpr_function(*trace_args)
            iargs, ikwargs = self._interpreter_args
            self._jaxpr_interpreters[shape_structure] = JaxprInterpreter(jaxpr, *iargs, resident_inputs=resident, **ikwargs)
            # (pytree of ShapeDtypeStruct, its structure, its leaves) -- flattened once, used on every call
            self._output_shapes[shape_structure] = (output_shapes, self._tree.tree_structure(output_shapes),
                                                    self._tree.tree_leaves(output_shapes))
        return self._jaxpr_interpreters[shape_structure], self._output_shapes[shape_structure], leaves

    def _restore_shapes(self, x, targetshapes, gathered=False):
        _, structure, flat_shapes = targetshapes
        def restore(a, s):
            if isinstance(a, DeviceArray):
                return a
            shape = tuple(s.shape)
            a = np.asarray(a)
            if gathered and a.size != int(np.prod(shape, dtype=np.int64)):
                # outputs of all ranks were all-gathered along the leading (batch) dimension
                shape = (-1,) + shape[1:] if len(shape) else (-1,)
            return a.reshape(shape)
        x = [restore(a, s) for a, s in zip(x, flat_shapes)]
        return self._tree.tree_unflatten(structure, x)


def _abstract_leaf(x):
    if isinstance(x, DeviceArray):
        return np.zeros(x.shape, x.dtype)       # tracing needs shape/dtype only
    return x


def _hashable(a):
    try:
        hash(a)
        return a
    except TypeError:
        return repr(a)


def wrap(function: tp.Callable, *args, **kwargs):
    return Function(function, *args, **kwargs)
