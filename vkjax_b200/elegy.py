"""`vkjax_b200.elegy.vkModel` (≙ reference vkjax/elegy.py:6-35).

The reference subclasses `elegy.Model` and overrides `jit_step()` so that Elegy's five jitted step functions go through
`vkjax.wrap` with the same `static_argnums` (pred / summary [2, 3], test / train [5, 6], init []).  When Elegy is
importable that is exactly what happens here.  Elegy 0.7.1 cannot be installed in this image, so otherwise `vkModel`
derives from `_Model`, a stand-in for the slice of `elegy.Model` the reference's tests drive
(tests/test_elegy_mlp.py:37-148, tests/test_elegy_conv.py:27-94, tests/test_elegy_resnet.py:18-32):

    call_pred_step(x, states, initializing, training)                                   -> (y_pred, states)
    call_test_step(x, y_true, sample_weight, class_weight, states, initializing, training)  -> (loss, logs, states)
    call_train_step(x, y_true, sample_weight, class_weight, states, initializing, training) -> (logs, states)
    call_init_step(x, rng)                                                              -> states
    init / predict / predict_on_batch / test_on_batch / train_on_batch / evaluate / fit

for modules of vkjax_b200.nets (`init_traced(key, x)` / `init(rng)` / `apply(params, x)`).  All five step functions are
traced by the JAX-free front end (gradients by vkjax_b200.frontend.autodiff, random initialisers by
vkjax_b200.frontend.random) and run on the GPU; `losses` / `optimizers` below are the minimal counterparts of
`elegy.losses.SparseCategoricalCrossentropy` and `optax.sgd` those tests use.

`states` of the stand-in: the parameter pytree for prediction (as in round 1), `TrainStates(params, opt)` for the
train step (the optimizer state travels with the parameters, as Elegy's `States.optimizer_states` does).
"""
import typing as tp

import numpy as np

from . import function, tree_util
from .interpreter import device_put, DeviceArray


# =================================================================================================
class losses:
    class SparseCategoricalCrossentropy:
        """≙ elegy.losses.SparseCategoricalCrossentropy(from_logits=True): mean over the batch of -log softmax(logits)[label]
        (the function reference tests/test_elegy_mlp.py:62-83 differentiates)."""
        def __init__(self, from_logits=True):
            assert from_logits, 'only from_logits=True is used by the reference tests'

        def __call__(self, y_true, y_pred):
            from .frontend import jnp, nn
            logp = nn.log_softmax(y_pred, axis=-1)
            picked = jnp.take_along_axis(logp, y_true.reshape(-1, 1), axis=-1)
            return -jnp.mean(picked)

    class MeanSquaredError:
        def __call__(self, y_true, y_pred):
            from .frontend import jnp
            d = y_pred - y_true
            return jnp.mean(d * d)


class optimizers:
    class sgd:
        """≙ optax.sgd(learning_rate, momentum=None): init(params) -> state, update(grads, state, params) -> (params, state)"""
        def __init__(self, learning_rate, momentum=None):
            self.lr, self.momentum = float(learning_rate), momentum

        def init(self, params):
            if self.momentum is None:
                return ()
            return tree_util.tree_map(lambda p: np.zeros(np.shape(p), np.float32), params)

        def update(self, grads, state, params):
            lr = np.float32(self.lr)
            if self.momentum is None:
                return tree_util.tree_map(lambda p, g: p - lr * g, params, grads), ()
            m = np.float32(self.momentum)
            new_state = tree_util.tree_map(lambda t, g: m * t + g, state, grads)
            return tree_util.tree_map(lambda p, t: p - lr * t, params, new_state), new_state


class TrainStates(tp.NamedTuple):
    params: tp.Any
    opt: tp.Any


try:                                    # pragma: no cover - Elegy is not available in this image
    import elegy as _elegy
    _Base = _elegy.Model
    HAVE_ELEGY = True
except ImportError:
    HAVE_ELEGY = False

    class _Model:
        """The slice of elegy.Model the reference's tests touch (see the module docstring)."""
        def __init__(self, module, loss=None, optimizer=None, seed=42, **wrap_kwargs):
            self.module = module
            self.loss = loss
            self.optimizer = optimizer
            self.seed = seed
            self.states = None
            self.optimizer_states = None
            self.initialized = False
            self.jitted_members = set()
            self._wrap_kwargs = wrap_kwargs
            self.jit_step()

        # -- step functions (elegy.Model.call_*_step) ---------------------------------------------
        def call_pred_step(self, x, states, initializing, training):
            y_pred = self.module.apply(states, x)
            return y_pred, states

        def call_summary_step(self, x, states, initializing, training):
            return self.call_pred_step(x, states, initializing, training)

        def call_init_step(self, x, rng):
            """Initial parameters from the module's random initialisers, traced (threefry2x32 -> uniform -> erf_inv chains of
            vkjax_b200.frontend.random) and run on the device (≙ elegy.Model.call_init_step, reference tests/test_elegy_mlp.py:134-148)."""
            return self.module.init_traced(rng, x)

        def _loss_and_logs(self, params, x, y_true, sample_weight, class_weight):
            if self.loss is None:
                raise ValueError('vkModel(module, loss=...) is required for test / train steps')
            if sample_weight is not None or class_weight is not None:
                raise NotImplementedError('sample_weight / class_weight')
            y_pred = self.module.apply(params, x)
            loss = self.loss(y_true, y_pred)
            return loss, {'loss': loss, 'y_pred': y_pred}

        def call_test_step(self, x, y_true, sample_weight, class_weight, states, initializing, training):
            params = states.params if isinstance(states, TrainStates) else states
            loss, logs = self._loss_and_logs(params, x, y_true, sample_weight, class_weight)
            return loss, logs, states

        def call_train_step(self, x, y_true, sample_weight, class_weight, states, initializing, training):
            """One optimisation step: jax.value_and_grad(loss)(params) -> optimizer.update (≙ elegy.Model.call_train_step,
            reference tests/test_elegy_mlp.py:87-118)."""
            from .frontend import value_and_grad
            if self.optimizer is None:
                raise ValueError('vkModel(module, loss=..., optimizer=...) is required for the train step')
            (loss, logs), grads = value_and_grad(lambda p: self._loss_and_logs(p, x, y_true, sample_weight, class_weight),
                                                 has_aux=True)(states.params)
            new_params, new_opt = self.optimizer.update(grads, states.opt, states.params)
            return logs, TrainStates(new_params, new_opt)

        def jit_step(self):
            pass

        # -- user API ---------------------------------------------------------------------------------
        def init(self, x=None, y=None, seed=None, host=None):
            """Initialises `states` (device resident).  Modules with `init_traced` are initialised ON THE DEVICE through
            call_init_step_jit (needs `x` for the input shape); `host=True` (or a module without `init_traced`) draws the
            synthetic numpy initialisation of vkjax_b200.nets instead."""
            seed = self.seed if seed is None else seed
            traced = hasattr(self.module, 'init_traced') and x is not None if host is None else not host
            if traced:
                from .frontend import random
                params = self.call_init_step_jit(np.asarray(x) if not hasattr(x, 'shape') else x, random.PRNGKey(seed))
            else:
                params = self.module.init(np.random.default_rng(seed))
            self.states = device_put(params)
            if self.optimizer is not None:
                self.optimizer_states = device_put(self.optimizer.init(tree_util.tree_map(np.asarray, params)))
            self.initialized = True

        def predict_on_batch(self, x):
            if not self.initialized:
                self.init(x)
            y_pred, _ = self.call_pred_step_jit(x, self.states, False, False)
            return y_pred

        def predict(self, x, batch_size=None, initialize=False):
            if initialize and not self.initialized:
                self.init(x)
            x = np.asarray(x) if not hasattr(x, 'shape') else x
            if batch_size is None or x.shape[0] <= batch_size:
                return self.predict_on_batch(x)
            if not self.initialized:
                self.init(x)
            # full batches are pipelined: the upload of batch i+1 overlaps the replay of batch i (Function.map)
            n_full = x.shape[0] // batch_size
            outs = self.call_pred_step_jit.map([(x[i * batch_size:(i + 1) * batch_size], self.states, False, False)
                                                for i in range(n_full)])
            ys = [y for y, _ in outs]
            if n_full * batch_size < x.shape[0]:
                ys.append(self.predict_on_batch(x[n_full * batch_size:]))
            return np.concatenate(ys)

        def test_on_batch(self, x, y):
            if not self.initialized:
                self.init(x, y)
            loss, logs, _ = self.call_test_step_jit(x, y, None, None, self.states, False, False)
            return {'loss': logs['loss']}

        evaluate = test_on_batch

        def train_on_batch(self, x, y):
            """≙ elegy.Model.train_on_batch: one step; the new parameters / optimizer state replace `states`."""
            if not self.initialized:
                self.init(x, y)
            if self.optimizer_states is None and self.optimizer is not None:
                self.optimizer_states = device_put(self.optimizer.init(tree_util.tree_map(np.asarray, self.states)))
            logs, new = self.call_train_step_jit(x, y, None, None, TrainStates(self.states, self.optimizer_states), False, True)
            self.states, self.optimizer_states = device_put(new.params), device_put(new.opt)
            return {'loss': logs['loss']}

        def fit(self, x, y, epochs=1, batch_size=32):
            history = []
            n = x.shape[0] // batch_size
            for _ in range(epochs):
                for i in range(n):
                    history.append(float(self.train_on_batch(x[i * batch_size:(i + 1) * batch_size],
                                                             y[i * batch_size:(i + 1) * batch_size])['loss']))
            return history

    _Base = _Model


class vkModel(_Base):
    def jit_step(self):
        kw = getattr(self, '_wrap_kwargs', {})
        self.call_summary_step_jit = function.wrap(self.call_summary_step, static_argnums=[2, 3], **kw)
        self.call_pred_step_jit = function.wrap(self.call_pred_step, static_argnums=[2, 3], **kw)
        self.call_test_step_jit = function.wrap(self.call_test_step, static_argnums=[5, 6], **kw)
        self.call_train_step_jit = function.wrap(self.call_train_step, static_argnums=[5, 6], **kw)
        self.call_init_step_jit = function.wrap(self.call_init_step, static_argnums=[], **kw)
        self.jitted_members |= {
            'call_summary_step_jit',
            'call_pred_step_jit',
            'call_test_step_jit',
            'call_train_step_jit',
            'call_init_step_jit',
        }
