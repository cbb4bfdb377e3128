"""`vkjax_b200.elegy.vkModel` (≙ reference vkjax/elegy.py:6-35).

The reference subclasses `elegy.Model` and overrides `jit_step()` so that Elegy's five jitted step
functions go through `vkjax.wrap` with the same `static_argnums`.  When Elegy is importable that is
exactly what happens here.  Elegy 0.7.1 cannot be installed in this image, so otherwise `vkModel`
derives from `_Model`, a minimal stand-in that keeps the members the predict path uses
(`states`, `initialized`, `jit_step`, `call_pred_step`, `call_pred_step_jit`, `predict`,
`predict_on_batch`, `jitted_members`) for modules with `init(rng)` / `apply(states, x)` (vkjax_b200/nets.py).
"""
import numpy as np

from . import function
from .interpreter import device_put

try:                                    # pragma: no cover - Elegy is not available in this image
    import elegy as _elegy
    _Base = _elegy.Model
    HAVE_ELEGY = True
except ImportError:
    HAVE_ELEGY = False

    class _Model:
        """The slice of elegy.Model the reference's predict path touches."""
        def __init__(self, module, seed=42, **wrap_kwargs):
            self.module = module
            self.seed = seed
            self.states = None
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

        def call_init_step(self, x):
            raise NotImplementedError('weights are initialised on the host: model.init(x)')

        def call_test_step(self, *a):
            raise NotImplementedError('test/train steps need Elegy (SURVEY.md §8f rank 2)')

        call_train_step = call_test_step

        def jit_step(self):
            pass

        # -- user API ---------------------------------------------------------------------------------
        def init(self, x=None, seed=None):
            self.states = device_put(self.module.init(np.random.default_rng(self.seed if seed is None else seed)))
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
