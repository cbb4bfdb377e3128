"""≙ jax.nn pieces used by the reference's models (relu is a custom_jvp there, which is why the
reference carries a `custom_jvp_call_jaxpr` handler, reference vkjax/ops.py:246-260)."""
import numpy as np

from . import lax, jnp
from .tracing import custom_jvp


@custom_jvp
def relu(x):
    return lax.max(x, 0.0)


def softmax(x, axis=-1):
    m = jnp.max(x, axis=axis, keepdims=True)
    e = jnp.exp(jnp.subtract(x, lax.stop_gradient(m)))
    return jnp.true_divide(e, jnp.sum(e, axis=axis, keepdims=True))


def log_softmax(x, axis=-1):
    shifted = jnp.subtract(x, lax.stop_gradient(jnp.max(x, axis=axis, keepdims=True)))
    return jnp.subtract(shifted, jnp.log(jnp.sum(jnp.exp(shifted), axis=axis, keepdims=True)))
