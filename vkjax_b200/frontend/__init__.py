"""JAX-free front end: produces jaxprs in the dialect `vkjax_b200.ops` consumes.

`jax` is not installable in this image.  When it *is* importable, `vkjax.wrap`
uses `jax.make_jaxpr` instead (vkjax_b200/function.py) and this package is unused.
"""
from . import tracing, lax, jnp, random, nn, autodiff
from .tracing import make_jaxpr, jit, custom_jvp, Tracer
from .autodiff import grad, value_and_grad, vjp
