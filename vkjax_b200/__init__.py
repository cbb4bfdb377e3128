"""vkjax_b200 -- a B200-native jaxpr executor behind the vkJAX API (`wrap`, `Function`,
`vkjax_b200.elegy.vkModel`).  See DESIGN.md."""
from .function import Function, wrap
from .interpreter import JaxprInterpreter, DeviceArray, device_put
from . import frontend

__all__ = ['Function', 'wrap', 'JaxprInterpreter', 'DeviceArray', 'device_put', 'frontend']
