"""Drop-in alias: `import vkjax; vkjax.wrap(fun)` resolves to the B200-native implementation."""
from vkjax_b200 import Function, wrap, JaxprInterpreter, DeviceArray, device_put  # noqa: F401
