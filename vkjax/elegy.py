"""Drop-in alias of vkjax_b200.elegy (≙ `from vkjax.elegy import vkModel`, reference README.md:31)."""
from vkjax_b200.elegy import vkModel  # noqa: F401
