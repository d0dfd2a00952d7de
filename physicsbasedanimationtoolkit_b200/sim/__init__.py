"""``pbat.sim`` -- only the ``vbd`` sub-module is in scope."""
from . import vbd  # noqa: F401
