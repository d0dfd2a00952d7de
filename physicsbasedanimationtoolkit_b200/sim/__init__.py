"""``pbat.sim`` -- the ``vbd`` sub-module (the hot path) and ``xpbd`` (section 8f rank 4)."""
from . import vbd, xpbd  # noqa: F401
