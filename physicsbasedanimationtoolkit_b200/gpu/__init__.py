"""``pbat.gpu`` -- only the ``vbd`` sub-module is in scope."""
from . import vbd  # noqa: F401

from . import geometry  # noqa: F401,E402
