"""Host-side mirror of ``brevitas.function`` (src/brevitas/function/__init__.py) on the B200 kernels."""
from .ops import *  # noqa: F401,F403
from .ops_ste import *  # noqa: F401,F403
from .shape import *  # noqa: F401,F403
