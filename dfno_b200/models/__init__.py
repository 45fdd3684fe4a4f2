from .linear import *             # noqa: F401,F403
from .norm import *               # noqa: F401,F403
from .fno import *                # noqa: F401,F403
from .loss import *               # noqa: F401,F403
