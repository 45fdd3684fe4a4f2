from .decomposition import *      # noqa: F401,F403
from .partition import *          # noqa: F401,F403
from .planner import *            # noqa: F401,F403
from .primitives import *         # noqa: F401,F403
