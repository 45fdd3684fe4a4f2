from .datasets import *     # noqa: F401,F403
