from . import backend                                      # noqa: F401
