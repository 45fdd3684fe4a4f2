from dfno_b200.parallel.primitives import ZeroVolumeCorrectorFunction      # noqa: F401
