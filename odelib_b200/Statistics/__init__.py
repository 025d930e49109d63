from . import Samplers, stats  # noqa: F401
