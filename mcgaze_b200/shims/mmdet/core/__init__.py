from . import bbox  # noqa: F401
