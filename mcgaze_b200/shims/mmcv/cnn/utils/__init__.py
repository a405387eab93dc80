from . import flops_counter  # noqa: F401
