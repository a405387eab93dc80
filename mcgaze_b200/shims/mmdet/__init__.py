"""`mmdet` import surface of the reference's inference tools on the B200 backend (see mcgaze_b200.shims)."""
from . import apis, core, datasets, models, utils  # noqa: F401

__version__ = '2.25.0'
