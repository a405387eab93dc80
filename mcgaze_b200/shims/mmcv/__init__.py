"""`mmcv` import surface of the reference's inference tools, backed by mcgaze_b200.compat (mmcv-full 1.4.8 is not
installable offline).  Version string inside the range mmdet/__init__.py:19-20 accepts."""
from mcgaze_b200.compat import Config, ConfigDict, DictAction  # noqa: F401
from mcgaze_b200.compat.runner import dump, load, mkdir_or_exist  # noqa: F401

from . import cnn, parallel, runner, utils  # noqa: F401

__version__ = '1.4.8'
