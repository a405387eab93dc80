from mcgaze_b200.compat.runner import fuse_conv_bn  # noqa: F401

from . import utils  # noqa: F401
