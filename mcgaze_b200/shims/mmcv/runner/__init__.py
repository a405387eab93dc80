from mcgaze_b200.compat.checkpoint import load_checkpoint, load_state_dict  # noqa: F401
from mcgaze_b200.compat.runner import get_dist_info, init_dist, wrap_fp16_model  # noqa: F401
