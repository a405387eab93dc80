from mcgaze_b200.compat.checkpoint import load_checkpoint, load_state_dict  # noqa: F401
