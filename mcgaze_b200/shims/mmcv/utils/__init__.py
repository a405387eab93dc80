from mcgaze_b200.compat import Config, ConfigDict, DictAction, Registry, build_from_cfg  # noqa: F401
