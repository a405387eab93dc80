from mcgaze_b200.compat.parallel import DataContainer, collate, scatter  # noqa: F401
