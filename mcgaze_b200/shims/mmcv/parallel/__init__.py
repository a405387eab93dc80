from mcgaze_b200.compat.parallel import DataContainer, collate, scatter  # noqa: F401
from mcgaze_b200.compat.runner import MMDataParallel, MMDistributedDataParallel  # noqa: F401
