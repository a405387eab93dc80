"""Host-side stand-ins for the small part of mmcv-full 1.4.x that the reference's inference
entry points touch: python-dict configs with `_base_`
inheritance, the string->class registry, non-strict checkpoint loading with `revise_keys`,
and DataContainer/collate/scatter.  No compute lives here."""
from .config import Config, ConfigDict, DictAction  # noqa: F401
from .registry import Registry, build_from_cfg  # noqa: F401
from .checkpoint import load_checkpoint, load_state_dict  # noqa: F401
from .parallel import DataContainer, collate, scatter  # noqa: F401
