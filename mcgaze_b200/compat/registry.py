"""mmcv.utils.Registry work-alike: the `type='...'` string -> class plugin API the reference's
configs use (mmdet/models/builder.py:1-59, mmdet/models/utils/builder.py:5-10)."""
from __future__ import annotations

import inspect
from typing import Any, Callable, Dict, Optional


def build_from_cfg(cfg: Dict[str, Any], registry: 'Registry', default_args: Optional[Dict[str, Any]] = None):
    if not isinstance(cfg, dict):
        raise TypeError(f'cfg must be a dict, but got {type(cfg)}')
    if 'type' not in cfg and not (default_args and 'type' in default_args):
        raise KeyError(f'`cfg` or `default_args` must contain the key "type", but got {cfg}\n{default_args}')
    args = dict(cfg)
    if default_args is not None:
        for k, v in default_args.items():
            args.setdefault(k, v)
    obj_type = args.pop('type')
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError(f'{obj_type} is not in the {registry.name} registry')
    elif inspect.isclass(obj_type) or inspect.isfunction(obj_type):
        obj_cls = obj_type
    else:
        raise TypeError(f'type must be a str or valid type, but got {type(obj_type)}')
    try:
        return obj_cls(**args)
    except Exception as e:
        raise type(e)(f'{obj_cls.__name__}: {e}') from e


class Registry:
    def __init__(self, name: str, build_func: Optional[Callable] = None, parent: Optional['Registry'] = None,
                 scope: Optional[str] = None):
        self._name = name
        self._module_dict: Dict[str, Any] = {}
        self._children: Dict[str, 'Registry'] = {}
        self.parent = parent
        self.scope = scope
        if build_func is None:
            build_func = parent.build_func if parent is not None else build_from_cfg
        self.build_func = build_func
        if parent is not None:
            parent._children[name] = self

    @property
    def name(self) -> str:
        return self._name

    @property
    def module_dict(self) -> Dict[str, Any]:
        return self._module_dict

    def __len__(self):
        return len(self._module_dict)

    def __contains__(self, key):
        return self.get(key) is not None

    def __repr__(self):
        return f'Registry(name={self._name}, items={sorted(self._module_dict)})'

    def get(self, key: str):
        if key in self._module_dict:
            return self._module_dict[key]
        if '.' in key:                      # "scope.Type"
            key = key.split('.', 1)[1]
            if key in self._module_dict:
                return self._module_dict[key]
        if self.parent is not None:
            return self.parent.get(key)
        return None

    def build(self, *args, **kwargs):
        return self.build_func(*args, **kwargs, registry=self)

    def _register_module(self, module_class, module_name=None, force=False):
        if module_name is None:
            module_name = module_class.__name__
        names = [module_name] if isinstance(module_name, str) else module_name
        for n in names:
            if not force and n in self._module_dict:
                raise KeyError(f'{n} is already registered in {self.name}')
            self._module_dict[n] = module_class

    def register_module(self, name=None, force: bool = False, module=None):
        if not isinstance(force, bool):
            raise TypeError(f'force must be a boolean, but got {type(force)}')
        if module is not None:
            self._register_module(module, name, force)
            return module

        def _register(cls):
            self._register_module(cls, name, force)
            return cls

        return _register
