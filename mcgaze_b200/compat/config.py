"""mmcv.Config work-alike (python-dict config files).

Behaviour restated from mmcv 1.4.x `mmcv/utils/config.py` (not under /root/reference):
  * `Config.fromfile(path)` executes the .py file and collects its public, non-module names;
  * `_base_` (str or list of str, relative to the file) is loaded first and deep-merged,
    child keys win, dicts merge recursively, lists are replaced wholesale;
  * a dict carrying `_delete_=True` replaces the base value instead of merging into it;
  * attribute access on nested dicts (`cfg.model.backbone.depth`);
  * `merge_from_dict({'a.b.c': v})` for `--cfg-options` (tools/test_gaze360_gaze.py:33-42,
    mmdet/apis/inference.py:36-37).
"""
from __future__ import annotations

import argparse
import ast
import copy
import os
import types
from typing import Any, Dict

BASE_KEY = '_base_'
DELETE_KEY = '_delete_'


class ConfigDict(dict):
    """dict with attribute access (addict.Dict semantics as far as mmcv uses them)."""

    def __getattr__(self, name: str) -> Any:
        try:
            return self[name]
        except KeyError:
            raise AttributeError(f"'ConfigDict' object has no attribute '{name}'") from None

    def __setattr__(self, name: str, value: Any) -> None:
        self[name] = _wrap(value)

    def __delattr__(self, name: str) -> None:
        try:
            del self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __deepcopy__(self, memo):
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def copy(self):
        return ConfigDict(super().copy())

    def to_dict(self) -> dict:
        return _unwrap(self)


def _wrap(v: Any) -> Any:
    if isinstance(v, ConfigDict):
        return v
    if isinstance(v, dict):
        return ConfigDict({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, list):
        return [_wrap(x) for x in v]
    if isinstance(v, tuple):
        return tuple(_wrap(x) for x in v)
    return v


def _unwrap(v: Any) -> Any:
    if isinstance(v, dict):
        return {k: _unwrap(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_unwrap(x) for x in v]
    if isinstance(v, tuple):
        return tuple(_unwrap(x) for x in v)
    return v


def _merge_a_into_b(a: Dict[str, Any], b: Dict[str, Any]) -> Dict[str, Any]:
    b = dict(b)
    for k, v in a.items():
        if isinstance(v, dict):
            if k in b and isinstance(b[k], dict) and not v.get(DELETE_KEY, False):
                b[k] = _merge_a_into_b(v, b[k])
            else:
                v = dict(v)
                v.pop(DELETE_KEY, None)
                b[k] = _merge_a_into_b(v, {})
        else:
            b[k] = v
    return b


def _file2dict(filename: str) -> Dict[str, Any]:
    filename = os.path.abspath(os.path.expanduser(filename))
    if not os.path.isfile(filename):
        raise FileNotFoundError(f'config file {filename} does not exist')
    if not filename.endswith('.py'):
        raise IOError('Only .py configs are supported')
    with open(filename, 'r', encoding='utf-8') as f:
        src = f.read()
    ast.parse(src, filename)           # syntax errors surface with the file name
    scope: Dict[str, Any] = {'__file__': filename}
    exec(compile(src, filename, 'exec'), scope)  # noqa: S102 - config files are python by design
    cfg = {k: v for k, v in scope.items()
           if not k.startswith('__') and not isinstance(v, (types.ModuleType, types.FunctionType, type))}
    if BASE_KEY in cfg:
        base = cfg.pop(BASE_KEY)
        base = base if isinstance(base, list) else [base]
        merged: Dict[str, Any] = {}
        for b in base:
            bd = _file2dict(os.path.join(os.path.dirname(filename), b))
            dup = merged.keys() & bd.keys()
            if dup:
                raise KeyError(f'Duplicate key is not allowed among bases: {sorted(dup)}')
            merged.update(bd)
        cfg = _merge_a_into_b(cfg, merged)
    return cfg


class Config:
    def __init__(self, cfg_dict: Dict[str, Any] | None = None, filename: str | None = None):
        object.__setattr__(self, '_cfg_dict', _wrap(cfg_dict or {}))
        object.__setattr__(self, '_filename', filename)

    @staticmethod
    def fromfile(filename: str) -> 'Config':
        return Config(_file2dict(filename), filename=filename)

    @property
    def filename(self):
        return self._filename

    def merge_from_dict(self, options: Dict[str, Any]) -> None:
        nested: Dict[str, Any] = {}
        for full_key, v in options.items():
            d = nested
            keys = full_key.split('.')
            for sub in keys[:-1]:
                d = d.setdefault(sub, {})
            d[keys[-1]] = v
        object.__setattr__(self, '_cfg_dict', _wrap(_merge_a_into_b(nested, _unwrap(self._cfg_dict))))

    def get(self, key, default=None):
        return self._cfg_dict.get(key, default)

    def __getattr__(self, name):
        if name.startswith('__') or name in ('_cfg_dict', '_filename'):     # copy / pickle probe blank instances
            raise AttributeError(name)
        return getattr(self._cfg_dict, name)

    def __setattr__(self, name, value):
        self._cfg_dict[name] = _wrap(value)

    def __deepcopy__(self, memo):
        """tools/analysis_tools/benchmark.py:148 deep-copies the config per repetition (mmcv.Config supports it)."""
        import copy
        return Config(copy.deepcopy(_unwrap(self._cfg_dict), memo), filename=self._filename)

    def __copy__(self):
        return Config(_unwrap(self._cfg_dict), filename=self._filename)

    def __getstate__(self):
        return (_unwrap(self._cfg_dict), self._filename)

    def __setstate__(self, state):
        object.__setattr__(self, '_cfg_dict', _wrap(state[0]))
        object.__setattr__(self, '_filename', state[1])

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __setitem__(self, name, value):
        self._cfg_dict[name] = _wrap(value)

    def __contains__(self, name):
        return name in self._cfg_dict

    def __iter__(self):
        return iter(self._cfg_dict)

    def __len__(self):
        return len(self._cfg_dict)

    def __repr__(self):
        return f'Config (path: {self._filename}): {self._cfg_dict!r}'

    def to_dict(self) -> dict:
        return _unwrap(self._cfg_dict)


class DictAction(argparse.Action):
    """argparse action for `--cfg-options k=v k2=[a,b]` (mmcv.DictAction)."""

    @staticmethod
    def _parse(val: str):
        for cast in (int, float):
            try:
                return cast(val)
            except ValueError:
                pass
        if val.lower() in ('true', 'false'):
            return val.lower() == 'true'
        if val == 'None':
            return None
        return val

    @classmethod
    def _parse_iter(cls, val: str):
        val = val.strip()
        if val[:1] in '([' and val[-1:] in ')]':
            inner = val[1:-1]
            items, depth, cur = [], 0, ''
            for ch in inner:
                if ch in '([':
                    depth += 1
                elif ch in ')]':
                    depth -= 1
                if ch == ',' and depth == 0:
                    items.append(cur)
                    cur = ''
                else:
                    cur += ch
            if cur.strip():
                items.append(cur)
            parsed = [cls._parse_iter(i) for i in items]
            return tuple(parsed) if val[0] == '(' else parsed
        if ',' in val:
            return [cls._parse(v.strip()) for v in val.split(',')]
        return cls._parse(val)

    def __call__(self, parser, namespace, values, option_string=None):
        options = {}
        for kv in values:
            key, val = kv.split('=', maxsplit=1)
            options[key] = self._parse_iter(val)
        setattr(namespace, self.dest, options)
