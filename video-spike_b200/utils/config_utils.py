"""YAML configuration handling with the semantics of the reference's src/utils/config_utils.py:
dot-access dictionaries (:6-15), `include:<path>` expansion (:20-30), recursive override merge
(:36-52), `update_config` (:59-75), CLI `key=value` parsing (:94-141).  The YAML schema of
config/model/*.yaml and config/train/*.yaml is unchanged, so reference config files load as-is.
"""
from __future__ import annotations

import argparse

import yaml

_INCLUDE = "include"


class DictConfig(dict):
    """dict whose items can be read as attributes; nested dicts are wrapped on access."""

    def __getattr__(self, name):
        try:
            item = self[name]
        except KeyError:
            raise KeyError(name)
        return DictConfig(item) if isinstance(item, dict) else item

    def get_dict(self):
        return super()


def _read_yaml(path):
    with open(path, "r") as fh:
        return yaml.safe_load(fh)


def unpack_config_rec(node):
    """Replace every string of the form 'include:<file>' by the parsed YAML of that file, depth first."""
    if isinstance(node, str):
        head, _, rest = node.partition(":")
        if head == _INCLUDE and rest:
            node = _read_yaml(node.split(":")[1])
    if isinstance(node, dict):
        for key in node:
            node[key] = unpack_config_rec(node[key])
    return node


def update_config_rec(base, override):
    """Overlay `override` on `base`.  Dict overrides recurse (creating missing keys, and replacing a
    non-dict base by {}); anything else replaces the base value outright."""
    if not isinstance(override, dict):
        return override
    if not isinstance(base, dict):
        base = {}
    for key, val in override.items():
        base[key] = update_config_rec(base.get(key, {}) if key in base else {}, val)
    return base


def update_config(default_config, config=None):
    """default_config / config may each be a dict or a path to a YAML file.  With config=None the
    defaults are returned with their includes expanded.  A non-dict, non-path `default_config`
    (e.g. an argparse Namespace, as src/train.py:30 passes) is discarded by the merge, which makes
    that call a plain copy of `config` -- behaviour kept on purpose."""
    if isinstance(default_config, str):
        default_config = _read_yaml(default_config)
    if config is None:
        config = default_config
    if isinstance(config, str):
        config = _read_yaml(config)
    return DictConfig(update_config_rec(unpack_config_rec(default_config), unpack_config_rec(config)))


class ParseKwargs(argparse.Action):
    """argparse action collecting `key=value` tokens into a dict."""

    def __call__(self, parser, namespace, values, option_string=None):
        collected = {}
        for token in values:
            key, val = token.split("=")
            collected[key] = val
        setattr(namespace, self.dest, collected)


def convert_to_dtype(value):
    """String flag -> list / None / bool / int / float / str (in that order of attempts)."""
    value = value.strip()
    if value[0] == "[" and value[-1] == "]":
        return [convert_to_dtype(v) for v in value[1:-1].split(",")]
    if value in ("null", "None", "none"):
        return None
    if value in ("true", "True"):
        return True
    if value in ("false", "False"):
        return False
    if value.isdigit() or value.replace("-", "").isdigit():
        return int(value)
    try:
        return float(value)
    except Exception:
        return value


def config_from_kwargs(kwargs):
    """{'a.b.c': 'v'} -> DictConfig({'a': {'b': {'c': v}}}) with values converted by convert_to_dtype."""
    tree = {}
    for dotted, raw in (kwargs or {}).items():
        *parents, leaf = dotted.split(".")
        node = tree
        for part in parents:
            node = node.setdefault(part, {})
        node[leaf] = convert_to_dtype(raw)
    return DictConfig(tree)
