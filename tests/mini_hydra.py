"""A few-dozen-line stand-in for hydra.compose + OmegaConf (neither is installable here), enough to compose the reference's
own YAML groups the way scripts/eval.py:45-58 does: `defaults:` lists (relative / absolute group paths, the `robot_base`
registration quirk of eval.py:47-56), `# @package _global_` files merged at the root in defaults order, `${a.b.c}`
interpolations resolved to the SAME node object (so that `env.config.robot` and `robot` alias each other like OmegaConf's
lazy interpolation does), the `${len:...}` resolver, and a dict type with attribute access.  TEST INFRASTRUCTURE."""
from __future__ import annotations

import re
from pathlib import Path
from typing import List

import yaml


class Node(dict):
    """dict with attribute access (what the env layer needs from a DictConfig)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def copy(self):
        return Node(dict.copy(self))


def _wrap(x):
    if isinstance(x, dict):
        return Node({k: _wrap(v) for k, v in x.items()})
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x


def _merge(dst: Node, src: Node) -> Node:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v
    return dst


def _find(search: List[Path], name: str) -> Path:
    for root in search:
        p = root / f"{name}.yaml"
        if p.exists():
            return p
    raise FileNotFoundError(name)


def _load(search: List[Path], name: str, out: Node) -> None:
    path = _find(search, name)
    raw = yaml.safe_load(path.read_text()) or {}
    group = str(Path(name).parent)
    for d in raw.pop("defaults", []):
        if d == "_self_":
            continue
        if isinstance(d, dict):                      # "group: option"
            (g, opt), = d.items()
            d = f"{g}/{opt}"
        if d.startswith("/"):
            child = d[1:]
        else:
            child = f"{group}/{d}" if group != "." else d
            try:
                _find(search, child)
            except FileNotFoundError:                # robot/go2/go2.yaml -> `robot_base` lives one level up (eval.py:47-56)
                child = f"{Path(group).parent}/{d}"
        _load(search, child, out)
    _merge(out, _wrap(raw))


_REF = re.compile(r"^\$\{([^:{}]+)\}$")
_LEN = re.compile(r"^\$\{len:\$\{([^{}]+)\}\}$")


def _lookup(root: Node, dotted: str):
    cur = root
    for k in dotted.split("."):
        if not isinstance(cur, dict) or k not in cur:
            raise KeyError(dotted)
        cur = cur[k]
    return cur


def _resolve(root: Node, node, depth=0):
    items = node.items() if isinstance(node, dict) else enumerate(node)
    for k, v in list(items):
        if isinstance(v, str):
            for _ in range(8):                       # chains of references
                m, ml = _REF.match(v) if isinstance(v, str) else None, _LEN.match(v) if isinstance(v, str) else None
                try:
                    if m:
                        v = _lookup(root, m.group(1))
                    elif ml:
                        tgt = _lookup(root, ml.group(1))
                        v = len(tgt) if tgt is not None else 0
                    else:
                        break
                except KeyError:
                    break                            # stays an unresolved string, like a lazy OmegaConf node nobody reads
            node[k] = v
        if isinstance(node[k], (dict, list)) and depth < 24:
            _resolve(root, node[k], depth + 1)


def compose(config_name: str, search: List[Path]) -> Node:
    cfg = Node()
    _load([Path(s) for s in search], config_name, cfg)
    _resolve(cfg, cfg)
    return cfg
