"""A ~150-line stand-in for `hydra.compose` + `hydra.utils.instantiate` (Hydra/OmegaConf are not installed in this image).

It reads the REFERENCE's YAML tree unmodified (`configs/model/sam_pt.yaml` and its config groups) -- defaults lists with
`group: option`, `group@package: option`, `override group: option`, `_self_`; relative interpolations `${..key}`;
`${hydra:runtime.cwd}`; `_target_` / `_partial_` instantiation -- which is everything those files use.  With it, the
drop-in packages are constructed from the very dotted paths the reference's configs name."""
from __future__ import annotations

import importlib
import os
import re
from functools import partial
from typing import Any, Dict, List, Optional

import yaml


def _deep_merge(dst: Dict, src: Dict) -> Dict:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _deep_merge(dst[k], v)
        else:
            dst[k] = v
    return dst


def _set_at(root: Dict, package: str, value: Dict) -> None:
    if not package:
        _deep_merge(root, value)
        return
    node = root
    parts = package.split(".")
    for p in parts[:-1]:
        node = node.setdefault(p, {})
    if isinstance(node.get(parts[-1]), dict):
        _deep_merge(node[parts[-1]], value)
    else:
        node[parts[-1]] = value


def _load_group(config_dir: str, group_dir: str, option: str, overrides: Dict[str, str]) -> Dict:
    """Load `<config_dir>/<group_dir>/<option>.yaml`, resolving its own defaults list relative to `group_dir`."""
    path = os.path.join(config_dir, group_dir, option + ".yaml")
    with open(path) as f:
        raw = yaml.safe_load(f) or {}
    defaults = raw.pop("defaults", [])
    out: Dict[str, Any] = {}
    self_done = False
    # `override x: y` entries replace the option chosen for group x by an earlier (inherited) defaults entry
    local_over = dict(overrides)
    for d in defaults:
        if isinstance(d, dict):
            (k, v), = d.items()
            if k.startswith("override "):
                local_over[k[len("override "):].strip().split("@")[0]] = v
    for d in defaults:
        if d == "_self_":
            _deep_merge(out, raw)
            self_done = True
            continue
        if isinstance(d, str):  # sibling file in the same group (e.g. `- sam_vit_base`)
            _deep_merge(out, _load_group(config_dir, group_dir, d, local_over))
            continue
        (k, v), = d.items()
        if k.startswith("override "):
            continue
        grp, _, pkg = k.partition("@")
        grp = grp.strip()
        opt = local_over.get(grp, overrides.get(os.path.join(group_dir, grp), v))
        sub = _load_group(config_dir, os.path.join(group_dir, grp), opt, {})
        _set_at(out, pkg.strip() if pkg else grp.split("/")[-1], sub)
    if not self_done:
        _deep_merge(out, raw)
    return out


_INTERP = re.compile(r"\$\{\s*([^}]+?)\s*\}")


def _resolve(root: Dict, cwd: str) -> Dict:
    def lookup(path_keys: List[str], rel: str):
        dots = len(rel) - len(rel.lstrip("."))
        key = rel.lstrip(".")
        base = path_keys[: len(path_keys) - dots] if dots else []
        node: Any = root
        for p in base + key.split("."):
            node = node[int(p)] if isinstance(node, list) else node[p]
        return node

    def walk(node, path_keys):
        if isinstance(node, dict):
            return {k: walk(v, path_keys + [k]) for k, v in node.items()}
        if isinstance(node, list):
            return [walk(v, path_keys + [str(i)]) for i, v in enumerate(node)]
        if isinstance(node, str):
            m = _INTERP.fullmatch(node.strip())
            if m:
                expr = m.group(1)
                if expr.startswith("hydra:runtime.cwd"):
                    return cwd
                return walk(lookup(path_keys, expr), path_keys)
            return _INTERP.sub(lambda mm: cwd if mm.group(1).startswith("hydra:runtime.cwd") else str(lookup(path_keys, mm.group(1))), node)
        return node

    return walk(root, [])


def compose_model(config_dir: str, overrides: Optional[Dict[str, Any]] = None, cwd: Optional[str] = None) -> Dict:
    """Compose `configs/model/sam_pt.yaml`.  `overrides`: group choices ("point_tracker": "pips",
    "sam@sam_predictor.sam_model": "sam_vit_base") and dotted value overrides ("sam_predictor._target_": "...")."""
    overrides = dict(overrides or {})
    groups = {k.split("@")[0]: v for k, v in overrides.items() if k.split("@")[0] in ("point_tracker", "sam")}
    cfg = _load_group(config_dir, "model", "sam_pt", groups)
    for k, v in overrides.items():
        if k.split("@")[0] in ("point_tracker", "sam"):
            continue
        node = cfg
        parts = k.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = v
    return _resolve(cfg, cwd or os.getcwd())


def _locate(path: str):
    mod, _, name = path.rpartition(".")
    return getattr(importlib.import_module(mod), name)


def instantiate(cfg: Any) -> Any:
    """hydra.utils.instantiate (recursive, `_target_` / `_partial_`)."""
    if isinstance(cfg, list):
        return [instantiate(v) for v in cfg]
    if not isinstance(cfg, dict):
        return cfg
    if "_target_" not in cfg:
        return {k: instantiate(v) for k, v in cfg.items()}
    kwargs = {k: instantiate(v) for k, v in cfg.items() if k not in ("_target_", "_partial_", "_recursive_")}
    target = _locate(cfg["_target_"])
    if cfg.get("_partial_", False):
        return partial(target, **kwargs)
    return target(**kwargs)
