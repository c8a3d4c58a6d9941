"""TEST INFRASTRUCTURE ONLY -- loads the *real* reference mapping module.

Only usable in the build container, where /root/reference exists (it does not
exist on the GPU box).  Used by tests/golden/make_golden.py to generate the
golden fixtures and by the `-m "not gpu"` tests that pin the C oracle against
the unmodified reference code.  Never imported by the product package.

The reference (`ivlnce_baselines/common/mapping_module/mapper.py`) does not
import cleanly here: `torch_scatter`, `habitat` and friends are missing and
`mapper.py:182-183` uses mutable dataclass defaults that Python >= 3.11
rejects.  SURVEY.md Appendix C lists the three load-time shims; this file is
our implementation of that recipe.  Nothing in the reference's arithmetic is
touched.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("IVLN_REFERENCE_ROOT", "/root/reference")
_MM_DIR = os.path.join(
    REFERENCE_ROOT, "ivlnce_baselines", "common", "mapping_module"
)
_PKG = "ivlnce_baselines.common.mapping_module"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_MM_DIR, "mapper.py"))


def scatter_max_first_index(src: torch.Tensor, index: torch.Tensor):
    """Stand-in for `torch_scatter.scatter_max` (torch-scatter==2.0.9, the
    reference's un-vendored dependency, requirements.txt:24; call site
    mapper.py:471-474).

    Published CPU semantics restated: groups are `index` values, the output has
    `index.max()+1` slots, `out[g]` is the maximum of the group, `arg[g]` the
    position of the FIRST element attaining it (serial loop with a strict `>`
    update) and `src.size(0)` for an empty group.
    """
    n = src.shape[0]
    groups = int(index.max().item()) + 1 if n > 0 else 0
    lowest = torch.finfo(src.dtype).min
    out = torch.full((groups,), lowest, dtype=src.dtype)
    out = out.scatter_reduce(0, index, src, "amax", include_self=True)
    positions = torch.arange(n, dtype=torch.long)
    is_max = src == out[index]
    cand = torch.where(is_max, positions, torch.full_like(positions, n))
    arg = torch.full((groups,), n, dtype=torch.long)
    arg = arg.scatter_reduce(0, index, cand, "amin", include_self=True)
    return out, arg


def _namespace(name: str) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__path__ = []  # mark as package
    sys.modules[name] = mod
    return mod


def load_reference_mapper() -> types.ModuleType:
    """Return the reference `mapper` module, executing its own source."""
    key = _PKG + ".mapper"
    if key in sys.modules and getattr(sys.modules[key], "_ivln_shimmed", False):
        return sys.modules[key]
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")

    # (1) empty namespace packages so ivlnce_baselines/__init__.py never runs
    for name in ("ivlnce_baselines", "ivlnce_baselines.common", _PKG):
        if name not in sys.modules:
            _namespace(name)

    # (2) torch_scatter stand-in
    ts = types.ModuleType("torch_scatter")
    ts.scatter_max = scatter_max_first_index
    sys.modules["torch_scatter"] = ts

    # (3) rednet stub (rednet.py:4 needs a torchvision symbol that is gone)
    rn = types.ModuleType(_PKG + ".rednet")

    class RedNet(torch.nn.Module):  # placeholder, never run by GT/known modes
        def __init__(self, *a, **k):
            super().__init__()

    rn.RedNet = RedNet
    sys.modules[_PKG + ".rednet"] = rn

    # projector package, loaded by path (pure torch)
    proj_dir = os.path.join(_MM_DIR, "projector")
    spec = importlib.util.spec_from_file_location(
        _PKG + ".projector",
        os.path.join(proj_dir, "__init__.py"),
        submodule_search_locations=[proj_dir],
    )
    proj = importlib.util.module_from_spec(spec)
    sys.modules[_PKG + ".projector"] = proj
    spec.loader.exec_module(proj)

    # mapper.py: exec its own text with the two dataclass-default lines fixed
    with open(os.path.join(_MM_DIR, "mapper.py"), "r") as f:
        src = f.read()
    a = "current_state: RobotCurrentState = RobotCurrentState()"
    b = "start_state: RobotStartState = RobotStartState()"
    assert a in src and b in src, "reference mapper.py changed"
    src = src.replace(
        a, "current_state: RobotCurrentState = field(default_factory=RobotCurrentState)"
    ).replace(
        b, "start_state: RobotStartState = field(default_factory=RobotStartState)"
    )
    mod = types.ModuleType(key)
    mod.__file__ = os.path.join(_MM_DIR, "mapper.py")
    sys.modules[key] = mod
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    mod._ivln_shimmed = True
    return mod


def load_reference_function(rel_path: str, name: str, class_name: str = None):
    """A single function of the reference, compiled from ITS OWN source text (found with `ast`), for modules whose
    imports (gym, habitat, lmdb ...) do not exist here: `rel_path` relative to the reference root, `name` a
    top-level function or -- with `class_name` -- a method (returned as a plain function taking `self`)."""
    import ast
    import collections
    import typing

    import numpy as np

    path = os.path.join(REFERENCE_ROOT, rel_path)
    with open(path, "r") as f:
        src = f.read()
    tree = ast.parse(src)
    body = tree.body
    if class_name is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name).body
    wanted = [n for n in body if isinstance(n, ast.FunctionDef) and n.name == name]
    helpers = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name != name] if class_name is None else []
    assert wanted, f"{name} not found in {path}"
    ns = {"np": np, "torch": torch, "defaultdict": collections.defaultdict, "DictTree": dict}
    ns.update({k: getattr(typing, k) for k in ("Any", "DefaultDict", "Dict", "List", "Optional", "Set", "Tuple")})
    for node in helpers + wanted:
        try:
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, ns)
        except Exception:
            if node in wanted:
                raise
    return ns[name]
