"""Reference module paths that are one file here.

The reference splits several packages into many files (``flashinfer/gemm/gemm_base.py``, ``gemm/routergemm.py``,
``logits_processor/{types,op,processors,...}.py`` ...); this framework keeps each of those packages in one or two
modules.  ``install`` registers a view module per reference file name so ``from flashinfer_b200.gemm.routergemm import
mm_M1_16_K7168_N256`` keeps working: attribute access on the view resolves against the owning package."""
from __future__ import annotations

import sys
import types
from typing import Iterable


def install(package: str, names: Iterable[str]) -> None:
    pkg = sys.modules[package]
    for name in names:
        full = f"{package}.{name}"
        if full in sys.modules or hasattr(pkg, name):
            continue
        view = types.ModuleType(full, f"View of {package} under the reference's module path '{name}'.")
        view.__getattr__ = lambda attr, _pkg=pkg, _full=full: _resolve(_pkg, _full, attr)   # PEP 562
        view.__dir__ = lambda _pkg=pkg: [n for n in dir(_pkg) if not n.startswith("_")]
        view.__package__ = package
        sys.modules[full] = view
        setattr(pkg, name, view)


def _resolve(pkg, full: str, attr: str):
    if attr.startswith("__"):
        raise AttributeError(attr)
    try:
        return getattr(pkg, attr)
    except AttributeError:
        raise AttributeError(f"module '{full}' has no attribute '{attr}'") from None
