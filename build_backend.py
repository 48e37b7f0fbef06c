"""In-tree PEP 517 build backend (parity: reference build_backend.py).

* every build writes ``flashinfer_b200/_build_meta.json`` (version, git revision, CUDA toolkit, target arch);
* ``FIB200_AOT=1 pip wheel .`` (or ``python -m build``) first compiles every native module for sm_100a
  (``flashinfer_b200.jit.build_all``) so that the wheel ships ``flashinfer_b200/_lib/*.so`` and needs no nvcc at run time;
  without it the wheel ships sources only and modules JIT-compile on first use;
* editable installs work from the source tree (the libraries are built in-tree).
"""
import json
import os
import subprocess
import sys

from setuptools import build_meta as _orig

ROOT = os.path.dirname(os.path.abspath(__file__))


def _git_rev() -> str:
    try:
        return subprocess.check_output(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, stderr=subprocess.DEVNULL, text=True).strip()
    except Exception:  # noqa: BLE001
        return "unknown"


def _nvcc_version() -> str:
    try:
        out = subprocess.check_output(["nvcc", "--version"], stderr=subprocess.DEVNULL, text=True)
        return out.strip().splitlines()[-1]
    except Exception:  # noqa: BLE001
        return "unavailable"


def _prepare(aot: bool) -> None:
    sys.path.insert(0, ROOT)
    built = []
    if aot:
        from flashinfer_b200 import jit

        built = jit.build_all(verbose=bool(os.environ.get("FIB200_JIT_VERBOSE")))
    from flashinfer_b200.version import __version__

    meta = {"version": __version__, "git": _git_rev(), "nvcc": _nvcc_version(), "arch": "sm_100a", "aot_modules": list(built)}
    with open(os.path.join(ROOT, "flashinfer_b200", "_build_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)


def _aot() -> bool:
    return os.environ.get("FIB200_AOT", "0") == "1"


def get_requires_for_build_wheel(config_settings=None):
    return _orig.get_requires_for_build_wheel(config_settings)


def get_requires_for_build_sdist(config_settings=None):
    return _orig.get_requires_for_build_sdist(config_settings)


def get_requires_for_build_editable(config_settings=None):
    return _orig.get_requires_for_build_editable(config_settings)


def prepare_metadata_for_build_wheel(metadata_directory, config_settings=None):
    _prepare(False)
    return _orig.prepare_metadata_for_build_wheel(metadata_directory, config_settings)


def prepare_metadata_for_build_editable(metadata_directory, config_settings=None):
    _prepare(False)
    return _orig.prepare_metadata_for_build_editable(metadata_directory, config_settings)


def build_wheel(wheel_directory, config_settings=None, metadata_directory=None):
    _prepare(_aot())
    return _orig.build_wheel(wheel_directory, config_settings, metadata_directory)


def build_editable(wheel_directory, config_settings=None, metadata_directory=None):
    _prepare(_aot())
    return _orig.build_editable(wheel_directory, config_settings, metadata_directory)


def build_sdist(sdist_directory, config_settings=None):
    _prepare(False)
    return _orig.build_sdist(sdist_directory, config_settings)
