"""Toolchain probes, flag builders and a ninja writer (reference flashinfer/jit/cpp_ext.py).

The in-process builder (``build_module`` in this package) runs ``nvcc -shared`` directly: one translation unit per module, seconds
per file, nothing to schedule.  This file serves the other consumers the reference has: CI / packaging that wants a ``build.ninja``
for all modules at once (``generate_ninja_build_for_op`` + ``run_ninja``), tools that ask which CUDA they are building against, and
users who pass extra flags through the environment."""
from __future__ import annotations

import os
import re
import shutil
import subprocess
from pathlib import Path
from typing import List, Optional, Sequence

from . import ARCH_FLAGS, COMMON_NVCC_FLAGS, INCLUDE_DIRS, LIB_DIR, debug_flags


def parse_env_flags(env_var_name: str) -> List[str]:
    """Whitespace-separated extra flags from the environment (``FLASHINFER_EXTRA_CUDAFLAGS`` ...), quotes honoured."""
    import shlex

    raw = os.environ.get(env_var_name, "")
    try:
        return shlex.split(raw)
    except ValueError:
        return raw.split()


def get_cuda_path() -> str:
    """CUDA toolkit root: ``$CUDA_HOME`` / ``$CUDA_PATH``, else the directory two levels above ``nvcc``, else /usr/local/cuda."""
    for var in ("CUDA_HOME", "CUDA_PATH"):
        v = os.environ.get(var)
        if v and os.path.isdir(v):
            return v
    nvcc = shutil.which("nvcc")
    if nvcc:
        return str(Path(nvcc).resolve().parent.parent)
    if os.path.isdir("/usr/local/cuda"):
        return "/usr/local/cuda"
    raise RuntimeError("CUDA toolkit not found: set CUDA_HOME or put nvcc on PATH")


def get_cuda_version():
    """Version of the nvcc that would compile the kernels (``packaging.version.Version``); falls back to torch's CUDA version."""
    from packaging.version import Version

    nvcc = os.path.join(get_cuda_path(), "bin", "nvcc")
    try:
        out = subprocess.run([nvcc, "--version"], capture_output=True, text=True, check=True).stdout
        m = re.search(r"release (\d+\.\d+)", out)
        if m:
            return Version(m.group(1))
    except (OSError, subprocess.CalledProcessError):
        pass
    import torch

    if torch.version.cuda is None:
        raise RuntimeError("neither nvcc nor a CUDA build of torch is available")
    return Version(torch.version.cuda)


def is_cuda_version_at_least(version_str: str) -> bool:
    from packaging.version import Version

    return get_cuda_version() >= Version(version_str)


def get_nvcc_parallelism_flags() -> List[str]:
    """``--threads N`` for nvcc's internal per-arch parallelism (``FLASHINFER_NVCC_THREADS``, default 1: one arch here)."""
    n = os.environ.get("FLASHINFER_NVCC_THREADS", "1")
    return ["--threads", n] if n.isdigit() and int(n) > 1 else []


def join_multiline(vs: Sequence[str]) -> str:
    return " $\n    ".join(vs)


def get_system_includes(cuda_home: Optional[str] = None) -> List[Path]:
    home = Path(cuda_home or get_cuda_path())
    cands = [home / "include", home / "targets" / "x86_64-linux" / "include", home / "include" / "cccl"]
    return [p for p in cands if p.is_dir()]


def build_cuda_cflags(extra_cuda_cflags: Optional[Sequence[str]] = None) -> List[str]:
    """The full nvcc flag list of a module build: sm_100a arch flags, the package's common flags, debug flags when enabled,
    include directories, the caller's extras and ``$FLASHINFER_EXTRA_CUDAFLAGS``."""
    flags = [*ARCH_FLAGS, *COMMON_NVCC_FLAGS, *debug_flags(), *get_nvcc_parallelism_flags()]
    for inc in INCLUDE_DIRS:
        flags += ["-I", str(inc)]
    flags += list(extra_cuda_cflags or [])
    flags += parse_env_flags("FLASHINFER_EXTRA_CUDAFLAGS")
    return flags


def build_cflags(extra_cflags: Optional[Sequence[str]] = None) -> List[str]:
    """Host-compiler flags (forwarded by nvcc through ``-Xcompiler``) for the few pure C++ sources (planners, schedulers)."""
    return ["-O3", "-std=c++17", "-fPIC", *list(extra_cflags or []), *parse_env_flags("FLASHINFER_EXTRA_CFLAGS")]


def generate_ninja_build_for_op(name: str, sources: Sequence[Path], extra_cflags: Optional[Sequence[str]] = None,
                                extra_cuda_cflags: Optional[Sequence[str]] = None, extra_ldflags: Optional[Sequence[str]] = None,
                                extra_include_dirs: Optional[Sequence[Path]] = None, needs_device_linking: bool = False) -> str:
    """Text of a ``build.ninja`` that produces ``_lib/<name>.so`` from ``sources`` with the flags of the in-process builder:
    one nvcc compile edge per source (so ninja can run them in parallel and rebuild only what changed) and one link edge."""
    cuda_flags = build_cuda_cflags(extra_cuda_cflags)
    for inc in extra_include_dirs or []:
        cuda_flags += ["-I", str(inc)]
    host = ",".join(build_cflags(extra_cflags))
    nvcc = os.path.join(get_cuda_path(), "bin", "nvcc")
    out_dir = LIB_DIR / "ninja" / name
    lines = ["ninja_required_version = 1.3", f"nvcc = {nvcc}", "", f"cuda_cflags = {join_multiline(cuda_flags)}", f"host_cflags = {host}",
             f"ldflags = {join_multiline(['-shared', *list(extra_ldflags or [])])}", "",
             "rule cuda_compile", "  depfile = $out.d", "  deps = gcc",
             "  command = $nvcc --generate-dependencies-with-compile --dependency-output $out.d $cuda_cflags -Xcompiler $host_cflags "
             + ("-dc" if needs_device_linking else "-c") + " $in -o $out", "",
             "rule link", "  command = $nvcc $in $ldflags -o $out", ""]
    objs = []
    for src in sources:
        obj = out_dir / (Path(src).stem + ".o")
        objs.append(str(obj))
        lines.append(f"build {obj}: cuda_compile {Path(src).resolve()}")
    lines += ["", f"build {LIB_DIR / (name + '.so')}: link {' '.join(objs)}", "", f"default {LIB_DIR / (name + '.so')}", ""]
    return "\n".join(lines)


def _get_num_workers() -> Optional[int]:
    v = os.environ.get("MAX_JOBS")
    return int(v) if v and v.isdigit() else None


def run_ninja(workdir: Path, ninja_file: Path, verbose: bool = False) -> None:
    """Run ninja on ``ninja_file`` (``MAX_JOBS`` bounds the parallelism); raises with the tool's output on failure."""
    Path(workdir).mkdir(parents=True, exist_ok=True)
    cmd = ["ninja", "-v", "-C", str(Path(workdir).resolve()), "-f", str(Path(ninja_file).resolve())]
    jobs = _get_num_workers()
    if jobs is not None:
        cmd += ["-j", str(jobs)]
    res = subprocess.run(cmd, capture_output=not verbose, text=True)
    if res.returncode != 0:
        raise RuntimeError("ninja build failed" + ("" if verbose else f":\n{res.stdout}\n{res.stderr}"))
