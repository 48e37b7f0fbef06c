"""Ahead-of-time build of every native module (reference flashinfer/aot.py:447-941).  With a single target
architecture there is no dtype x head-dim x arch cross product to enumerate: kernels dispatch on run-time dtype codes
and template instantiations live inside each translation unit, so AOT == build every ModuleSpec once."""
from __future__ import annotations

import argparse

from . import jit


def gen_all_modules():
    return list(jit.REGISTRY.values())


# ------------------------------------------------------------------ reference entry-point names (flashinfer/aot.py)
def parse_bool(s: str) -> bool:
    if s.lower() in ("true", "1", "yes", "on"):
        return True
    if s.lower() in ("false", "0", "no", "off"):
        return False
    raise ValueError(f"Invalid boolean value: {s}")


def parse_head_dim(head_dim: str):
    qk, vo = (int(x) for x in head_dim.split(","))
    return qk, vo


def detect_sm_capabilities() -> dict:
    """Which architectures the build targets: exactly one here."""
    return {"sm100a": True, "sm90": False, "sm103": False, "sm110": False, "sm120": False, "sm121": False}


def get_default_config() -> dict:
    return {"fa2_head_dim": [(128, 128)], "fa3_head_dim": [], "f16_dtype": ["float16", "bfloat16"], "f8_dtype": ["float8_e4m3fn", "float8_e5m2"],
            "use_sliding_window": [False, True], "use_logits_soft_cap": [False, True], "add_comm": True, "add_gemma": True,
            "add_oai_oss": True, "add_moe": True, "add_act": True, "add_misc": True, "add_xqa": True}


def _subset(*names):
    return [jit.REGISTRY[n] for n in names if n in jit.REGISTRY]


def gen_fa2(*args, **kwargs):
    """The reference enumerates fa2 template instances; the attention modules here carry their instances inside one TU each."""
    return _subset("decode_sm100", "prefill_sm100", "attention_generic")


def gen_fa3(*args, **kwargs):
    return []  # Hopper-only


def gen_attention(*args, **kwargs):
    return _subset("decode_sm100", "prefill_sm100", "mla_sm100", "pod_sm100", "attention_generic", "cascade", "page", "rope")


def gen_xqa(*args, **kwargs):
    return _subset("decode_sm100", "mla_sm100")


def register_default_modules() -> int:
    return len(gen_all_modules())


def compile_and_package_modules(out_dir=None, build_dir=None, project_root=None, config=None, verbose: bool = False,
                                skip_prebuilt: bool = True) -> None:
    """Build every module; with ``out_dir`` also copy the libraries there."""
    jit.build_all(verbose=verbose, force=not skip_prebuilt)
    if out_dir is not None:
        copy_built_kernels(out_dir)


def copy_built_kernels(out_dir) -> None:
    import shutil
    from pathlib import Path

    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    for so in jit.LIB_DIR.glob("*.so"):
        shutil.copy2(so, out / so.name)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser("flashinfer_b200.aot")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("-v", "--verbose", action="store_true")
    ap.add_argument("-j", "--jobs", type=int, default=None)
    a = ap.parse_args(argv)
    built = jit.build_all(verbose=a.verbose, jobs=a.jobs, force=a.force)
    print(f"[aot] {len(built)} modules in {jit.LIB_DIR}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
