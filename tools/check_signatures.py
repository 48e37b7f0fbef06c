"""Signature parity report: every public function / method of the reference package (parsed with ``ast`` - the reference is never
imported) against the callable of the same dotted name here.  Prints the reference parameters this library does not accept.

    python tools/check_signatures.py [/root/reference/flashinfer]

A parameter the reference accepts and this library lacks is a TypeError for a user who switches; the report is the work list.
Modules that only exist for other architectures / backends are listed in SKIP."""
from __future__ import annotations

import ast
import importlib
import inspect
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

SKIP_UNUSED = {"aot", "__main__", "_build_meta", "version", "tllm_enums", "tllm_utils", "jit", "cute_dsl", "triton", "gdn_kernels", "data", "profiler", "cli", "collect_env",
        "logits_processor", "testing", "tuning_configs", "artifacts", "compilation_context", "api_logging", "trace", "fi_trace", "dsv3_ops", "fused_moe", "gemm",
        "comm", "mamba", "quantization", "attention", "mla", "norm", "parallel_attention", "diffusion_ops", "topk"}


def ref_functions(path: Path):
    """{qualified name: [param names]} for top-level public defs and public methods of public classes of one source file."""
    tree = ast.parse(path.read_text())
    out = {}

    def params(fn):
        a = fn.args
        names = [x.arg for x in a.posonlyargs + a.args + a.kwonlyargs]
        pos = a.posonlyargs + a.args
        defaults = {x.arg: ast.unparse(d) for x, d in zip(pos[len(pos) - len(a.defaults):], a.defaults)}
        defaults.update({x.arg: ast.unparse(d) for x, d in zip(a.kwonlyargs, a.kw_defaults) if d is not None})
        return [n for n in names if n not in ("self", "cls")], defaults, [x.arg for x in pos if x.arg not in ("self", "cls")]

    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef)) and not node.name.startswith("_"):
            out[node.name] = params(node)
        elif isinstance(node, ast.ClassDef) and not node.name.startswith("_"):
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and (not sub.name.startswith("_") or sub.name == "__init__"):
                    out[f"{node.name}.{sub.name}"] = params(sub)
    return out


def _norm(text: str) -> str:
    """Comparable spelling of a default: enum members by their last name component, quotes unified, dtype prefixes dropped."""
    t = text.strip().replace('"', "'")
    if t.startswith("<") and ":" in t:                       # <Enum.Member: 3>
        t = t[1:].split(":")[0]
    for pre in ("torch.", "ActivationType.", "WeightLayout.", "RoutingMethodType.", "Fp8QuantizationType.", "SfLayout.", "GatedActType."):
        t = t.replace(pre, "")
    t = t.replace(".value", "")
    try:
        return repr(float(t)) if t.replace(".", "", 1).replace("-", "", 1).replace("e", "", 1).isdigit() else t
    except ValueError:
        return t


def main() -> int:
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    root = Path(args[0] if args else "/root/reference/flashinfer")
    files = sorted(p for p in root.rglob("*.py") if "gdn_kernels" not in p.parts and "data" not in p.parts and "triton" not in p.parts
                   and "cute_dsl" not in p.parts and "jit" not in p.parts and "tuning_configs" not in p.parts)
    n_checked = n_missing_fn = n_gap = 0
    for f in files:
        rel = f.relative_to(root).with_suffix("")
        mod_name = ".".join(rel.parts)
        if mod_name.endswith("__init__"):
            mod_name = mod_name[: -len(".__init__")] if "." in mod_name else ""
        if rel.parts[-1].startswith("_") and rel.parts[-1] != "__init__":
            continue
        try:
            mod = importlib.import_module("flashinfer_b200" + ("." + mod_name if mod_name else ""))
        except Exception:
            continue
        for qual, (names, ref_defaults, ref_pos) in ref_functions(f).items():
            obj = mod
            try:
                for part in qual.split("."):
                    obj = getattr(obj, part)
            except AttributeError:
                n_missing_fn += 1
                print(f"MISSING  {mod_name}.{qual}")
                continue
            try:
                sig = inspect.signature(inspect.unwrap(obj) if callable(obj) else obj)
            except (TypeError, ValueError):
                continue
            n_checked += 1
            ours = sig.parameters
            if any(p.kind is inspect.Parameter.VAR_KEYWORD for p in ours.values()):
                continue
            lack = [n for n in names if n not in ours]
            if lack:
                n_gap += 1
                print(f"PARAMS   {mod_name}.{qual}: {', '.join(lack)}")
            if "--defaults" in sys.argv:
                diffs = []
                for n, d in ref_defaults.items():
                    p = ours.get(n)
                    if p is None or p.default is inspect.Parameter.empty:
                        continue
                    if _norm(d) != _norm(repr(p.default)):
                        diffs.append(f"{n}: ref {d} / here {p.default!r}")
                if diffs:
                    print(f"DEFAULTS {mod_name}.{qual}: " + "; ".join(diffs))
            if "--order" in sys.argv:
                mine = [n for n, p in ours.items() if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD) and n not in ("self", "cls")]
                common = [n for n in ref_pos if n in mine]
                if common != [n for n in mine if n in common]:
                    print(f"ORDER    {mod_name}.{qual}: ref {common} / here {[n for n in mine if n in common]}")
    print(f"{n_checked} callables compared, {n_gap} with missing parameters, {n_missing_fn} missing callables")
    return 0


if __name__ == "__main__":
    sys.exit(main())
