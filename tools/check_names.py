"""Undefined-global check for the package (no linter in the image): every name a function loads as a global must exist in its
module's namespace after import (or be a builtin).  ``python tools/check_names.py [package.module ...]``; exit code 1 on findings."""
import builtins
import dis
import importlib
import pathlib
import sys
import types

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def global_loads(code: types.CodeType):
    instructions = list(dis.get_instructions(code))
    local = {i.argval for i in instructions if i.opname == "STORE_NAME"}          # class bodies: names bound in the same body
    for ins in instructions:
        if ins.opname == "LOAD_GLOBAL" or (ins.opname == "LOAD_NAME" and ins.argval not in local):
            yield ins.argval, code.co_name, ins.positions.lineno if ins.positions else code.co_firstlineno
    for c in code.co_consts:
        if isinstance(c, types.CodeType):
            yield from global_loads(c)


def check(modname: str):
    mod = importlib.import_module(modname)
    path = getattr(mod, "__file__", None)
    if not path or not path.endswith(".py"):
        return []
    code = compile(pathlib.Path(path).read_text(), path, "exec")
    have = set(vars(mod)) | set(dir(builtins)) | {"__class__", "__file__", "__name__", "__doc__", "__annotations__", "__qualname__", "__module__"}
    return [(modname, fn, line, name) for name, fn, line in global_loads(code) if name not in have]


def main(argv):
    mods = argv or []
    if not mods:
        pkg = ROOT / "flashinfer_b200"
        for f in sorted(pkg.rglob("*.py")):
            rel = f.relative_to(ROOT).with_suffix("")
            if any(p in ("csrc", "_lib") for p in rel.parts):
                continue
            name = ".".join(rel.parts)
            mods.append(name[: -len(".__init__")] if name.endswith(".__init__") else name)
    bad = []
    for m in mods:
        try:
            bad += check(m)
        except Exception as exc:  # noqa: BLE001
            bad.append((m, "<import>", 0, f"{type(exc).__name__}: {exc}"))
    for m, fn, line, name in bad:
        print(f"{m}:{line} in {fn}: {name}")
    print(f"{len(mods)} modules checked, {len(bad)} findings")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
