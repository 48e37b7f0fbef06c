"""No undefined globals anywhere in the package (the image has no linter; tools/check_names.py disassembles every function and checks
the names it loads against its module's namespace) - the safety net for code motion between modules."""
import importlib.util
import os


def test_every_global_name_resolves():
    spec = importlib.util.spec_from_file_location("check_names", os.path.join(os.path.dirname(__file__), "..", "tools", "check_names.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.main([]) == 0
