"""Build selected native modules: python tools/b.py name [name ...] (verbose nvcc output)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flashinfer_b200 import jit
for n in sys.argv[1:]:
    t = time.time()
    jit.build_module(jit.REGISTRY[n], verbose=True)
    print(n, "ok", round(time.time() - t, 1), "s")
