"""Build a variant of the library with extra -D flags and run a tool script against it:
    python tools/exp_variant.py NAME "-DFLAG1 -DFLAG2" [--build-only] -- script.py args...
The variant lives in brief_pytorch_b200/_lib/libbrief_NAME.so (git-ignored) and never replaces the product library."""
import os, runpy, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brief_pytorch_b200 import build as _b
name, flags = sys.argv[1], sys.argv[2].split()
rest = sys.argv[3:]
lib = os.path.join(_b.LIBDIR, f"libbrief_{name}.so")
base = [f for f in _b.NVCC_FLAGS if f != "--use_fast_math=false"]
srcs = [os.path.join(_b.CSRC, s) for s in _b.SOURCES]
if not os.path.exists(lib) or any(os.path.getmtime(os.path.join(_b.CSRC, f)) > os.path.getmtime(lib) for f in os.listdir(_b.CSRC)):
    subprocess.run([_b._nvcc(), *base, *flags, "-o", lib] + srcs, check=True)
if "--build-only" in rest:
    sys.exit(0)
os.environ["BRIEF_NO_BUILD"] = "1"
_b.LIBPATH = lib
i = rest.index("--")
sys.argv = rest[i + 1:]
runpy.run_path(sys.argv[0], run_name="__main__")
