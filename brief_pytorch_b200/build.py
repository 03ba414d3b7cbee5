"""Build libbrief_b200.so (the C-ABI shared library, include/brief_b200.h) in-tree with nvcc for sm_100a.

    python -m brief_pytorch_b200.build [--force] [--verbose]

The library has no torch dependency (plain CUDA runtime, static cudart); the built .so lives in
brief_pytorch_b200/_lib/ so that it travels with the source snapshot and is git-ignored.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIBPATH = os.path.join(LIBDIR, "libbrief_b200.so")
SOURCES = ["brief_capi.cu", "brief_simt.cu", "brief_opt.cu", "brief_tc.cu", "brief_tc_wide.cu", "brief_data.cu", "brief_deblock.cu", "brief_quality.cu",
           "brief_preprocess.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libbrief_b200.so cannot be built (set NVCC=/path/to/nvcc)")


def _stale() -> bool:
    if not os.path.exists(LIBPATH):
        return True
    t = os.path.getmtime(LIBPATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
           [os.path.join(os.path.dirname(HERE), "include", "brief_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source into the one shared library. Returns its path."""
    if not force and not _stale():
        return LIBPATH
    os.makedirs(LIBDIR, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    cmd = [_nvcc(), *flags, "-o", LIBPATH + ".tmp"] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        print(res.stdout, res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(LIBPATH + ".tmp", LIBPATH)
    return LIBPATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
