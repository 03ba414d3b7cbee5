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
SOURCES = ["brief_capi.cu", "brief_simt.cu", "brief_opt.cu", "brief_tc.cu", "brief_tc_wide.cu", "brief_tc_lw.cu", "brief_data.cu", "brief_deblock.cu", "brief_quality.cu",
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
    """Compile every CUDA source (one object per translation unit, in parallel, rebuilt only when the source or a
    header is newer) and link the one shared library. Returns its path."""
    if not force and not _stale():
        return LIBPATH
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f not in ("--use_fast_math=false", "-shared")]
    if verbose:
        flags += ["-Xptxas", "-v"]
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + \
              [os.path.join(os.path.dirname(HERE), "include", "brief_b200.h")]
    newest_header = max(os.path.getmtime(h) for h in headers)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(objdir, src[:-3] + ".o")
        path = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), newest_header):
            return obj, ""
        res = subprocess.run([nvcc, *flags, "-c", "-o", obj, path], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
        return obj, res.stdout + res.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), max(1, (os.cpu_count() or 2)))) as pool:
        results = list(pool.map(compile_one, SOURCES))
    if verbose:
        for _, log in results:
            print(log)
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-o",
                          LIBPATH + ".tmp"] + [o for o, _ in results], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    os.replace(LIBPATH + ".tmp", LIBPATH)
    return LIBPATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
