"""Import shim for the UNMODIFIED reference at /root/reference (test infrastructure only).

The reference's hot-path modules import six third-party packages that are absent in this
image and that the SIREN path never touches (compressai, matplotlib, omegaconf, py7zr,
tifffile, gurobipy).  This shim registers empty stand-ins for exactly those
names and puts /root/reference on sys.path, so `utils.Networks`, `utils.misc`, `utils.dataset`,
`utils.io`, `utils.ModelSave`, `utils.ssim` and `utils.adaptive_blocking.cal_divide_num` import and
run unmodified.

/root/reference exists only in the build container, never on the GPU box: only
`oracle/gen_golden.py` (fixture generation) and the `-m "not gpu"` tests that are skipped when
the directory is missing may call `load_reference()`.  Nothing in the product imports this file.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("BRIEF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utils", "Networks.py"))


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__brief_stub__ = True
        sys.modules[name] = mod
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


def _install_stubs() -> None:
    class _Absent:  # placeholder class: instantiating it means a non-hot-path feature was hit
        def __init__(self, *a, **k):
            raise RuntimeError("stubbed third-party class used: outside the SIREN hot path")

    def _have(name):
        try:
            importlib.import_module(name)
            return True
        except Exception:
            return False

    if not _have("compressai"):
        _stub("compressai")
        _stub("compressai.entropy_models", EntropyBottleneck=_Absent, GaussianConditional=_Absent)
    if not _have("matplotlib"):
        _stub("matplotlib")
        _stub("matplotlib.pyplot")
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not _have("omegaconf"):
        _stub("omegaconf", OmegaConf=_Absent)
        _stub("omegaconf.listconfig", ListConfig=list)
        _stub("omegaconf.dictconfig", DictConfig=dict)
        sys.modules["omegaconf"].listconfig = sys.modules["omegaconf.listconfig"]
        sys.modules["omegaconf"].dictconfig = sys.modules["omegaconf.dictconfig"]
    if not _have("py7zr"):
        _stub("py7zr", FILTER_BZIP2=0, FILTER_LZMA=1, FILTER_ZSTD=2, SevenZipFile=_Absent)
    if not _have("tifffile"):
        _stub("tifffile")
    if not _have("gurobipy"):
        _stub("gurobipy", GRB=_Absent)
    if not _have("prettytable"):
        _stub("prettytable", PrettyTable=_Absent)
    if not _have("skimage"):
        _stub("skimage")
        _stub("skimage.metrics")


_REF = None


def load_reference() -> types.SimpleNamespace:
    """Return a namespace of the reference's own hot-path modules (imported, not copied)."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    ns.Networks = importlib.import_module("utils.Networks")
    ns.misc = importlib.import_module("utils.misc")
    ns.dataset = importlib.import_module("utils.dataset")
    ns.io = importlib.import_module("utils.io")
    ns.ModelSave = importlib.import_module("utils.ModelSave")
    ns.ssim = importlib.import_module("utils.ssim")
    ns.adaptive_blocking = importlib.import_module("utils.adaptive_blocking")
    ns.tool = importlib.import_module("utils.tool")
    _REF = ns
    return ns


def load_main_samplers() -> types.SimpleNamespace:
    """The LIVE sampler classes (main.py:38-163) and loss (main.py:176-182) of the reference.  main.py cannot be imported
    (argparse / OmegaConf globals at module level), so these definitions are cut out of its syntax tree and executed
    unmodified in a namespace that holds exactly the names they use (torch, np, F, rearrange, typing,
    utils.dataset.create_*coords)."""
    import ast
    load_reference()
    with open(os.path.join(REFERENCE_ROOT, "main.py")) as fh:
        tree = ast.parse(fh.read())
    ns: dict = {}
    exec("import torch\nimport numpy as np\nfrom einops import rearrange\nfrom typing import *\n"
         "from utils.dataset import create_coords, create_flattened_coords\n", ns)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in ("RandomCubeSampler", "RandompointSampler"):
            exec(compile(ast.Module([node], []), os.path.join(REFERENCE_ROOT, "main.py"), "exec"), ns)
    # the loss is a closure inside NFGR.set_loss (main.py:176-182); it uses nothing but torch.nn.functional as F
    exec("import torch.nn.functional as F\n", ns)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "datal2":
            exec(compile(ast.Module([node], []), os.path.join(REFERENCE_ROOT, "main.py"), "exec"), ns)
    return types.SimpleNamespace(RandomCubeSampler=ns["RandomCubeSampler"], RandompointSampler=ns["RandompointSampler"],
                                 datal2=ns["datal2"])
