"""ctypes binding of libbrief_b200.so — one Python prototype per symbol of include/brief_b200.h.

There is no fallback: if the shared library is missing it is built with nvcc (brief_pytorch_b200.build),
and if that is impossible importing this module raises.  Every compute entry point fails with
BriefError when no CUDA device is present.
"""
import ctypes as C
import os

from . import build as _build

c_i32, c_i64, c_u64, c_f32, c_vp = C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_void_p

PREC_FP32, PREC_F16, PREC_AUTO = 0, 1, 2
DT_U8, DT_U16, DT_F32 = 0, 1, 2
SAMPLE_FULL_BLOCK, SAMPLE_RANDOM_POINTS = 0, 1
OPT_ADAMAX, OPT_ADAM, OPT_SGD = 0, 1, 2
MAX_WEIGHT_RULES = 4


class BriefError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libbrief_b200 error {code}: {msg}")
        self.code = code


class NetDesc(C.Structure):
    _fields_ = [("coords_channel", c_i32), ("data_channel", c_i32), ("features", c_i32), ("layers", c_i32),
                ("w0", c_f32), ("w_hidden", c_f32), ("dims", c_i32 * 3)]


class WeightRule(C.Structure):
    _fields_ = [("lo", c_f32), ("hi", c_f32), ("scale", c_f32)]


class OptConfig(C.Structure):
    _fields_ = [("kind", c_i32), ("lr", c_f32), ("beta1", c_f32), ("beta2", c_f32), ("eps", c_f32),
                ("n_milestones", c_i32), ("milestones", c_i64 * 8), ("gamma", c_f32)]


# name -> (restype, argtypes); mirrors include/brief_b200.h one to one
PROTOTYPES = {
    "brief_abi_version": (c_i32, []),
    "brief_last_error": (C.c_char_p, []),
    "brief_device_count": (c_i32, [C.POINTER(c_i32)]),
    "brief_linspace": (c_i32, [c_f32, c_f32, c_i32, C.POINTER(c_f32)]),
    "brief_group_create": (c_i32, [C.POINTER(NetDesc), c_i32, c_i32, c_i32, C.POINTER(c_vp)]),
    "brief_group_destroy": (None, [c_vp]),
    "brief_group_num_nets": (c_i32, [c_vp]),
    "brief_group_param_count": (c_i32, [c_vp, c_i32]),
    "brief_group_precision": (c_i32, [c_vp, c_i32]),
    "brief_group_batch": (c_i32, [c_vp, c_i32]),
    "brief_group_set_params": (c_i32, [c_vp, c_i32, c_vp, c_vp]),
    "brief_group_get_params": (c_i32, [c_vp, c_i32, c_vp, c_vp]),
    "brief_group_get_grads": (c_i32, [c_vp, c_i32, c_vp, c_vp]),
    "brief_group_set_grads": (c_i32, [c_vp, c_i32, c_vp, c_vp]),
    "brief_group_get_opt_state": (c_i32, [c_vp, c_i32, c_vp, c_vp, c_vp]),
    "brief_group_reset_opt_state": (c_i32, [c_vp, c_vp]),
    "brief_group_set_axes": (c_i32, [c_vp, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "brief_group_bind_volume": (c_i32, [c_vp, c_i32, c_vp, c_i32, c_f32, c_f32, c_f32, c_f32, c_vp,
                                        C.POINTER(WeightRule), c_i32, c_f32]),
    "brief_group_set_sampler": (c_i32, [c_vp, c_i32, c_i32, c_i32]),
    "brief_group_set_cube_sampler": (c_i32, [c_vp, c_i32, c_i32, C.POINTER(c_i32)]),
    "brief_cube_indices": (c_i32, [c_vp, c_i32, c_vp, c_u64, c_u64, c_vp, c_vp]),
    "brief_group_set_stream": (c_i32, [c_vp, c_i32, C.c_uint32]),
    "brief_group_set_slicing": (c_i32, [c_vp, c_i32]),
    "brief_fit_step": (c_i32, [c_vp, c_vp, c_u64, c_u64, c_vp, c_vp]),
    "brief_fit_kernels": (c_i32, [c_vp, c_vp, c_u64, c_u64, c_vp]),
    "brief_opt_step": (c_i32, [c_vp, c_i32, c_f32, c_f32, c_f32, c_f32, c_i64, c_vp]),
    "brief_fit_run": (c_i32, [c_vp, C.POINTER(OptConfig), c_u64, c_i64, c_i64, c_vp, c_vp]),
    "brief_fit_step_host": (c_i32, [c_vp, c_vp, C.POINTER(OptConfig), c_u64, c_i64, c_vp, c_vp]),
    "brief_block_histogram": (c_i32, [c_vp, c_i64, c_i32, c_vp, c_i32, c_vp]),
    "brief_forward": (c_i32, [c_vp, c_i32, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "brief_group_set_denorm": (c_i32, [c_vp, c_i32, c_f32, c_f32, c_f32, c_f32]),
    "brief_decompress": (c_i32, [c_vp, C.POINTER(c_vp), c_i32, c_vp]),
    "brief_gather": (c_i32, [c_vp, c_i32, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "brief_sample_indices": (c_i32, [c_u64, c_u64, c_i32, c_i64, c_i64, c_vp, c_vp]),
    "brief_block_stats": (c_i32, [C.POINTER(c_vp), C.POINTER(c_i64), c_i32, c_i32, c_i32, C.POINTER(C.c_double), c_vp]),
    "brief_deblock": (c_i32, [c_vp, c_i32, c_i32, c_i32, C.POINTER(c_i32), c_i32, c_i32, c_i32, c_i32, C.POINTER(c_i32), c_i32, c_vp]),
    "brief_volume_quality": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, C.c_double, C.POINTER(C.c_double), c_i32, c_vp]),
    "brief_preprocess_scratch_bytes": (c_i64, [c_i32, c_i32, c_i32]),
    "brief_preprocess": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_i32, C.c_double, C.POINTER(c_i32), C.c_double, C.c_double, c_vp,
                                 c_i32, c_vp]),
    "brief_launch_count": (c_i64, []),
    "brief_reset_launch_count": (None, []),
}

_lib = None


def lib_path() -> str:
    return _build.LIBPATH


def load():
    """Load (building first if stale) and prototype the shared library."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIBPATH
    if os.environ.get("BRIEF_NO_BUILD") != "1":
        path = _build.build()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing and could not be built; brief_pytorch_b200 has no fallback path")
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.brief_abi_version() != 1:
        raise ImportError("libbrief_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int):
    if rc < 0:
        raise BriefError(rc, load().brief_last_error().decode("utf-8", "replace"))
    return rc
