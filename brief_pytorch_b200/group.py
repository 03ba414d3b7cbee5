"""SirenGroup — Python handle on a BriefGroup (include/brief_b200.h): many independent per-block SIREN
networks fitted / evaluated together by grouped sm_100a kernel launches.

Replaces the reference's per-block process farm (main.py:547-579, utils/TasksManager.py) and the
per-step Python loop of main.py:385-400.  torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _cabi
from ._cabi import (DT_F32, DT_U16, DT_U8, OPT_ADAM, OPT_ADAMAX, OPT_SGD, PREC_AUTO, PREC_F16, PREC_FP32,
                    SAMPLE_FULL_BLOCK, SAMPLE_RANDOM_POINTS, check)

_PREC = {"fp32": PREC_FP32, "f16": PREC_F16, "auto": PREC_AUTO}  # "f16": tcgen05 kind::f16 with fp16 operands
_OPT = {"Adamax": OPT_ADAMAX, "Adam": OPT_ADAM, "SGD": OPT_SGD}
_NP2DT = {"uint8": DT_U8, "uint16": DT_U16, "float32": DT_F32}
_DT2TORCH = {DT_U8: torch.uint8, DT_U16: torch.int16, DT_F32: torch.float32}  # u16 held as int16 bit patterns


@dataclass
class NetSpec:
    """Architecture + block geometry of one network (SIREN kwargs, utils/Networks.py:246)."""
    features: int
    layers: int
    w0: float
    dims: Sequence[int]            # (d,h,w) or (h,w)
    coords_channel: int = 3
    data_channel: int = 1
    w_hidden: float = 30.0

    def param_count(self) -> int:
        c, o, f, L = self.coords_channel, self.data_channel, self.features, self.layers
        return c * f + f + (L - 2) * (f * f + f) + f * o + o


def schedule_window(lr: float, milestones: Sequence[int], gamma: float, steps_done: int):
    """BriefOptConfig carries at most 8 MultiStepLR milestones.  Returns (lr0, window, horizon): the configuration that
    reproduces the schedule `milestones` from step steps_done + 1 until `horizon` steps are complete (None: for ever).
    Up to 8 milestones pass through unchanged; longer lists (a StepLR expressed as milestones, utils/misc.py:191-192)
    are served window by window, with the decays already behind folded into lr0 by the same chained multiplication."""
    ms = sorted(int(m) for m in milestones)
    if len(ms) <= 8:
        return float(lr), ms, None
    lr0 = float(np.float32(lr))
    for m in ms:
        if m <= steps_done:
            lr0 *= float(np.float32(gamma))
    pending = [m for m in ms if m > steps_done]
    return lr0, pending[:8], (pending[8] if len(pending) > 8 else None)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def pack_module_params(module) -> np.ndarray:
    """Flatten a SIREN-like module (`.net[l][0].weight/.bias`) in the order utils/ModelSave.py writes."""
    parts = []
    for l in range(len(module.net)):
        parts.append(module.net[l][0].weight.detach().to("cpu", torch.float32).reshape(-1).numpy())
        parts.append(module.net[l][0].bias.detach().to("cpu", torch.float32).reshape(-1).numpy())
    return np.ascontiguousarray(np.concatenate(parts), dtype=np.float32)


def unpack_module_params(module, flat: np.ndarray) -> None:
    off = 0
    with torch.no_grad():
        for l in range(len(module.net)):
            lin = module.net[l][0]
            for p in (lin.weight, lin.bias):
                n = p.numel()
                p.data = torch.from_numpy(flat[off:off + n].reshape(tuple(p.shape)).copy()).to(p.device, p.dtype)
                off += n
    assert off == flat.size


class SirenGroup:
    def __init__(self, specs: Sequence[NetSpec], device: int | str | torch.device = 0, precision: str = "auto"):
        self._lib = _cabi.load()
        self._h = C.c_void_p()
        if not torch.cuda.is_available():
            raise _cabi.BriefError(-2, "no CUDA device: brief_pytorch_b200 has no CPU path")
        self.device = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        self.specs = list(specs)
        descs = (_cabi.NetDesc * len(self.specs))()
        for i, s in enumerate(self.specs):
            dims = tuple(int(x) for x in s.dims)
            if len(dims) == 2:
                dims = (1,) + dims
            descs[i] = _cabi.NetDesc(s.coords_channel, s.data_channel, s.features, s.layers, float(s.w0),
                                     float(s.w_hidden), (C.c_int32 * 3)(*dims))
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        with torch.cuda.device(self.device):
            torch.cuda.current_stream()  # make sure the primary context exists
            check(self._lib.brief_group_create(descs, len(self.specs), idx, _PREC[precision], C.byref(self._h)))
        self._keep: Dict[int, tuple] = {}  # tensors bound to the group must outlive it
        self._cube_count: Dict[int, int] = {}
        self.steps_done = 0

    # ---- life cycle -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.brief_group_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return len(self.specs)

    def precision(self, net: int) -> str:
        return {PREC_FP32: "fp32", PREC_F16: "f16"}[check(self._lib.brief_group_precision(self._h, net))]

    def param_count(self, net: int) -> int:
        return check(self._lib.brief_group_param_count(self._h, net))

    # ---- parameters ---------------------------------------------------------------------------------
    def set_params(self, net: int, flat: np.ndarray) -> None:
        flat = np.ascontiguousarray(flat, dtype=np.float32)
        assert flat.size == self.param_count(net), (flat.size, self.param_count(net))
        check(self._lib.brief_group_set_params(self._h, net, flat.ctypes.data_as(C.c_void_p), _stream(self.device)))

    def _get(self, fn, net: int) -> np.ndarray:
        out = np.empty(self.param_count(net), dtype=np.float32)
        check(fn(self._h, net, out.ctypes.data_as(C.c_void_p), _stream(self.device)))
        return out

    def get_params(self, net: int) -> np.ndarray:
        return self._get(self._lib.brief_group_get_params, net)

    def get_grads(self, net: int) -> np.ndarray:
        return self._get(self._lib.brief_group_get_grads, net)

    def set_grads(self, net: int, flat: np.ndarray) -> None:
        flat = np.ascontiguousarray(flat, dtype=np.float32)
        assert flat.size == self.param_count(net)
        check(self._lib.brief_group_set_grads(self._h, net, flat.ctypes.data_as(C.c_void_p), _stream(self.device)))

    def get_opt_state(self, net: int):
        m = np.empty(self.param_count(net), dtype=np.float32)
        v = np.empty_like(m)
        check(self._lib.brief_group_get_opt_state(self._h, net, m.ctypes.data_as(C.c_void_p),
                                                  v.ctypes.data_as(C.c_void_p), _stream(self.device)))
        return m, v

    def reset_opt_state(self) -> None:
        check(self._lib.brief_group_reset_opt_state(self._h, _stream(self.device)))
        self.steps_done = 0

    def load_module(self, net: int, module) -> None:
        self.set_params(net, pack_module_params(module))

    def store_module(self, net: int, module) -> None:
        unpack_module_params(module, self.get_params(net))

    def set_axes(self, net: int, coords_mode: str = "-1,1") -> None:
        """Per-axis tables from torch.linspace on the CPU — bit-identical to the reference's create_coords."""
        from .dataset import axis_table
        s = self.specs[net]
        dims = tuple(int(x) for x in s.dims)
        if len(dims) == 2:
            dims = (1,) + dims
        tabs = [np.ascontiguousarray(axis_table(n, coords_mode).numpy(), dtype=np.float32) for n in dims]
        check(self._lib.brief_group_set_axes(self._h, net, *[t.ctypes.data_as(C.c_void_p) for t in tabs],
                                             _stream(self.device)))

    # ---- data -------------------------------------------------------------------------------------------
    def bind_volume(self, net: int, raw: torch.Tensor, vmin: float, vmax: float, lo: float = 0.0, hi: float = 100.0,
                    weight: Optional[torch.Tensor] = None, rules: Sequence[Sequence[float]] = (), tau: float = 0.0,
                    np_dtype: Optional[str] = None) -> None:
        """raw: CUDA tensor with the block's voxels in d,h,w order (uint8 / int16-viewed uint16 / float32)."""
        assert raw.is_cuda and raw.is_contiguous()
        s = self.specs[net]
        assert raw.numel() == int(np.prod(s.dims)), (raw.shape, s.dims)
        if np_dtype is None:
            np_dtype = {torch.uint8: "uint8", torch.int16: "uint16", torch.float32: "float32"}[raw.dtype]
        if hasattr(torch, "uint16") and raw.dtype == getattr(torch, "uint16"):
            np_dtype = "uint16"
        if weight is not None:
            assert weight.is_cuda and weight.dtype == torch.float32 and weight.is_contiguous()
            assert weight.numel() == raw.numel()
        if len(rules) > _cabi.MAX_WEIGHT_RULES:
            raise NotImplementedError("more than 4 value rules: pass an explicit weight volume")
        arr = (_cabi.WeightRule * max(1, len(rules)))()
        for i, r in enumerate(rules):
            arr[i] = _cabi.WeightRule(float(r[0]), float(r[1]), float(r[2]))
        check(self._lib.brief_group_bind_volume(self._h, net, _ptr(raw), _NP2DT[np_dtype], float(vmin), float(vmax),
                                                float(lo), float(hi), _ptr(weight), arr, len(rules), float(tau)))
        self._keep[net] = (raw, weight)

    def set_denorm(self, net: int, vmin: float, vmax: float, lo: float = 0.0, hi: float = 100.0) -> None:
        check(self._lib.brief_group_set_denorm(self._h, net, float(vmin), float(vmax), float(lo), float(hi)))

    def set_sampler(self, net: int, mode: str, batch: int = 0) -> None:
        m = {"randomcube": SAMPLE_FULL_BLOCK, "full": SAMPLE_FULL_BLOCK, "randompoint": SAMPLE_RANDOM_POINTS}[mode]
        check(self._lib.brief_group_set_sampler(self._h, net, m, int(batch)))
        self._cube_count.pop(net, None)

    def set_cube_sampler(self, net: int, cube_count: int, cube_len: Sequence[int]) -> None:
        """The general RandomCubeSampler (main.py:38-125): `cube_count` windows of `cube_len` voxels (clamped to the
        block) per step.  cube_len covers the block's axes: (d,h,w), or (h,w) for a 2-D network."""
        clen = [int(c) for c in cube_len]
        if len(clen) == 2:
            clen = [1] + clen
        check(self._lib.brief_group_set_cube_sampler(self._h, net, int(cube_count), (C.c_int32 * 3)(*clen)))
        self._cube_count[net] = int(cube_count)

    def cube_indices(self, net: int, cube_ids: Optional[torch.Tensor] = None, seed: int = 0, step: int = 0) -> torch.Tensor:
        """Voxel indices (int64, cube after cube) of one step of a cube network: `cube_ids` = the step's window draws
        (torch.randint(0, pop_size, (cube_count,)), main.py:114) or None for the on-device stream at (seed, step)."""
        if cube_ids is not None:
            assert cube_ids.numel() == self._cube_count.get(net), "one window index per cube of the step"
            cube_ids = cube_ids.to(self.device, torch.int64).contiguous()
        out = torch.empty(self.batch(net), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            check(self._lib.brief_cube_indices(self._h, net, _ptr(cube_ids), seed, step, _ptr(out), _stream(self.device)))
        return out

    def batch(self, net: int) -> int:
        """Samples per step of the network under its current sampler."""
        return check(self._lib.brief_group_batch(self._h, net))

    def set_stream(self, net: int, stream_id: int) -> None:
        """Key of the network's on-device sampler stream (pass the block's global index when blocks are sharded)."""
        check(self._lib.brief_group_set_stream(self._h, net, int(stream_id)))

    def set_slicing(self, per_network: bool) -> None:
        """per_network=True: every network's result is bit-identical whatever else shares the GPU (slower)."""
        check(self._lib.brief_group_set_slicing(self._h, 1 if per_network else 0))

    # ---- hot path -----------------------------------------------------------------------------------------
    def fit_step(self, idx: Optional[torch.Tensor] = None, seed: int = 0, step: int = 0) -> torch.Tensor:
        """gather + forward + weighted L2 + backward for every network; returns the per-network loss."""
        loss = torch.empty(len(self.specs), dtype=torch.float32, device=self.device)
        if idx is not None:
            assert idx.is_cuda and idx.dtype == torch.int64 and idx.is_contiguous()
        with torch.cuda.device(self.device):
            check(self._lib.brief_fit_step(self._h, _ptr(idx), seed, step, _ptr(loss), _stream(self.device)))
        return loss

    def fit_kernel_only(self, idx: Optional[torch.Tensor] = None, seed: int = 0, step: int = 0) -> None:
        """Bench hook: launch only the fit kernel(s) (no partial reduction, no optimiser)."""
        with torch.cuda.device(self.device):
            check(self._lib.brief_fit_kernels(self._h, _ptr(idx), seed, step, _stream(self.device)))

    def opt_step(self, kind: str = "Adamax", lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 t: Optional[int] = None) -> None:
        if t is None:
            t = self.steps_done + 1
        with torch.cuda.device(self.device):
            check(self._lib.brief_opt_step(self._h, _OPT[kind], lr, betas[0], betas[1], eps, t, _stream(self.device)))
        self.steps_done = t

    def fit_run(self, n_steps: int, kind: str = "Adamax", lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                milestones: Sequence[int] = (), gamma: float = 0.2, seed: int = 42,
                loss_history: bool = False) -> Optional[torch.Tensor]:
        """n_steps iterations of main.py:385-400 enqueued from C without host synchronisation."""
        hist = torch.empty((n_steps, len(self.specs)), dtype=torch.float32, device=self.device) if loss_history else None
        done = 0
        while True:
            lr0, window, horizon = schedule_window(lr, milestones, gamma, self.steps_done)
            n = n_steps - done if horizon is None else min(n_steps - done, horizon - self.steps_done)
            cfg = _cabi.OptConfig(_OPT[kind], lr0, betas[0], betas[1], eps, len(window), (C.c_int64 * 8)(*window), gamma)
            with torch.cuda.device(self.device):
                check(self._lib.brief_fit_run(self._h, C.byref(cfg), seed, self.steps_done, n,
                                              _ptr(hist[done:]) if hist is not None and done < n_steps else None,
                                              _stream(self.device)))
            self.steps_done += n
            done += n
            if done >= n_steps:
                return hist

    def fit_step_host(self, host_idx: Optional[torch.Tensor], host_loss: Optional[torch.Tensor], kind: str = "Adamax",
                      lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, milestones: Sequence[int] = (),
                      gamma: float = 0.2, seed: int = 42) -> None:
        """One whole step (main.py:385-401) from HOST buffers as one CUDA-graph launch: pinned int64 sampler indices
        (all RANDOM_POINTS networks concatenated; None = on-device sampler) -> device, fit, optimiser, per-network loss
        -> pinned `host_loss`.  Asynchronous: synchronise the current stream (or an event recorded on it) before reading
        `host_loss` or refilling `host_idx`; alternate two buffer sets to keep the GPU busy."""
        for t in (host_idx, host_loss):
            if t is not None:
                assert (not t.is_cuda) and t.is_pinned() and t.is_contiguous()
        if host_idx is not None:
            assert host_idx.dtype == torch.int64
        if host_loss is not None:
            assert host_loss.dtype == torch.float32 and host_loss.numel() >= len(self.specs)
        lr0, window, _ = schedule_window(lr, milestones, gamma, self.steps_done)
        cfg = _cabi.OptConfig(_OPT[kind], lr0, betas[0], betas[1], eps, len(window), (C.c_int64 * 8)(*window), gamma)
        with torch.cuda.device(self.device):
            check(self._lib.brief_fit_step_host(self._h, _ptr(host_idx), C.byref(cfg), seed, self.steps_done,
                                                _ptr(host_loss), _stream(self.device)))
        self.steps_done += 1

    def forward(self, net: int, coords: torch.Tensor, return_layers: bool = False):
        """SIREN.forward on explicit coordinates [..., coords_channel] -> [..., 1] (fp32, CUDA)."""
        s = self.specs[net]
        assert coords.is_cuda and coords.dtype == torch.float32
        flat = coords.reshape(-1, s.coords_channel).contiguous()
        n = flat.shape[0]
        out = torch.empty((n, 1), dtype=torch.float32, device=self.device)
        layers = torch.zeros((s.layers - 1, n, s.features), dtype=torch.float32, device=self.device) if return_layers else None
        with torch.cuda.device(self.device):
            check(self._lib.brief_forward(self._h, net, _ptr(flat), n, _ptr(out), _ptr(layers), _stream(self.device)))
        out = out.reshape(*coords.shape[:-1], 1)
        return (out, layers) if return_layers else out

    def decompress(self, out_dtype: str = "uint16", out: Optional[List[Optional[torch.Tensor]]] = None,
                   nets: Optional[Sequence[int]] = None) -> List[Optional[torch.Tensor]]:
        """Dense-grid evaluation + inverse normalisation + truncating cast for every network (one launch
        per kernel family).  Returns one tensor per network shaped like its block (uint16 as int16 bits).
        `nets`: decode only these networks (the others' entries stay None / untouched) — lets a caller overlap the
        device->host copy of one block with the decode of the next."""
        dt = _NP2DT[out_dtype]
        want = set(range(len(self.specs))) if nets is None else set(int(i) for i in nets)
        if out is None:
            out = [torch.empty(tuple(int(x) for x in s.dims), dtype=_DT2TORCH[dt], device=self.device) if i in want else None
                   for i, s in enumerate(self.specs)]
        ptrs = (C.c_void_p * len(out))(*[(t.data_ptr() if (t is not None and i in want) else None) for i, t in enumerate(out)])
        with torch.cuda.device(self.device):
            check(self._lib.brief_decompress(self._h, ptrs, dt, _stream(self.device)))
        return out

    def decompress_to_host(self, out_dtype: str = "uint16", host_out: Optional[List[torch.Tensor]] = None,
                           dev_out: Optional[List[torch.Tensor]] = None, post=None) -> List[torch.Tensor]:
        """Decode block by block and overlap every block's device->host copy (copy stream, pinned destination) with the
        decode of the next one.  `post(i, dev_block)` (optional) runs on the decode stream between block i's decode and
        its copy (the Decompress.postprocess step).  Returns the pinned host tensors (uint16 as int16 bit patterns)."""
        dt = _DT2TORCH[_NP2DT[out_dtype]]
        n = len(self.specs)
        shapes = [tuple(int(x) for x in s.dims) for s in self.specs]
        if host_out is None:
            host_out = [torch.empty(sh, dtype=dt).pin_memory() for sh in shapes]
        if dev_out is None:
            dev_out = [torch.empty(sh, dtype=dt, device=self.device) for sh in shapes]
        main = torch.cuda.current_stream(self.device)
        copy = torch.cuda.Stream(device=self.device)
        for i in range(n):
            self.decompress(out_dtype, out=dev_out, nets=[i])
            if post is not None:
                post(i, dev_out[i])
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(copy):
                copy.wait_event(ev)
                host_out[i].copy_(dev_out[i], non_blocking=True)
        copy.synchronize()
        return host_out

    def gather(self, net: int, idx: Optional[torch.Tensor], batch: Optional[int] = None):
        """The reference sampler's (coords, data, weight) for the given voxel indices (main.py:156-160)."""
        s = self.specs[net]
        n = int(idx.numel()) if idx is not None else int(batch)
        coords = torch.empty((n, s.coords_channel), dtype=torch.float32, device=self.device)
        data = torch.empty((n, 1), dtype=torch.float32, device=self.device)
        weight = torch.empty((n, 1), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self._lib.brief_gather(self._h, net, _ptr(idx), n, _ptr(coords), _ptr(data), _ptr(weight),
                                         _stream(self.device)))
        return coords, data, weight


def sample_indices(seed: int, step: int, net: int, batch: int, pop: int, device="cuda") -> torch.Tensor:
    """Index stream of the on-device sampler (Philox4x32-10), int64 [batch]."""
    lib = _cabi.load()
    out = torch.empty(batch, dtype=torch.int64, device=device)
    with torch.cuda.device(out.device):
        check(lib.brief_sample_indices(seed, step, net, batch, pop, _ptr(out), _stream(out.device)))
    return out


def block_stats(blocks: Sequence[torch.Tensor], np_dtype: Optional[str] = None) -> np.ndarray:
    """min, max, sum, sum of squares of every raw block (CUDA tensors, contiguous) in one launch -> float64 [n, 4].
    Replaces the host numpy passes of normalize_data / alloc_param (utils/io.py:67-80, utils/misc.py:402-422)."""
    lib = _cabi.load()
    if not blocks:
        return np.zeros((0, 4))
    dev = blocks[0].device
    if np_dtype is None:
        np_dtype = {torch.uint8: "uint8", torch.int16: "uint16", torch.float32: "float32"}[blocks[0].dtype]
    for t in blocks:
        assert t.is_cuda and t.is_contiguous() and t.device == dev and t.dtype == blocks[0].dtype
    n = len(blocks)
    ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in blocks])
    sizes = (C.c_int64 * n)(*[t.numel() for t in blocks])
    out = np.zeros((n, 4), dtype=np.float64)
    with torch.cuda.device(dev):
        check(lib.brief_block_stats(ptrs, sizes, n, _NP2DT[np_dtype], dev.index if dev.index is not None else 0,
                                    out.ctypes.data_as(C.POINTER(C.c_double)), _stream(dev)))
    return out


def block_histogram(block: torch.Tensor, np_dtype: Optional[str] = None) -> np.ndarray:
    """Value histogram (int64 [256] / [65536]) of a raw uint8 / uint16 CUDA block (uint16 as int16 bit patterns) in one
    launch — the device half of the 'quantile' weight rule (utils/misc.py:298-305)."""
    lib = _cabi.load()
    assert block.is_cuda and block.is_contiguous()
    if np_dtype is None:
        np_dtype = {torch.uint8: "uint8", torch.int16: "uint16"}[block.dtype]
    out = np.zeros(256 if np_dtype == "uint8" else 65536, dtype=np.uint64)
    with torch.cuda.device(block.device):
        check(lib.brief_block_histogram(_ptr(block), block.numel(), _NP2DT[np_dtype], out.ctypes.data_as(C.c_void_p),
                                        block.device.index or 0, _stream(block.device)))
    return out.astype(np.int64)


def volume_quality(a: torch.Tensor, b: torch.Tensor, data_range: float, np_dtype: Optional[str] = None) -> dict:
    """mse / psnr / ssim of two CUDA volumes [D,H,W] (same dtype) in one launch — eval_performance (utils/misc.py:477-499)
    with the reference's SSIM (utils/ssim.py).  uint16 volumes may be held as int16 bit patterns."""
    lib = _cabi.load()
    assert a.is_cuda and b.is_cuda and a.shape == b.shape and a.dtype == b.dtype and a.dim() == 3
    a, b = a.contiguous(), b.contiguous()
    if np_dtype is None:
        np_dtype = {torch.uint8: "uint8", torch.int16: "uint16", torch.float32: "float32"}[a.dtype]
    out = (C.c_double * 3)()
    d, h, w = (int(x) for x in a.shape)
    with torch.cuda.device(a.device):
        check(lib.brief_volume_quality(_ptr(a), _ptr(b), _NP2DT[np_dtype], d, h, w, float(data_range), out,
                                       a.device.index or 0, _stream(a.device)))
    mse = out[0] / (d * h * w)
    return {"mse": mse, "psnr": float(-10.0 * np.log10(mse / (data_range * data_range))) if mse > 0 else float("inf"),
            "ssim": out[1] / out[2]}


def preprocess_(vol: torch.Tensor, denoise_level, denoise_close, clip_range, np_dtype: Optional[str] = None,
                scratch: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The reference's preprocess (utils/misc.py:244-254) on a CUDA block [D,H,W] (or [H,W]) IN PLACE: zero the
    binary opening of (v <= level) under a ones(close) structure (close False: the plain threshold), then clip.
    uint16 volumes may be held as int16 bit patterns.  Asynchronous on the current stream; returns `vol`."""
    lib = _cabi.load()
    assert vol.is_cuda and vol.is_contiguous() and vol.dim() in (2, 3)
    if np_dtype is None:
        np_dtype = {torch.uint8: "uint8", torch.int16: "uint16", torch.uint16: "uint16"}[vol.dtype]
    if vol.dim() == 2:  # image [H,W]: structure close[:2] (utils/misc.py:251)
        d, (h, w) = 1, (int(x) for x in vol.shape)
        close = None if denoise_close is False else [1, int(denoise_close[0]), int(denoise_close[1])]
    else:
        d, h, w = (int(x) for x in vol.shape)
        close = None if denoise_close is False else [int(c) for c in denoise_close[:3]]
    if vol.numel() == 0:
        return vol
    need = int(lib.brief_preprocess_scratch_bytes(d, h, w))
    if scratch is None or scratch.numel() * scratch.element_size() < need:
        scratch = torch.empty(need, dtype=torch.uint8, device=vol.device)
    c_close = (C.c_int32 * 3)(*close) if close is not None else None
    with torch.cuda.device(vol.device):
        check(lib.brief_preprocess(_ptr(vol), _NP2DT[np_dtype], d, h, w, float(denoise_level), c_close,
                                   float(clip_range[0]), float(clip_range[1]), _ptr(scratch), vol.device.index or 0,
                                   _stream(vol.device)))
    scratch.record_stream(torch.cuda.current_stream(vol.device))
    return vol


def launch_count() -> int:
    return int(_cabi.load().brief_launch_count())


def reset_launch_count() -> None:
    _cabi.load().brief_reset_launch_count()
