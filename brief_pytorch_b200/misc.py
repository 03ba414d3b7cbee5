"""Host-side helpers with the reference's signatures (utils/misc.py): checkpoints, loss weights, optimiser /
scheduler configuration, dense reconstruction, block partition and merge, quality metrics."""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass, field
from typing import Callable, Iterable, List, Optional, Sequence

import numpy as np
import torch

from .dataset import create_flattened_coords
from .io import get_type_max


# ---- checkpoints / weights ------------------------------------------------------------------------------
def parse_checkpoints(checkpoints, max_steps: int) -> List[int]:
    """'none' | 'every_N' | 'a,b,c' -> sorted step list that always ends with max_steps (utils/misc.py:255-271;
    like the reference, a bare int is rejected by the 'every' membership test with TypeError)."""
    if checkpoints == "none":
        return [max_steps]
    if "every" in checkpoints:
        every = int(checkpoints.split("_")[1])
        return list(range(every, max_steps, every)) + [max_steps]
    return [int(tok) for tok in checkpoints.split(",") if int(tok) < max_steps] + [max_steps]


def _limits(data: np.ndarray, lo: float, hi: float):
    if not (0 <= lo <= hi <= get_type_max(data)):
        raise AssertionError("Improper range setting!")
    return lo, hi


def parse_weight(data: np.ndarray, weight_type_list: Iterable[str]) -> np.ndarray:
    """Per-voxel loss weights from the rule strings of Compress.loss.weight (utils/misc.py:272-307)."""
    data = np.asarray(data)
    weight = np.ones(data.shape, dtype=np.float32)
    for rule in weight_type_list:
        if rule == "none":
            continue
        kind, *args = rule.split("_")
        vals = [float(a) for a in args]
        if kind == "value":
            lo, hi = _limits(data, vals[0], vals[1])
            weight[(data >= lo) & (data <= hi)] = vals[2]
        elif kind == "quantile":
            sel = data[data >= vals[0]]
            lo, hi = _limits(data, np.quantile(sel, vals[1]), np.quantile(sel, vals[2]))
            weight[(data >= lo) & (data <= hi)] = vals[3]
        elif kind == "exp":
            weight = np.exp(-(-np.log(vals[1]) / vals[0]) * data)
        else:
            raise NotImplementedError(rule)
    return weight


def preprocess_is_identity(np_dtype, denoise_level, clip_range) -> bool:
    """For unsigned data, level <= 0 only ever rewrites zeros with zeros, and a clip to the whole dtype range changes
    nothing: the shipped configs (level 0, clip [0, 65535]) are this case."""
    tmax = 255 if np.dtype(np_dtype) == np.uint8 else 65535
    return np.dtype(np_dtype) in (np.uint8, np.uint16) and denoise_level <= 0 and clip_range[0] <= 0 and clip_range[1] >= tmax


def preprocess(data: np.ndarray, denoise_level, denoise_close, clip_range, device="cuda") -> np.ndarray:
    """Same call as the reference's preprocess (utils/misc.py:244-254) for uint8 / uint16 data [D,H,W,1] / [H,W,1]:
    the block goes to the device, the threshold + binary opening + clip run there (brief_preprocess, one bit per voxel
    for the morphology), and the result comes back.  The reference also zeroes its INPUT array in place; this mirror
    leaves the input untouched (main.py uses the returned array only)."""
    from .group import preprocess_
    data = np.asarray(data)
    if data.dtype not in (np.uint8, np.uint16) or data.shape[-1] != 1 or data.ndim not in (3, 4):
        raise NotImplementedError("preprocess on the device: uint8 / uint16 data with one channel")
    _limits(data, clip_range[0], clip_range[1])  # range_limit (utils/tool.py:26-30)
    flat = np.ascontiguousarray(data[..., 0])
    t = torch.from_numpy(flat.view(np.int16) if flat.dtype == np.uint16 else flat).to(device)
    preprocess_(t, denoise_level, denoise_close, clip_range, data.dtype.name)
    out = t.cpu().numpy()
    return (out.view(np.uint16) if data.dtype == np.uint16 else out)[..., None]


def quantile_from_histogram(hist: np.ndarray, ge_thres: float, q: float) -> float:
    """np.quantile(data[data >= ge_thres], q) (default 'linear' method) for integer-valued data given only its
    histogram `hist[v] = #voxels == v`: the two neighbouring order statistics are read off the cumulative counts and
    interpolated with numpy's own lerp formula, so the result is the float64 numpy returns."""
    hist = np.asarray(hist, dtype=np.int64)
    first = int(np.ceil(max(ge_thres, 0.0)))
    counts = hist.copy()
    counts[:min(first, counts.size)] = 0
    n = int(counts.sum())
    if n == 0:
        return float("nan")  # numpy: quantile of an empty selection
    cum = np.cumsum(counts)
    vidx = (n - 1) * np.float64(q)
    prev = int(np.floor(vidx))
    gamma = vidx - prev
    nxt = min(prev + 1, n - 1)
    a = np.float64(np.searchsorted(cum, prev + 1, side="left"))   # value of the order statistic with 0-based rank prev
    b = np.float64(np.searchsorted(cum, nxt + 1, side="left"))
    diff = b - a
    return float(b - diff * (1 - gamma)) if gamma >= 0.5 else float(a + diff * gamma)


def weight_rules_for_kernel(data, weight_type_list: Iterable[str], np_dtype=None):
    """Translate rule strings into on-chip (lo, hi, scale) triples; returns None when a rule needs the
    explicit weight volume ('exp', or more than 4 rules).  `data`: the raw block as a numpy array or as a CUDA tensor
    (uint16 as int16 bit patterns, with `np_dtype`); only a 'quantile' rule looks at the voxels — through a device
    histogram (brief_block_histogram) when the block is a CUDA tensor of uint8 / uint16."""
    dt = np.dtype(np_dtype) if np_dtype is not None else np.asarray(data).dtype
    probe = np.zeros(0, dt)
    rules = []
    hist = None
    for rule in weight_type_list:
        if rule == "none":
            continue
        kind, *args = rule.split("_")
        vals = [float(a) for a in args]
        if kind == "value":
            rules.append(_limits(probe, vals[0], vals[1]) + (vals[2],))
        elif kind == "quantile":
            if isinstance(data, torch.Tensor) and data.is_cuda and dt in (np.uint8, np.uint16):
                if hist is None:
                    from .group import block_histogram
                    hist = block_histogram(data, dt.name)
                ql, qh = quantile_from_histogram(hist, vals[0], vals[1]), quantile_from_histogram(hist, vals[0], vals[2])
            else:
                host = data.cpu().numpy() if isinstance(data, torch.Tensor) else np.asarray(data)
                host = host.view(np.uint16) if (dt == np.uint16 and host.dtype == np.int16) else host
                sel = host[host >= vals[0]]
                ql, qh = float(np.quantile(sel, vals[1])), float(np.quantile(sel, vals[2]))
            rules.append(_limits(probe, ql, qh) + (vals[3],))
        else:
            return None
    return rules if len(rules) <= 4 else None


def check_float_preprocess_is_identity(blocks, pre) -> None:
    """brief_preprocess handles uint8 / uint16 volumes.  For float32 blocks the configured preprocess is accepted only
    when it provably changes nothing (no voxel at or below the level, clip range covering the data); anything else is
    refused loudly instead of being skipped."""
    level, (lo, hi) = pre["denoise"]["level"], pre["clip"]
    for b in blocks:
        arr = b.data if b.data is not None else None
        vmin = float(arr.min()) if arr is not None else float(b.dev.min())
        vmax = float(arr.max()) if arr is not None else float(b.dev.max())
        if vmin <= level or vmin < lo or vmax > hi:
            raise NotImplementedError("Compress.preprocess on float32 data that the threshold / clip would change: "
                                      "brief_preprocess covers uint8 / uint16 volumes")


# ---- optimiser / schedule configuration -------------------------------------------------------------------
@dataclass
class FusedOptimizer:
    """What configure_optimizer returns here: the settings the fused per-network kernel consumes
    (torch.optim defaults: betas (0.9, 0.999), eps 1e-8, no weight decay)."""
    name: str
    lr: float
    betas: tuple = (0.9, 0.999)
    eps: float = 1e-8
    milestones: tuple = ()
    gamma: float = 1.0
    step_size: int = 0   # StepLR: a decay every step_size steps (milestones are then generated per horizon)
    cyclic: Optional[dict] = None   # CyclicLR keyword arguments: lr AND beta1 change every step (per_step)
    _cyc: Optional[tuple] = field(default=None, repr=False, compare=False)

    def per_step(self, first_step_1based: int, n: int):
        """CyclicLR (utils/misc.py:189-190): the (lr, beta1) of steps first .. first + n - 1, read off torch's OWN scheduler
        driven on a dummy parameter — host scalars only, exactly the values the reference's optimiser would hold (with
        cycle_momentum, torch cycles beta1 of Adam / Adamax between base_momentum and max_momentum)."""
        if self._cyc is None or self._cyc[2] > first_step_1based - 1:
            dummy = torch.nn.Parameter(torch.zeros(1))
            topt = getattr(torch.optim, self.name)([dummy], lr=self.lr)
            self._cyc = [topt, torch.optim.lr_scheduler.CyclicLR(topt, **self.cyclic), 0]
        topt, sch, _ = self._cyc
        out = []
        for t in range(first_step_1based, first_step_1based + n):
            while self._cyc[2] < t - 1:
                topt.step()
                sch.step()
                self._cyc[2] += 1
            grp = topt.param_groups[0]
            out.append((float(grp["lr"]), float(grp["betas"][0]) if "betas" in grp else self.betas[0]))
        return out

    def milestones_until(self, n_steps: int) -> tuple:
        """The schedule as MultiStepLR milestones for a run of n_steps steps (what SirenGroup.fit_run consumes)."""
        if self.step_size > 0:
            return tuple(range(self.step_size, int(n_steps), self.step_size))
        return self.milestones

    def lr_at(self, step_1based: int) -> float:
        if self.cyclic is not None:
            return self.per_step(step_1based, 1)[0][0]
        lr = self.lr
        for m in self.milestones_until(step_1based):
            if m <= step_1based - 1:
                lr *= self.gamma
        return lr


def configure_optimizer(parameters, optimizer: str, lr: float) -> FusedOptimizer:
    if optimizer not in ("Adam", "Adamax", "SGD"):
        raise NotImplementedError(optimizer)
    return FusedOptimizer(optimizer, float(lr))


def configure_lr_scheduler(optimizer: FusedOptimizer, lr_scheduler_opt) -> FusedOptimizer:
    opt = copy.deepcopy(dict(lr_scheduler_opt))
    name = opt.pop("name")
    if name == "MultiStepLR":
        optimizer.milestones = tuple(sorted(int(m) for m in opt["milestones"]))
        optimizer.gamma = float(opt.get("gamma", 0.1))
        optimizer.step_size, optimizer.cyclic = 0, None
    elif name == "StepLR":  # lr * gamma every step_size steps: MultiStepLR at the multiples of step_size
        optimizer.milestones, optimizer.step_size, optimizer.cyclic = (), int(opt["step_size"]), None
        optimizer.gamma = float(opt.get("gamma", 0.1))
        if optimizer.step_size < 1:
            raise ValueError("StepLR.step_size must be positive")
    elif name == "none":
        optimizer.milestones, optimizer.gamma, optimizer.step_size, optimizer.cyclic = (), 1.0, 0, None
    elif name == "CyclicLR":  # lr (and beta1) of every step come from torch's scheduler itself, see per_step
        if optimizer.name == "SGD" and opt.get("cycle_momentum", True):
            raise NotImplementedError("CyclicLR with cycle_momentum turns torch's SGD into momentum SGD; the fused SGD has none")
        optimizer.milestones, optimizer.gamma, optimizer.step_size, optimizer.cyclic = (), 1.0, 0, opt
        optimizer._cyc = None
    else:
        raise NotImplementedError(name)  # utils/misc.py:195-196
    return optimizer


# ---- dense reconstruction -----------------------------------------------------------------------------------
def reconstruct_flattened(data_shape: Sequence[int], sample_size: int, sample_nf: Callable, device: str = "cuda",
                          half: bool = False, coords_mode: str = "-1,1") -> torch.Tensor:
    """Evaluate the network on the dense grid of `data_shape` ([d,h,w,C] / [h,w,C]) -> fp32 tensor of that shape.
    When `sample_nf` is the forward of a fused SIREN the whole grid is produced by ONE decompress launch with
    on-chip coordinates (sample_size is then irrelevant to the result); any other callable is fed coordinate
    chunks of `sample_size` like the reference (utils/misc.py:59-92).  half=True: an fp16 tensor comes back, like the
    reference's; the fused path still evaluates fp32 coordinates (the reference rounds them to fp16 first)."""
    *cshape, channels = [int(x) for x in data_shape]
    owner = getattr(sample_nf, "__self__", None)
    from .Networks import SIREN
    if isinstance(owner, SIREN) and getattr(sample_nf, "__name__", "") == "forward":
        dims = tuple(cshape) if len(cshape) == 3 else (1, *cshape)
        grp = owner.fused_group(dims)
        grp.set_axes(0, coords_mode)
        out = grp.decompress("float32")[0].reshape(*cshape, channels)
        return out.half() if half else out   # half=True returns an fp16 tensor like the reference (utils/misc.py:69-70)
    with torch.no_grad():
        coords = create_flattened_coords(tuple(cshape), coords_mode).to(device)
        flat = torch.zeros((coords.shape[0], channels), device=device, dtype=torch.float16 if half else torch.float32)
        for s in range(0, coords.shape[0], sample_size):
            c = coords[s:s + sample_size]
            flat[s:s + sample_size] = sample_nf(c.half() if half else c)
    return flat.reshape(*cshape, channels)


# ---- block partition (the multi-network batch) ------------------------------------------------------------------
def cal_divide_num(d: int, h: int, w: int, Nb: int, param_size: float) -> np.ndarray:
    """Grid (nd,nh,nw) of proper divisors whose product is the largest <= Nb, ties -> most cubic blocks
    (utils/adaptive_blocking.py:425-460; Nb <= 0 -> param bytes / (4*1361))."""
    if Nb <= 0:
        Nb = max(1, int(param_size / (4 * 1361)))

    def divisors(n):
        return [1] + [i for i in range(2, n) if n % i == 0]

    best, best_key = None, None
    for nd in divisors(d):
        for nh in divisors(h):
            for nw in divisors(w):
                num = nd * nh * nw
                if num > Nb:
                    continue
                ext = np.array([d / nd, h / nh, w / nw])
                var = ((ext - ext.mean()) ** 2).mean()
                if best is None or num > best_key[0] or (num == best_key[0] and var < best_key[1]):
                    best, best_key = np.array([nd, nh, nw]), (num, var)
    return best


def divide_data(data: np.ndarray, divide_type: str):
    """'total_nd_nh_nw' / 'every_d_h_w' -> list of chunk dicts with inclusive 'd','h','w' ranges and the
    reference's chunk names (utils/misc.py:329-366, 3-D).  Returns (chunks, None): the preview image the
    reference also returns is not produced."""
    if data.ndim != 4:
        raise NotImplementedError("3-D volumes [d,h,w,c] only")
    kind, a, b, c = divide_type.split("_")
    a, b, c = int(a), int(b), int(c)
    if "total" in kind:
        ext = (int(data.shape[0] / a), int(data.shape[1] / b), int(data.shape[2] / c))
    elif "every" in kind:
        ext = (a, b, c)
    else:
        raise NotImplementedError(divide_type)
    cuts = [list(range(0, data.shape[k], ext[k])) + [data.shape[k]] for k in range(3)]
    chunks = []
    for z0, z1 in zip(cuts[0][:-1], cuts[0][1:]):
        for y0, y1 in zip(cuts[1][:-1], cuts[1][1:]):
            for x0, x1 in zip(cuts[2][:-1], cuts[2][1:]):
                blk = data[z0:z1, y0:y1, x0:x1]
                chunks.append({"data": blk, "d": [z0, z1 - 1], "h": [y0, y1 - 1], "w": [x0, x1 - 1],
                               "total_size": data.size, "size": blk.size,
                               "name": f"d_{z0}_{z1 - 1}-h_{y0}_{y1 - 1}-w_{x0}_{x1 - 1}"})
    return chunks, None


def variance_from_sums(s1: float, s2: float, n: int) -> float:
    """Population variance of a block from its exact integer sums (brief_block_stats): E[x^2] - mean^2 in float64.
    The reference's two-pass ((x - mean)^2).mean() agrees to ~1e-14 relative; budgets are not a bit-exact quantity."""
    m = s1 / n
    return float(s2 / n - m * m)


def cal_feature(image: np.ndarray) -> float:
    """utils/adaptive_blocking.py:16-24 for a [d,h,w,c] block: max / sum of the 3-D FFT magnitudes (the DC share of
    the spectrum), with the reference's int() truncations.  Host numpy like the reference: it runs once per block
    before the fit and numpy's FFT is what makes the budgets bit-identical."""
    if image.ndim != 4:
        raise NotImplementedError("cal_feature: [d,h,w,c] blocks")
    f = np.abs(np.fft.fft(np.fft.fft(np.fft.fft(image, axis=0), axis=1), axis=2))
    return int(f.max()) / int(f.sum())


def alloc_param(data_chunk_list: List[dict], param_size: float, param_alloc: str, param_size_thres: float):
    """Split the byte budget over blocks (equal | by_size | by_var | by_d | by_dv) and drop blocks below the
    threshold, re-allocating until stable (utils/misc.py:395-428).  A chunk may carry a precomputed 'var' (from the
    device statistics kernel) or 'feature'; otherwise they are computed from chunk['data'] like the reference."""
    chunks = list(data_chunk_list)

    def var_of(c):
        if "var" not in c:
            c["var"] = ((c["data"] - c["data"].mean()) ** 2).mean()
        return c["var"]

    def feat_of(c):
        if "feature" not in c:
            c["feature"] = cal_feature(c["data"])
        return c["feature"]

    while True:
        if param_alloc == "equal":
            for c in chunks:
                c["param_size"] = param_size / len(chunks)
        elif param_alloc == "by_size":
            for c in chunks:
                c["param_size"] = param_size * c["size"] / c["total_size"]
        elif param_alloc == "by_var":
            total = 0
            for c in chunks:
                total += var_of(c)
            for c in chunks:
                c["param_size"] = float(param_size * var_of(c) / total)
        elif param_alloc == "by_d":
            total = 0
            for c in chunks:
                total += 1 / feat_of(c)
            for c in chunks:
                c["param_size"] = float(param_size * (1 / feat_of(c)) / total)
        elif param_alloc == "by_dv":
            total = 0
            for c in chunks:
                total += c["size"] / feat_of(c)
            for c in chunks:
                c["param_size"] = float(param_size * (c["size"] / feat_of(c)) / total)
        else:
            raise NotImplementedError(param_alloc)
        kept = [c for c in chunks if c["param_size"] >= param_size_thres]
        if len(kept) == len(chunks):
            return kept
        chunks = kept


def merge_divided_data(decompressed_data_chunk_list: List[dict], data_shape) -> np.ndarray:
    """Zero canvas, add each block at its inclusive range, clip to the dtype maximum, cast (utils/misc.py:430-445)."""
    first = decompressed_data_chunk_list[0]["data"]
    canvas = np.zeros(data_shape, dtype=np.float32)
    for c in decompressed_data_chunk_list:
        canvas[c["d"][0]:c["d"][1] + 1, c["h"][0]:c["h"][1] + 1, c["w"][0]:c["w"][1] + 1] += c["data"]
    return canvas.clip(None, get_type_max(first)).astype(first.dtype)


# ---- quality metrics ------------------------------------------------------------------------------------------
def cal_mse(data1: np.ndarray, data2: np.ndarray):
    return ((data1 - data2) ** 2).mean()


def cal_psnr(origin_data: np.ndarray, decompressed_data: np.ndarray, data_range) -> float:
    err = origin_data / data_range - decompressed_data / data_range
    return -10 * np.log10(np.mean(np.power(err, 2)))


def _gauss_window(size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    x = torch.arange(size, dtype=torch.float) - size // 2
    g = torch.exp(-(x ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _ssim_slice(x: torch.Tensor, y: torch.Tensor, data_range: float) -> torch.Tensor:
    """2-D SSIM of [1,C,H,W] images: separable 11-tap Gaussian, valid convolution, K=(0.01,0.03) (utils/ssim.py)."""
    import torch.nn.functional as F
    ch = x.shape[1]
    win = _gauss_window().to(x.device, x.dtype)

    def blur(t):
        for axis, n in ((2, t.shape[2]), (3, t.shape[3])):
            if n >= win.numel():
                shape = [ch, 1, 1, 1]
                shape[axis] = -1
                t = F.conv2d(t, win.view(1, 1, -1, 1).expand(ch, 1, -1, 1) if axis == 2
                             else win.view(1, 1, 1, -1).expand(ch, 1, 1, -1), groups=ch)
        return t

    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mx, my = blur(x), blur(y)
    sxx, syy, sxy = blur(x * x) - mx * mx, blur(y * y) - my * my, blur(x * y) - mx * my
    cs = (2 * sxy + c2) / (sxx + syy + c2)
    return (((2 * mx * my + c1) / (mx * mx + my * my + c1)) * cs).flatten(2).mean(-1).mean()


def cal_ssim(origin_data: np.ndarray, decompressed_data: np.ndarray, data_range, device: str = "cpu") -> float:
    """Mean over depth slices of the 2-D SSIM (utils/misc.py:458-475)."""
    a = torch.from_numpy(np.ascontiguousarray(origin_data)).to(device)
    b = torch.from_numpy(np.ascontiguousarray(decompressed_data)).to(device)
    if a.dim() == 3:
        return float(_ssim_slice(a.permute(2, 0, 1)[None], b.permute(2, 0, 1)[None], data_range))
    total = 0.0
    for i in range(a.shape[0]):
        total += _ssim_slice(a[i].permute(2, 0, 1)[None], b[i].permute(2, 0, 1)[None], data_range)
    return float(total / a.shape[0])


def eval_performance(steps: int, data1: np.ndarray, data2: np.ndarray, Log=None, mse=True, psnr=True, ssim=True,
                     device: str = "cpu") -> dict:
    out = {"steps": steps}
    rng = get_type_max(data1)
    if str(device).startswith("cuda") and data1.dtype == data2.dtype and data1.dtype in (np.uint8, np.uint16) \
            and data1.size == int(np.prod(data1.shape[:3])) and min(data1.shape[1:3]) >= 11:
        # one fused launch on the device (brief_volume_quality) instead of float32 copies + five conv2d passes per slice
        from .group import volume_quality
        view = (lambda x: np.ascontiguousarray(x.reshape(x.shape[:3])).view(np.int16 if x.dtype == np.uint16 else x.dtype))
        q = volume_quality(torch.from_numpy(view(data1)).to(device), torch.from_numpy(view(data2)).to(device), float(rng),
                           data1.dtype.name)
        out.update({k: q[k] for k, on in (("mse", mse), ("psnr", psnr), ("ssim", ssim)) if on})
        if Log is not None:
            Log.log_metrics({k: v for k, v in out.items() if k != "steps"}, steps)
        return out
    a, b = data1.astype(np.float32), data2.astype(np.float32)
    if mse:
        out["mse"] = cal_mse(a, b)
    if psnr:
        out["psnr"] = cal_psnr(a, b, rng)
    if ssim:
        out["ssim"] = cal_ssim(a, b, rng, device)
    if Log is not None:
        Log.log_metrics({k: v for k, v in out.items() if k != "steps"}, steps)
    return out
