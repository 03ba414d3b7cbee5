"""CPU oracle: a restatement of BRIEF's SIREN fit / decompress hot path (TEST INFRASTRUCTURE ONLY).

This file restates, on the CPU in fp32 with torch + numpy, the algorithm of the reference
RichealYoung/BRIEF_PyTorch for the path named in BASELINE.json (`north_star`).  It is the
checker, never the product: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it.  The product package
`brief_pytorch_b200` never imports anything from `oracle/`.

Where the arithmetic lives: the reference's arithmetic is third-party `torch` (unpinned in
`requirements.txt:1`; this image has torch 2.11.0).  The oracle therefore calls the same torch
CPU primitives (`nn.Linear` init, `F.linear`, `torch.sin`, `F.mse_loss`, autograd,
`torch.optim.Adamax/Adam/SGD`, `MultiStepLR`, `torch.linspace`, `torch.randint`) in the same
order as the reference, and is pinned by `oracle/gen_golden.py`, which runs the UNMODIFIED
reference modules (imported from /root/reference through `oracle/refshim.py`) on the same seeds
and asserts bit-equality before writing `tests/golden/*.npz`.  Parity status: PINNED against
the reference's own code run in the build container (the reference ships no golden vectors or
tests of its own).

Every function cites the reference file:line it follows (relative to /root/reference).
"""
from __future__ import annotations

import copy
import math
import os
import shutil
import struct
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

HIDDEN_OMEGA = 30.0  # Sine() default, utils/Networks.py:227-229


# --------------------------------------------------------------------------------------
# Network: constructor, init, forward            utils/Networks.py:215-271, 291-314, 800-802
# --------------------------------------------------------------------------------------
class _Sine(nn.Module):
    """utils/Networks.py:227-234 : sin(w0 * x)."""

    def __init__(self, w0: float = HIDDEN_OMEGA):
        super().__init__()
        self.w0 = w0

    def forward(self, x):
        return torch.sin(self.w0 * x)


def _sine_init(m):  # utils/Networks.py:215-220
    with torch.no_grad():
        if hasattr(m, "weight"):
            n = m.weight.size(-1)
            m.weight.uniform_(-np.sqrt(6 / n) / 30, np.sqrt(6 / n) / 30)


def _first_layer_sine_init(m):  # utils/Networks.py:221-226
    with torch.no_grad():
        if hasattr(m, "weight"):
            n = m.weight.size(-1)
            m.weight.uniform_(-1 / n, 1 / n)


class OracleSIREN(nn.Module):
    """utils/Networks.py:246-271.  Same module tree, same RNG consumption order:
    L x nn.Linear default init, then sine_init over every Linear in order, then
    first_layer_sine_init on layer 0 (`net.apply` visits children before self)."""

    def __init__(self, coords_channel=3, data_channel=1, features=256, layers=5, w0=30,
                 res=False, output_act=False, **kwargs):
        super().__init__()
        if res:
            raise NotImplementedError("res=True (HalfResidual) is outside the hot path (SURVEY 2 #23)")
        net = [nn.Sequential(nn.Linear(coords_channel, features), _Sine(w0))]
        for _ in range(layers - 2):
            net.append(nn.Sequential(nn.Linear(features, features), _Sine()))
        if output_act:
            net.append(nn.Sequential(nn.Linear(features, data_channel), _Sine()))
        else:
            net.append(nn.Sequential(nn.Linear(features, data_channel)))
        self.net = nn.Sequential(*net)
        self.net.apply(_sine_init)
        self.net[0].apply(_first_layer_sine_init)

    def forward(self, coords):
        return self.net(coords)


def calc_param_count(coords_channel, data_channel, features, layers, res=False, **kw) -> int:
    """utils/Networks.py:291-297 (res=False branch)."""
    return int(coords_channel * features + features + (layers - 2) * (features ** 2 + features)
               + features * data_channel + data_channel)


def calc_features(param_count, coords_channel, data_channel, layers, res=False, **kw) -> int:
    """utils/Networks.py:298-314 (res=False branch): positive root of a f^2 + b f + c = 0, rounded."""
    a = layers - 2
    b = coords_channel + 1 + layers - 2 + data_channel
    c = -param_count + data_channel
    if a == 0:
        return round(-c / b)
    return round((-b + math.sqrt(b ** 2 - 4 * a * c)) / (2 * a))


def init_phi(kwargs) -> OracleSIREN:
    """utils/Networks.py:800-802 (only 'SIREN' is on the hot path)."""
    kwargs = copy.deepcopy(dict(kwargs))
    name = kwargs.pop("name")
    return {"SIREN": OracleSIREN}[name](**kwargs)


def estimate_module_size(ideal_module_size: float, phi_kwargs: dict, half: bool = False):
    """main.py:214-246, SIREN branch: byte budget -> (features, param_count, theory bytes)."""
    ideal = ideal_module_size / (2.0 if half else 4.0)
    kw = {k: v for k, v in phi_kwargs.items() if k not in ("name", "features")}
    f = calc_features(param_count=ideal, **kw)
    p = calc_param_count(features=f, **kw)
    return f, p, p * (2.0 if half else 4.0)


def siren_params(module: nn.Module) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """[(W_l [out,in], b_l [out])] in layer order (the layout ModelSave writes)."""
    return [(module.net[l][0].weight.detach(), module.net[l][0].bias.detach())
            for l in range(len(module.net))]


def omegas(layers: int, w0: float) -> List[float]:
    """utils/Networks.py:249,259: first Sine(w0), hidden Sine() = 30; last layer has no activation."""
    return [float(w0)] + [HIDDEN_OMEGA] * (layers - 2)


def forward_layers(params, coords: torch.Tensor, w0: float):
    """Functional forward returning every pre-activation z_l and activation a_l."""
    om = omegas(len(params), w0)
    a = coords
    zs, acts = [], []
    for l, (W, b) in enumerate(params):
        z = F.linear(a, W, b)
        zs.append(z)
        if l < len(params) - 1:
            a = torch.sin(om[l] * z)
            acts.append(a)
        else:
            a = z
    return a, zs, acts


# --------------------------------------------------------------------------------------
# Coordinates                                                   utils/dataset.py:11-62
# --------------------------------------------------------------------------------------
def _parse_mode(mode):
    if mode == "n11":
        return -1, 1
    if mode == "0p1":
        return 0, 1
    lo, hi = mode.split(",")
    return float(lo), float(hi)


def axis_coords(n: int, mode: str = "-1,1") -> torch.Tensor:
    lo, hi = _parse_mode(mode)
    return torch.linspace(lo, hi, n)


def create_coords(coords_shape: Sequence[int], mode: str = "n11") -> torch.Tensor:
    """utils/dataset.py:11-35: meshgrid(ij) of per-axis linspace, stacked last -> [..., ndim]."""
    axes = [axis_coords(int(n), mode) for n in coords_shape]
    if len(axes) not in (2, 3):
        raise NotImplementedError
    return torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1)


def create_flattened_coords(coords_shape: Sequence[int], mode: str = "n11") -> torch.Tensor:
    """utils/dataset.py:36-62."""
    c = create_coords(coords_shape, mode)
    return c.reshape(-1, c.shape[-1])


# --------------------------------------------------------------------------------------
# Normalisation                                                utils/io.py:65-80, 111-147
# --------------------------------------------------------------------------------------
def normalize_data(data: np.ndarray, name: str, min=None, max=None):
    """utils/io.py:67-80 ('minmaxany_lo_hi' only): fp32 (x-min)/(max-min)*(hi-lo)+lo."""
    if "minmaxany" not in name:
        raise NotImplementedError
    lo, hi = [float(s) for s in name.split("_")[1:]]
    dtype = data.dtype.name
    data = data.astype(np.float32)
    if min is None:
        min = float(data.min())
    if max is None:
        max = float(data.max())
    data = (data - min) / (max - min)
    data *= (hi - lo)
    data += lo
    t = torch.tensor(data, dtype=torch.float)
    return t, {"dtype": dtype, "min": min, "max": max,
               "normalized_min": t.min().item(), "normalized_max": t.max().item()}


def invnormalize_data(data: torch.Tensor, sideinfos: dict, name: str) -> np.ndarray:
    """utils/io.py:111-147 ('minmaxany'): (x-lo)/(hi-lo) -> clip[0,1] -> *(max-min)+min -> TRUNCATING cast."""
    if "minmaxany" not in name:
        raise NotImplementedError
    lo, hi = [float(s) for s in name.split("_")[1:]]
    data = data.clone()
    data -= lo
    data /= (hi - lo)
    data = torch.clip(data, 0, 1)
    data = data * (sideinfos["max"] - sideinfos["min"]) + sideinfos["min"]
    return np.array(data, dtype=sideinfos["dtype"])


def get_type_max(data: np.ndarray) -> int:
    """utils/tool.py:8-24."""
    return {"uint8": 255, "uint12": 4098, "uint16": 65535, "float32": 65535, "float64": 65535,
            "int16": 65535}[data.dtype.name]


# --------------------------------------------------------------------------------------
# Loss weights / checkpoints / preprocess                       utils/misc.py:244-307
# --------------------------------------------------------------------------------------
def parse_weight(data: np.ndarray, weight_type_list: Iterable[str]) -> np.ndarray:
    """utils/misc.py:272-307."""
    data = np.asarray(data)
    weight = np.ones_like(data).astype(np.float32)
    tmax = get_type_max(data)
    for wt in weight_type_list:
        if "quantile" in wt:
            _, ge, ql, qh, scale = wt.split("_")
            ge, ql, qh, scale = float(ge), float(ql), float(qh), float(scale)
            l = np.quantile(data[data >= ge], ql)
            h = np.quantile(data[data >= ge], qh)
            assert 0 <= l <= h <= tmax, "Improper range setting!"
            weight[(data >= l) * (data <= h)] = scale
        elif "value" in wt:
            _, l, h, scale = wt.split("_")
            l, h, scale = float(l), float(h), float(scale)
            assert 0 <= l <= h <= tmax, "Improper range setting!"
            weight[(data >= l) * (data <= h)] = scale
        elif "exp" in wt:
            _, mid_x, mid_value = wt.split("_")
            a = -np.log(float(mid_value)) / float(mid_x)
            weight = np.exp(-a * data)
        elif wt == "none":
            pass
        else:
            raise NotImplementedError
    return weight


def parse_checkpoints(checkpoints, max_steps: int) -> List[int]:
    """utils/misc.py:255-271."""
    if checkpoints == "none":
        return [max_steps]
    if "every" in checkpoints:  # NB: an int reaches this line first and raises TypeError, as in the reference
        interval = int(checkpoints.split("_")[1])
        out = list(range(interval, max_steps, interval))
        out.append(max_steps)
        return out
    if isinstance(checkpoints, int):  # unreachable in the reference too (utils/misc.py:263-267)
        return [max_steps] if checkpoints >= max_steps else [checkpoints, max_steps]
    out = [int(s) for s in checkpoints.split(",") if int(s) < max_steps]
    out.append(max_steps)
    return out


def preprocess(data: np.ndarray, denoise_level: int, denoise_close, clip_range):
    """utils/misc.py:244-254."""
    from scipy import ndimage
    if denoise_close is False:
        data[data <= denoise_level] = 0
    else:
        k = tuple(list(denoise_close)[: data.ndim - 1] + [1]) if data.ndim == 3 else tuple(list(denoise_close) + [1])
        data[ndimage.binary_opening(data <= denoise_level, structure=np.ones(k), iterations=1)] = 0
    l, h = clip_range
    assert 0 <= l <= h <= get_type_max(data)
    return data.clip(l, h)


def box_opening(mask: np.ndarray, size) -> np.ndarray:
    """What ndimage.binary_opening(mask, structure=np.ones(size), iterations=1) computes, written out: the erosion
    E[q] = AND of the mask over the box anchored at q (boxes that leave the array are 0: border_value = 0), the opening
    O[p] = OR of E over all boxes containing p.  Pinned against scipy in oracle/gen_golden_preprocess.py and in
    tests/test_oracle_golden.py; the device kernel (brief_preprocess.cu) is this on one bit per voxel."""
    mask = np.asarray(mask, dtype=bool)
    size = tuple(int(s) for s in size)
    assert mask.ndim == len(size)
    er = mask.copy()
    for ax, s in enumerate(size):      # separable AND over the box, anchored at the low corner
        acc = er.copy()
        for k in range(1, s):
            sh = np.zeros_like(er)
            sl_dst = [slice(None)] * er.ndim
            sl_src = [slice(None)] * er.ndim
            sl_dst[ax], sl_src[ax] = slice(0, max(er.shape[ax] - k, 0)), slice(k, None)
            sh[tuple(sl_dst)] = er[tuple(sl_src)]
            acc &= sh                 # positions whose box leaves the array read 0
        er = acc
    op = er.copy()
    for ax, s in enumerate(size):      # separable OR over the anchors whose box contains p
        acc = op.copy()
        for k in range(1, s):
            sh = np.zeros_like(op)
            sl_dst = [slice(None)] * op.ndim
            sl_src = [slice(None)] * op.ndim
            sl_dst[ax], sl_src[ax] = slice(k, None), slice(0, max(op.shape[ax] - k, 0))
            sh[tuple(sl_dst)] = op[tuple(sl_src)]
            acc |= sh
        op = acc
    return op


def preprocess_restated(data: np.ndarray, denoise_level, denoise_close, clip_range) -> np.ndarray:
    """preprocess without scipy (box_opening above); does not touch its input."""
    data = np.array(data, copy=True)
    m = data <= denoise_level
    if denoise_close is not False:
        k = tuple(list(denoise_close)[:2] + [1]) if data.ndim == 3 else tuple(list(denoise_close) + [1])
        m = box_opening(m, k)
    data[m] = 0
    l, h = clip_range
    assert 0 <= l <= h <= get_type_max(data)
    return data.clip(l, h)


# --------------------------------------------------------------------------------------
# Samplers                                   main.py:38-163 (== utils/sampler.py:9-94)
# --------------------------------------------------------------------------------------
class RandompointSampler:
    """main.py:126-163: B indices with replacement from torch's CPU generator, three gathers."""

    def __init__(self, data: torch.Tensor, weight: np.ndarray, coords_mode: str, sample_size: int,
                 sample_count: int, device: str = "cpu"):
        self.sample_size, self.sample_count = sample_size, sample_count
        shape = tuple(data.shape[:-1])
        if len(shape) not in (2, 3):
            raise NotImplementedError
        self.coords = create_flattened_coords(shape, mode=coords_mode).to(device)
        self.data = data.reshape(-1, data.shape[-1])
        w = torch.from_numpy(weight).to(device)
        self.weight = w.reshape(-1, w.shape[-1])
        self.pop_size = int(np.prod(shape))
        self.last_idx = None

    def __len__(self):
        return self.sample_count

    def __iter__(self):
        self.index = 0
        return self

    def __next__(self):
        if self.index >= len(self):
            raise StopIteration
        idx = torch.randint(0, self.pop_size, (self.sample_size,))
        self.last_idx = idx
        self.index += 1
        return self.coords[idx, :], self.data[idx, :], self.weight[idx, :]


class RandomCubeSampler:
    """main.py:38-125 (3-D branch :42-71, 2-D branch :73-102): cube_len is clamped to the block, the population is
    every stride-1 window position listed '(dc hc wc)' / '(hc wc)', and a step is `cube_count` draws with replacement.
    With the shipped cube_len (clamped to the block) and cube_count=1 the population is one cube == the whole block."""

    def __init__(self, data: torch.Tensor, weight: np.ndarray, coords_mode: str, cube_count: int,
                 cube_len: List[int], sample_count: int, device: str = "cpu", gpu_force: bool = False):
        nd = data.dim() - 1
        if nd not in (2, 3):
            raise NotImplementedError
        self.sample_count = sample_count
        cube_len = [min(int(cube_len[i]), data.shape[i]) for i in range(nd)]
        self.cube_len = cube_len
        coords = create_coords(tuple(data.shape[:nd]), mode=coords_mode)
        wt = torch.from_numpy(weight)

        def cubes(t):
            u = t
            for k in range(nd):
                u = u.unfold(k, cube_len[k], 1)
            # [dc,hc,wc,c,ds,hs,ws] -> [(dc hc wc), ds, hs, ws, c]   (2-D: [hc,wc,c,hs,ws] -> [(hc wc), hs, ws, c])
            u = u.permute(*range(nd), *range(nd + 1, 2 * nd + 1), nd)
            return u.reshape(-1, *cube_len, t.shape[-1])

        self.coords_cubes, self.data_cubes, self.weight_cubes = cubes(coords), cubes(data), cubes(wt)
        self.pop_size = self.data_cubes.shape[0]
        self.cube_count = cube_count
        self.last_idx = None

    def __len__(self):
        return self.sample_count

    def __iter__(self):
        self.index = 0
        return self

    def __next__(self):
        if self.index >= len(self):
            raise StopIteration
        idx = torch.randint(0, self.pop_size, (self.cube_count,))
        self.last_idx = idx
        self.index += 1
        return self.coords_cubes[idx], self.data_cubes[idx], self.weight_cubes[idx]


def cube_voxel_indices(shape: Sequence[int], cube_len: Sequence[int], cube_ids) -> np.ndarray:
    """Flat voxel indices [n_cubes, prod(cube_len)] of windows `cube_ids` of RandomCubeSampler's population
    (main.py:61-69: unfold along every axis with stride 1, cubes listed row-major over their origins, voxels of a cube
    row-major).  This is what brief_common.cuh's brief_cube_voxel restates on the device."""
    shape = [int(n) for n in shape]
    clen = [min(int(c), n) for c, n in zip(cube_len, shape)]
    counts = [n - c + 1 for n, c in zip(shape, clen)]
    origin = np.stack(np.unravel_index(np.asarray(cube_ids, dtype=np.int64), counts), axis=-1)      # [n, nd]
    offs = np.stack(np.unravel_index(np.arange(int(np.prod(clen)), dtype=np.int64), clen), axis=-1)  # [v, nd]
    pos = origin[:, None, :] + offs[None, :, :]
    return np.ravel_multi_index(tuple(pos[..., k] for k in range(len(shape))), shape).astype(np.int64)


# --------------------------------------------------------------------------------------
# Loss / optimiser / schedule / train step   main.py:171-197, 385-400; utils/misc.py:174-197
# --------------------------------------------------------------------------------------
def datal2(data_gt, data_hat, weight, weight_thres):
    """main.py:176-182.  NB: mutates `weight` in place, exactly like the reference."""
    loss = F.mse_loss(data_hat, data_gt, reduction="none")
    if weight_thres:
        weight[data_hat <= weight_thres] = 1
    loss = loss * weight
    return loss.mean()


def configure_optimizer(parameters, optimizer: str, lr: float):
    """utils/misc.py:174-183."""
    if optimizer == "Adam":
        return torch.optim.Adam(parameters, lr=lr)
    if optimizer == "Adamax":
        return torch.optim.Adamax(parameters, lr=lr)
    if optimizer == "SGD":
        return torch.optim.SGD(parameters, lr=lr)
    raise NotImplementedError


def configure_lr_scheduler(optimizer, lr_scheduler_opt: dict):
    """utils/misc.py:184-197 (MultiStepLR / CyclicLR / StepLR / none)."""
    opt = copy.deepcopy(dict(lr_scheduler_opt))
    name = opt.pop("name")
    if name == "MultiStepLR":
        return torch.optim.lr_scheduler.MultiStepLR(optimizer, **opt)
    if name == "StepLR":
        return torch.optim.lr_scheduler.StepLR(optimizer, **opt)
    if name == "CyclicLR":
        return torch.optim.lr_scheduler.CyclicLR(optimizer, **opt)
    if name == "none":
        return torch.optim.lr_scheduler.MultiStepLR(optimizer, milestones=[100000000000])
    raise NotImplementedError


def train_step(module, optimizer, scheduler, coords, data, weight, weight_thres) -> torch.Tensor:
    """main.py:385-400 (half=False): zero_grad -> forward -> datal2 -> backward -> step -> sched.step."""
    optimizer.zero_grad()
    data_hat = module.forward(coords)
    loss = datal2(data, data_hat, weight, weight_thres)
    loss.backward()
    optimizer.step()
    scheduler.step()
    return loss


def loss_and_grads(params, coords, data, weight, weight_thres, w0):
    """fwd + datal2 + autograd backward for explicit parameter tensors (per-layer grad oracle)."""
    ps = [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in params]
    yhat, zs, acts = forward_layers(ps, coords, w0)
    loss = datal2(data, yhat, weight.clone(), weight_thres)
    flat = [t for Wb in ps for t in Wb]
    grads = torch.autograd.grad(loss, flat)
    return loss.detach(), yhat.detach(), [(grads[2 * i], grads[2 * i + 1]) for i in range(len(ps))], \
        [z.detach() for z in zs]


def weight_thres_normalized(weight_thres: float, normalize_name: str, vmin: float, vmax: float) -> float:
    """main.py:380-383: the raw threshold pushed through normalize_data with the block's min/max."""
    t, _ = normalize_data(np.array(weight_thres), normalize_name, min=vmin, max=vmax)
    return float(t)


# --------------------------------------------------------------------------------------
# Decompress                                  utils/misc.py:59-92, main.py:270-297
# --------------------------------------------------------------------------------------
def reconstruct_flattened(data_shape, sample_size: int, sample_nf: Callable, device="cpu",
                          half=False, coords_mode="-1,1") -> torch.Tensor:
    """utils/misc.py:59-92: chunked no-grad evaluation over the dense grid."""
    *cshape, ch = data_shape
    with torch.no_grad():
        coords = create_flattened_coords(tuple(cshape), coords_mode).to(device)
        pop = coords.shape[0]
        flat = torch.zeros((pop, ch), device=device)
        for i in range(math.ceil(pop / sample_size)):
            s, e = i * sample_size, min((i + 1) * sample_size, pop)
            flat[s:e, :] = sample_nf(coords[s:e, :])
    return flat.reshape(*cshape, ch)


def decompress_block(module, sideinfos: dict, normalize_name: str, sample_size=10000,
                     coords_mode="-1,1") -> np.ndarray:
    """main.py:270-297 minus file IO: evaluate, inverse-normalise (truncating), postprocess is a
    no-op for unsigned dtypes (utils/misc.py:244-254 with level 0 / clip [0,max])."""
    out = reconstruct_flattened(sideinfos["data_shape"], sample_size, module.forward,
                                coords_mode=coords_mode).float().cpu()
    return invnormalize_data(out, sideinfos, normalize_name)


# --------------------------------------------------------------------------------------
# Compressed-parameter layout                                  utils/ModelSave.py:8-51
# --------------------------------------------------------------------------------------
def save_model(model, save_path: str) -> None:
    """utils/ModelSave.py:32-51: rmtree + mkdir; 'weight-l-out-in' / 'bias-l-n' raw native fp32."""
    if os.path.exists(save_path):
        shutil.rmtree(save_path)
    os.mkdir(save_path)
    for l in range(len(model.net)):
        w = np.array(model.net[l][0].weight.data.to("cpu"))
        b = np.array(model.net[l][0].bias.data.to("cpu"))
        with open(os.path.join(save_path, f"weight-{l}-{w.shape[0]}-{w.shape[1]}"), "wb") as f:
            flat = w.reshape(-1)
            f.write(struct.pack("f" * len(flat), *flat))
        with open(os.path.join(save_path, f"bias-{l}-{len(b)}"), "wb") as f:
            f.write(struct.pack("f" * len(b), *b))


def load_model(model, model_path: str):
    """utils/ModelSave.py:8-30: shapes parsed from file names."""
    for file in os.listdir(model_path):
        with open(os.path.join(model_path, file), "rb") as fh:
            raw = fh.read()
        if "weight" in file:
            _, l, s0, s1 = file.split("-")
            l, s0, s1 = int(l), int(s0), int(s1)
            w = np.array(struct.unpack("f" * s0 * s1, raw)).astype(np.float32).reshape(s0, s1)
            model.net[l][0].weight.data = torch.tensor(w)
        elif "bias" in file:
            _, l, n = file.split("-")
            b = np.array(struct.unpack("f" * int(n), raw)).astype(np.float32)
            model.net[int(l)][0].bias.data = torch.tensor(b)
    return model


# --------------------------------------------------------------------------------------
# Block partition ("next" row f-1)   utils/misc.py:329-445, utils/adaptive_blocking.py:425-460
# --------------------------------------------------------------------------------------
def cal_factor(n: int) -> List[int]:
    """utils/adaptive_blocking.py:425-430 (proper divisors only: n itself is excluded)."""
    return [1] + [i for i in range(2, n) if n % i == 0]


def cal_divide_num(d, h, w, Nb, param_size):
    """utils/adaptive_blocking.py:432-460: factor triple with the largest product <= Nb, most cubic."""
    if Nb <= 0:
        Nb = int(param_size / (4 * 1361))
        if Nb <= 0:
            Nb = 1
    num_max, number, var_min = 0, None, None
    for nd in cal_factor(d):
        for nh in cal_factor(h):
            for nw in cal_factor(w):
                num = nd * nh * nw
                if num > Nb:
                    continue
                size = np.array([d / nd, h / nh, w / nw])
                var = ((size - size.mean()) ** 2).mean()
                if num > num_max:
                    num_max, number, var_min = num, np.array([nd, nh, nw]), var
                elif num == num_max and var < var_min:
                    number, var_min = np.array([nd, nh, nw]), var
    return number


def divide_data(data: np.ndarray, divide_type: str) -> List[dict]:
    """utils/misc.py:329-366, 3-D branch (the divide_img preview is not part of the hot path)."""
    if data.ndim != 4:
        raise NotImplementedError
    if "total" in divide_type:
        nd, nh, nw = [int(s) for s in divide_type.split("_")[1:]]
        cd, ch, cw = int(data.shape[0] / nd), int(data.shape[1] / nh), int(data.shape[2] / nw)
    elif "every" in divide_type:
        cd, ch, cw = [int(s) for s in divide_type.split("_")[1:]]
    else:
        raise NotImplementedError
    ds = [i for i in range(data.shape[0]) if i % cd == 0] + [data.shape[0]]
    hs = [i for i in range(data.shape[1]) if i % ch == 0] + [data.shape[1]]
    ws = [i for i in range(data.shape[2]) if i % cw == 0] + [data.shape[2]]
    out = []
    for di in range(len(ds) - 1):
        for hi in range(len(hs) - 1):
            for wi in range(len(ws) - 1):
                blk = data[ds[di]:ds[di + 1], hs[hi]:hs[hi + 1], ws[wi]:ws[wi + 1]]
                c = {"data": blk, "d": [ds[di], ds[di + 1] - 1], "h": [hs[hi], hs[hi + 1] - 1],
                     "w": [ws[wi], ws[wi + 1] - 1], "total_size": data.size, "size": blk.size}
                c["name"] = "d_{}_{}-h_{}_{}-w_{}_{}".format(*c["d"], *c["h"], *c["w"])
                out.append(c)
    return out


def alloc_param(chunks: List[dict], param_size: float, param_alloc: str, thres: float) -> List[dict]:
    """utils/misc.py:395-428 (equal / by_size / by_var; by_d, by_dv need the FFT feature)."""
    if param_alloc == "equal":
        for c in chunks:
            c["param_size"] = param_size / len(chunks)
    elif param_alloc == "by_size":
        for c in chunks:
            c["param_size"] = param_size * c["size"] / c["total_size"]
    elif param_alloc == "by_var":
        var_total = 0
        for c in chunks:
            var_total += ((c["data"] - c["data"].mean()) ** 2).mean()
        for c in chunks:
            c["param_size"] = float(param_size * ((c["data"] - c["data"].mean()) ** 2).mean() / var_total)
    else:
        raise NotImplementedError
    kept = [c for c in chunks if c["param_size"] >= thres]
    if len(kept) < len(chunks):
        return alloc_param(kept, param_size, param_alloc, thres)
    return kept


def merge_divided_data(chunks: List[dict], data_shape) -> np.ndarray:
    """utils/misc.py:430-445: zeros fp32 canvas, += block, clip to dtype max, cast."""
    tmax = get_type_max(chunks[0]["data"])
    out = np.zeros(data_shape, dtype=np.float32)
    for c in chunks:
        (d0, d1), (h0, h1), (w0, w1) = c["d"], c["h"], c["w"]
        out[d0:d1 + 1, h0:h1 + 1, w0:w1 + 1] += c["data"]
    return out.clip(None, tmax).astype(chunks[0]["data"].dtype)


# --------------------------------------------------------------------------------------
# Quality metrics ("next" row f-2)              utils/misc.py:447-475, utils/ssim.py
# --------------------------------------------------------------------------------------
def cal_psnr(origin: np.ndarray, decompressed: np.ndarray, data_range) -> float:
    """utils/misc.py:451-456."""
    mse = np.mean(np.power(origin / data_range - decompressed / data_range, 2))
    return float(-10 * np.log10(mse))


def _gauss_1d(size=11, sigma=1.5):
    coords = torch.arange(size, dtype=torch.float) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return (g / g.sum())


def ssim_2d(x: torch.Tensor, y: torch.Tensor, data_range, K=(0.01, 0.03)) -> torch.Tensor:
    """utils/ssim.py `ssim` defaults (11-tap sigma-1.5 separable Gaussian, valid conv, mean)."""
    g = _gauss_1d().to(x.dtype)
    c = x.shape[1]

    def filt(t):
        t = F.conv2d(t, g.view(1, 1, -1, 1).repeat(c, 1, 1, 1), groups=c)
        return F.conv2d(t, g.view(1, 1, 1, -1).repeat(c, 1, 1, 1), groups=c)

    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mu1, mu2 = filt(x), filt(y)
    mu1_sq, mu2_sq, mu12 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = filt(x * x) - mu1_sq
    s2 = filt(y * y) - mu2_sq
    s12 = filt(x * y) - mu12
    cs = (2 * s12 + C2) / (s1 + s2 + C2)
    ssim_map = ((2 * mu12 + C1) / (mu1_sq + mu2_sq + C1)) * cs
    return torch.flatten(ssim_map, 2).mean(-1).mean()


def cal_ssim(origin: np.ndarray, decompressed: np.ndarray, data_range) -> float:
    """utils/misc.py:458-475: per-depth-slice 2-D SSIM averaged over depth (4-D input [d,h,w,c])."""
    o, r = torch.from_numpy(origin), torch.from_numpy(decompressed)
    if o.dim() == 3:
        return float(ssim_2d(o.permute(2, 0, 1)[None], r.permute(2, 0, 1)[None], data_range))
    tot = 0.0
    for i in range(o.shape[0]):
        tot += ssim_2d(o[i].permute(2, 0, 1)[None], r[i].permute(2, 0, 1)[None], data_range)
    return float(tot / o.shape[0])


# --------------------------------------------------------------------------------------
# Device sampler RNG (ours, not the reference's): Philox4x32-10, restated for index parity
# --------------------------------------------------------------------------------------
_PH_M0, _PH_M1 = 0xD2511F53, 0xCD9E8D57
_PH_W0, _PH_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(ctr: np.ndarray, key: Tuple[int, int]) -> np.ndarray:
    """Philox4x32-10 (Salmon et al. 2011).  ctr: uint32 [n,4]; key: two uint32.  Returns uint32 [n,4]."""
    c = ctr.astype(np.uint64).copy()
    k0, k1 = np.uint64(key[0] & 0xFFFFFFFF), np.uint64(key[1] & 0xFFFFFFFF)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(_PH_M0) * c[:, 0]
        p1 = np.uint64(_PH_M1) * c[:, 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        n0 = (hi1 ^ c[:, 1] ^ k0) & mask
        n1 = lo1
        n2 = (hi0 ^ c[:, 3] ^ k1) & mask
        n3 = lo0
        c = np.stack([n0, n1, n2, n3], axis=1)
        k0 = (k0 + np.uint64(_PH_W0)) & mask
        k1 = (k1 + np.uint64(_PH_W1)) & mask
    return c.astype(np.uint32)


def device_sample_indices(seed: int, step: int, net: int, batch: int, pop_size: int) -> np.ndarray:
    """The index stream of brief_b200's on-device sampler (csrc/sampler.cuh): sample s of network
    `net` at step `step` uses Philox counter (s>>2, step_lo, step_hi, net), key = seed lo/hi, word
    s&3, mapped to [0,pop) by the 32x32->hi multiply (with replacement, like torch.randint)."""
    s = np.arange(batch, dtype=np.uint64)
    ctr = np.stack([(s >> np.uint64(2)).astype(np.uint32),
                    np.full(batch, step & 0xFFFFFFFF, np.uint32),
                    np.full(batch, (step >> 32) & 0xFFFFFFFF, np.uint32),
                    np.full(batch, net & 0xFFFFFFFF, np.uint32)], axis=1)
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    word = r[np.arange(batch), (s & np.uint64(3)).astype(np.int64)].astype(np.uint64)
    return ((word * np.uint64(pop_size)) >> np.uint64(32)).astype(np.int64)
