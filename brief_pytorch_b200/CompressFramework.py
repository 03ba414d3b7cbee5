"""NFGR — the hot-path half of the reference's compression framework (class NFGR, main.py:164-651), re-built around
grouped kernels: the blocks of a volume are independent networks (main.py:484-532), so instead of one OS process
per block (main.py:547-579, utils/TasksManager.py) every block owned by this rank is fitted in ONE SirenGroup —
one fused fit launch + one optimiser launch per step for all of them — and decoded by one decompress launch.

Kept from the reference: the config tree (the `CompressFramework` sub-tree of opt/*.yaml, as a plain dict), the block
partition / budget allocation (`divide`, utils/misc.py), the width solver, the network constructor and its CPU-RNG
initialisation order, the optimiser / scheduler / checkpoint semantics, and the on-disk layout
`compressed/{sideinfos.yaml, module/<chunk>/module/<weight-l-o-i | bias-l-n>, sideinfos/<chunk>/sideinfos.yaml}`
(main.py:589-607, utils/ModelSave.py) — directories written here decode with the reference and vice versa.
Not here (out of the hot path): TIFF / video IO, logging, MIP images, the Gurobi partition.
"""
from __future__ import annotations

import copy
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import yaml

from . import misc, sharding
from .group import NetSpec, SirenGroup, block_stats, pack_module_params, preprocess_, unpack_module_params
from .io import get_type_max, normalized_threshold
from .ModelSave import load_model, save_model
from .Networks import ALL_CALC_PHI_FEATURES, ALL_CALC_PHI_PARAM_COUNT, get_nnmodule_param_count, init_phi

_DTYPES = {"uint8": np.uint8, "uint16": np.uint16, "float32": np.float32}


def _norm_range(name: str):
    if "minmaxany" not in name:
        raise NotImplementedError(f"normalisation '{name}' is outside the SIREN hot path")
    _, lo, hi = name.split("_")
    return float(lo), float(hi)


def _f32(x) -> float:
    return float(np.float32(x))


@dataclass
class Block:
    """One block of the partition and everything the fit produces for it."""
    name: str
    data: Optional[np.ndarray]             # [d,h,w,1] view of the (pre-processed) volume, original dtype (None: see dev)
    d: List[int]
    h: List[int]
    w: List[int]
    param_size: float                      # byte budget (alloc_param)
    features: int = 0
    module: Optional[torch.nn.Module] = None
    sideinfos: Dict = field(default_factory=dict)
    loss: float = float("nan")
    dev: Optional[torch.Tensor] = None     # the same voxels already on the device, contiguous [d,h,w] (uint16 as int16
                                           # bit patterns): the fit then never touches `data` for the voxel values
    np_dtype: str = ""                     # dtype name of the raw voxels when `data` is None

    @property
    def shape(self):
        if self.data is not None:
            return tuple(int(x) for x in self.data.shape[:3])
        return (self.d[1] - self.d[0] + 1, self.h[1] - self.h[0] + 1, self.w[1] - self.w[0] + 1)

    @property
    def dtype_name(self) -> str:
        return self.data.dtype.name if self.data is not None else self.np_dtype


def _merge(base: dict, over: dict) -> dict:
    """OmegaConf.merge for plain dicts (main.py:568-569): nested keys of `over` replace those of `base`."""
    out = copy.deepcopy(base)
    for k, v in over.items():
        out[k] = _merge(out[k], v) if isinstance(v, dict) and isinstance(out.get(k), dict) else copy.deepcopy(v)
    return out


def _to_device_raw(a: np.ndarray, device) -> torch.Tensor:
    a = np.ascontiguousarray(a)
    return torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a).to(device)


class NFGR:
    def __init__(self, opt: dict, device: int | str = 0, precision: str = "auto", reproducible: bool = False):
        """reproducible=True: a block's fitted parameters are bit-identical whichever rank owns it and whatever else
        shares its GPU (per-network slicing, SirenGroup.set_slicing); the default slicing is faster and reproducible
        run to run for a fixed sharding, but may differ in the last bits between shardings."""
        self.opt = copy.deepcopy(opt)
        self.device = device
        self.precision = precision
        self.reproducible = reproducible
        # Compress.half (main.py:388-398) is fp16 forward / backward around an fp32 optimiser step — the arithmetic of the
        # f16 tensor-core mode (fp16 operands, fp32 master weights; here with fp32 accumulation and fp32 coordinates).
        # What `half` changes on this path is the width rule: 2 bytes per parameter (main.py:217-220, 242-245).
        self.half = bool(self.opt["Compress"].get("half", False))
        if self.opt["Compress"]["loss"]["name"] != "datal2":
            raise NotImplementedError(self.opt["Compress"]["loss"]["name"])  # main.py:197
        if self.opt["Module"]["phi"]["name"] not in ALL_CALC_PHI_FEATURES:
            raise KeyError(self.opt["Module"]["phi"]["name"])

    # ---- byte budget -> width (main.py:199-246) -----------------------------------------------------------------
    def parse_param_size(self, orig_bytes: int) -> float:
        p = self.opt["Compress"]["param"]
        given, ratio = p.get("given_size", 0), p.get("filesize_ratio", 0)
        if given > 0 and ratio == 0:
            return float(given)
        if ratio > 0 and given == 0:
            return orig_bytes / ratio
        raise ValueError("exactly one of param.given_size / param.filesize_ratio must be set")

    def estimate_module_size(self, ideal_bytes: float):
        phi = {k: v for k, v in self.opt["Module"]["phi"].items() if k not in ("name", "features")}
        name = self.opt["Module"]["phi"]["name"]
        per = 2.0 if self.half else 4.0
        f = ALL_CALC_PHI_FEATURES[name](param_count=ideal_bytes / per, **phi)
        return f, ALL_CALC_PHI_PARAM_COUNT[name](features=f, **phi) * per

    # ---- partition (main.py:484-532) ------------------------------------------------------------------------------
    def divide(self, data: np.ndarray, param_size: float, dev_volume: Optional[torch.Tensor] = None) -> List[Block]:
        """Partition + budget allocation.  `dev_volume` (optional): the same (pre-processed) volume on the device,
        [d,h,w] in the raw dtype — the per-block variances of `by_var` then come from ONE brief_block_stats launch over
        the device blocks (which are kept in Block.dev for the fit) instead of two numpy passes per chunk.
        self.chunks_total = number of chunks BEFORE the param_size_thres filter (the reference's chunks_numbers)."""
        dv = self.opt["Compress"]["divide"]
        kind = dv["divide_type"]
        if kind == "none":
            d, h, w = data.shape[:3]
            chunks = [{"data": data, "d": [0, d - 1], "h": [0, h - 1], "w": [0, w - 1], "size": data.size,
                       "total_size": data.size, "name": f"d_0_{d - 1}-h_0_{h - 1}-w_0_{w - 1}", "param_size": param_size}]
            self.chunks_total = 1
        else:
            if kind.startswith("adaptive"):
                nb = int(kind.split("_")[-1])
                if nb >= 8:
                    raise NotImplementedError("the Gurobi oct-tree partition is outside the hot path; use adaptotal_*")
                kind = f"adaptotal_-1_-1_-1_{nb}"
            if kind.startswith("adaptotal"):
                _, nd, nh, nw, nb = kind.split("_")
                nd, nh, nw, nb = int(nd), int(nh), int(nw), int(nb)
                if -1 in (nd, nh, nw):
                    nd, nh, nw = misc.cal_divide_num(*data.shape[:3], nb, param_size)
                kind = f"total_{nd}_{nh}_{nw}"
            chunks, _ = misc.divide_data(data, kind)
            self.chunks_total = len(chunks)
            if dev_volume is not None:
                for c in chunks:
                    c["dev"] = dev_volume[c["d"][0]:c["d"][1] + 1, c["h"][0]:c["h"][1] + 1, c["w"][0]:c["w"][1] + 1].contiguous()
                if dv["param_alloc"] == "by_var":
                    st = block_stats([c["dev"] for c in chunks], data.dtype.name)
                    for c, row in zip(chunks, st):
                        c["var"] = misc.variance_from_sums(row[2], row[3], c["size"])
            chunks = misc.alloc_param(chunks, param_size, dv["param_alloc"], dv["param_size_thres"])
        return [Block(c["name"], c["data"], c["d"], c["h"], c["w"], float(c["param_size"]), dev=c.get("dev")) for c in chunks]

    # ---- fit (main.py:322-454 for every block at once) -----------------------------------------------------------
    def _sampler_name(self, n_vox: int, shape) -> str:
        """main.py:325-334: the cube sampler survives only while min(block voxels, configured cube voxels) <= 80^3."""
        sp = self.opt["Compress"]["sampler"]
        name = sp["name"]
        if name not in ("randomcube", "randompoint"):
            raise NotImplementedError(name)  # main.py:371
        if name == "randomcube":
            cube_len = [int(c) for c in sp.get("cube_len", [10000000] * 3)]
            cube = cube_len[0] * cube_len[1] * cube_len[2] if len(shape) == 3 else cube_len[1] * cube_len[2]
            if min(n_vox, cube) > 80 ** 3:
                return "randompoint"
        return name

    def _cube_config(self, shape):
        """(cube_count, cube_len clamped to the block) of the cube sampler (main.py:49-50, 367-369)."""
        sp = self.opt["Compress"]["sampler"]
        cube_len = [int(c) for c in sp.get("cube_len", [10000000] * 3)]
        return int(sp.get("cube_count", 1)), [min(c, int(n)) for c, n in zip(cube_len, shape)]

    def _step_batch(self, shape) -> int:
        """Samples one step draws from a block of this shape."""
        n_vox = int(np.prod(shape))
        if self._sampler_name(n_vox, shape) == "randompoint":
            return int(self.opt["Compress"]["sampler"]["sample_size"])
        count, clen = self._cube_config(shape)
        return count * int(np.prod(clen))

    def fit_blocks(self, blocks: Sequence[Block], max_steps: Optional[int] = None, seed: int = 42,
                   sampler_generator: str = "device", on_checkpoint=None,
                   stream_ids: Optional[Sequence[int]] = None, preprocessed: bool = False) -> SirenGroup:
        """Fit every block's network (grouped launches).  Initial weights are drawn block by block from torch's CPU
        generator in the reference's order (seed -> init_phi).  Returns the live group (parameters on the device);
        block.module / block.sideinfos / block.loss are filled in.
        preprocessed=True: the blocks were cut from a volume that already went through Compress.preprocess as a WHOLE
        (compress_divide, main.py:518-531, whose sub-tasks run with denoise level 0 / close False, main.py:557-558)."""
        C = self.opt["Compress"]
        lo, hi = _norm_range(self.opt["Normalize"]["name"])
        max_steps = int(C["max_steps"] if max_steps is None else max_steps)
        checkpoints = misc.parse_checkpoints(C.get("checkpoints", "none"), max_steps)
        specs = []
        for b in blocks:
            if not b.features:
                b.features, _ = self.estimate_module_size(b.param_size)
            kw = dict(self.opt["Module"]["phi"], features=b.features)
            torch.manual_seed(seed)   # every block's own process re-seeds in the reference (main.py:653-661)
            b.module = init_phi(kw)
            want = ALL_CALC_PHI_PARAM_COUNT[kw["name"]](**{k: v for k, v in kw.items() if k != "name"})
            assert get_nnmodule_param_count(b.module) == want  # main.py:261-262
            init_path = str(C["param"].get("init_net_path", "none"))
            if init_path != "none":  # warm start (main.py:349-354): weights only, no optimiser state
                if len(blocks) == 1:
                    load_model(b.module, init_path)
                elif os.path.isdir(os.path.join(init_path, b.name, "module")):  # a compressed/module tree of a divided run
                    load_model(b.module, os.path.join(init_path, b.name, "module"))
                else:
                    raise FileNotFoundError(f"init_net_path has no module for block {b.name}")
            specs.append(NetSpec(b.features, kw["layers"], kw["w0"], b.shape, kw["coords_channel"], kw["data_channel"]))
        grp = SirenGroup(specs, self.device, self.precision)
        grp.set_slicing(self.reproducible)
        if stream_ids is not None:  # global block indices: a block draws the same samples whichever rank owns it
            for i, sid in enumerate(stream_ids):
                grp.set_stream(i, sid)
        self._keep = []
        # raw blocks go to the device in their own dtype; min / max of every block in ONE launch (brief_block_stats)
        # instead of normalize_data's host numpy passes (utils/io.py:67-80)
        dtypes = {b.dtype_name for b in blocks}
        if len(dtypes) != 1:
            raise NotImplementedError("blocks of one volume share a dtype")
        dtype = dtypes.pop()
        np_dt = np.dtype(dtype)
        dev_raw = [b.dev if b.dev is not None else _to_device_raw(b.data[..., 0], grp.device) for b in blocks]
        weight_rules = list(C["loss"]["weight"])
        pre = C.get("preprocess")
        if pre and not preprocessed:
            if np_dt in (np.uint8, np.uint16):
                if not misc.preprocess_is_identity(np_dt, pre["denoise"]["level"], pre["clip"]):
                    # main.py:336-337 (single-task path): threshold + opening + clip on the device copy (brief_preprocess);
                    # min / max and the weights below are those of the preprocessed block
                    misc._limits(np.zeros(0, np_dt), pre["clip"][0], pre["clip"][1])
                    for t in dev_raw:
                        preprocess_(t, pre["denoise"]["level"], pre["denoise"]["close"], pre["clip"], dtype)
            else:
                misc.check_float_preprocess_is_identity(blocks, pre)  # raises NotImplementedError otherwise
        stats = block_stats(dev_raw, dtype)
        for i, b in enumerate(blocks):
            grp.set_axes(i, str(C["coords_mode"]))
            grp.load_module(i, b.module)
            vmin, vmax = float(stats[i, 0]), float(stats[i, 1])
            b.sideinfos = {"dtype": dtype, "min": vmin, "max": vmax, "data_shape": list(b.shape) + [1],
                           "phi_features": int(b.features), "phi_name": self.opt["Module"]["phi"]["name"]}
            t = dev_raw[i]
            rules = misc.weight_rules_for_kernel(t, weight_rules, np_dt)
            weight = None
            if rules is None:  # 'exp' rule or more than 4 rules: explicit per-voxel weights
                raw = t.cpu().numpy()
                raw = raw.view(np.uint16) if dtype == "uint16" else raw
                weight = torch.from_numpy(np.ascontiguousarray(misc.parse_weight(raw, weight_rules), np.float32).reshape(-1)).to(grp.device)
            thres = C["loss"].get("weight_thres", 0)
            assert thres <= get_type_max(np.zeros(0, np_dt))  # main.py:380
            # a constant block (max == min: an all-background block of a sparse volume) normalises to 0/0 in the
            # reference (NaN weights for ever, warnings only); here it is fitted against the constant target `lo`
            # (span 1) and decodes exactly to its constant because the inverse multiplies by (max - min) = 0
            flat = vmax == vmin
            fit_max = vmin + 1.0 if flat else vmax
            # main.py:380-383 + 178: the override is gated on the NORMALISED threshold's truthiness
            tau = normalized_threshold(thres, self.opt["Normalize"]["name"], vmin, fit_max)
            grp.bind_volume(i, t, vmin, fit_max, lo, hi, weight=weight, rules=rules or (), tau=tau, np_dtype=dtype)
            if flat:
                grp.set_denorm(i, vmin, vmax, lo, hi)
            self._keep.append((t, weight))
            if self._sampler_name(int(np.prod(b.shape)), b.shape) == "randomcube":
                # windows of cube_len slid over the block; the shipped cube_len >= block, cube_count 1 is the whole block
                grp.set_cube_sampler(i, *self._cube_config(b.shape))
            else:
                grp.set_sampler(i, "randompoint", int(C["sampler"]["sample_size"]))
        opt = misc.configure_lr_scheduler(misc.configure_optimizer(None, C["optimizer_name_phi"], C["lr_phi"]),
                                          C["lr_scheduler_phi"])
        done = 0
        for ck in checkpoints:
            if opt.cyclic is not None:  # CyclicLR: lr and beta1 of every step from torch's scheduler, one enqueue per step
                hist = None
                for lr_t, b1_t in opt.per_step(done + 1, ck - done):
                    hist = grp.fit_run(1, opt.name, lr_t, (b1_t, opt.betas[1]), opt.eps, seed=seed, loss_history=True)
            else:
                hist = grp.fit_run(ck - done, opt.name, opt.lr, opt.betas, opt.eps, opt.milestones_until(max_steps), opt.gamma,
                                   seed=seed, loss_history=True)
            done = ck
            last = hist[-1].cpu().numpy() if hist is not None and len(hist) else np.full(len(blocks), np.nan)
            for i, b in enumerate(blocks):
                b.loss = float(last[i])
                grp.store_module(i, b.module)
            if on_checkpoint is not None:
                on_checkpoint(ck, grp, blocks)
        return grp

    # ---- serialisation (main.py:404-414, 589-607) ----------------------------------------------------------------
    @staticmethod
    def save_compressed(blocks: Sequence[Block], data_shape: Sequence[int], compressed_dir: str,
                        chunks_total: Optional[int] = None, write_top: bool = True) -> None:
        """compressed/{sideinfos.yaml, module/<chunk>/module/*, sideinfos/<chunk>/sideinfos.yaml} (main.py:589-607).
        chunks_total: the reference's chunks_numbers — the chunk count of the partition BEFORE the param_size_thres
        filter (main.py:529), not the number of blocks this call writes; write_top=False on every rank but one when
        several ranks write their shares into the same directory."""
        os.makedirs(compressed_dir, exist_ok=True)
        if write_top:
            with open(os.path.join(compressed_dir, "sideinfos.yaml"), "w") as fh:
                yaml.safe_dump({"data_shape": [int(x) for x in data_shape],
                                "chunks_numbers": int(len(blocks) if chunks_total is None else chunks_total)}, fh)
        for b in blocks:
            os.makedirs(os.path.join(compressed_dir, "module", b.name), exist_ok=True)
            save_model(b.module, os.path.join(compressed_dir, "module", b.name, "module"))
            os.makedirs(os.path.join(compressed_dir, "sideinfos", b.name), exist_ok=True)
            with open(os.path.join(compressed_dir, "sideinfos", b.name, "sideinfos.yaml"), "w") as fh:
                yaml.safe_dump(b.sideinfos, fh)

    # ---- decompress (main.py:270-320) -----------------------------------------------------------------------------
    def decompress_modules(self, modules: Sequence[torch.nn.Module], sideinfos: Sequence[dict]) -> List[np.ndarray]:
        """reconstruct_flattened + invnormalize_data + postprocess for many blocks in one grouped launch."""
        lo, hi = _norm_range(self.opt["Normalize"]["name"])
        phi = self.opt["Module"]["phi"]
        specs = [NetSpec(int(s["phi_features"]), phi["layers"], phi["w0"], tuple(s["data_shape"][:-1]), phi["coords_channel"],
                         phi["data_channel"]) for s in sideinfos]
        grp = SirenGroup(specs, self.device, self.precision)
        dtypes = {s["dtype"] for s in sideinfos}
        if len(dtypes) != 1:
            raise NotImplementedError("blocks of one volume share a dtype")
        dtype = dtypes.pop()
        for i, (m, s) in enumerate(zip(modules, sideinfos)):
            grp.set_axes(i, str(self.opt["Compress"]["coords_mode"]))
            grp.set_params(i, pack_module_params(m))
            grp.set_denorm(i, float(s["min"]), float(s["max"]), lo, hi)
        post_opt = self.opt["Decompress"]["postprocess"]
        level, close, clip = post_opt["denoise"]["level"], post_opt["denoise"]["close"], post_opt["clip"]
        post = None
        if dtype in ("uint8", "uint16") and not misc.preprocess_is_identity(dtype, level, clip):
            misc._limits(np.zeros(0, dtype), clip[0], clip[1])
            post = lambda i, t: preprocess_(t, level, close, clip, dtype)  # main.py:295, on the device before the copy
        elif dtype not in ("uint8", "uint16") and level > 0:
            raise NotImplementedError(f"denoise postprocess of {dtype} blocks")
        outs = grp.decompress_to_host(dtype, post=post)  # block i's device->host copy runs under the decode of block i+1
        res = []
        for t, s in zip(outs, sideinfos):
            a = t.numpy()
            a = a.view(np.uint16) if dtype == "uint16" else a
            a = a.reshape(tuple(s["data_shape"]))
            res.append(a if dtype in ("uint8", "uint16") else a.clip(clip[0], clip[1]))
        grp.close()
        return res

    @staticmethod
    def decompress(opt_path: str, module_path: str, sideinfos_path: str, device: int | str = 0) -> np.ndarray:
        """Same call as the reference's static NFGR.decompress (main.py:270-297)."""
        with open(opt_path) as fh:
            opt = yaml.safe_load(fh)
        with open(sideinfos_path) as fh:
            side = yaml.safe_load(fh)
        cf = NFGR(opt["CompressFramework"], device)
        m = init_phi(dict(cf.opt["Module"]["phi"], features=side["phi_features"], name=side["phi_name"]))
        load_model(m, module_path)
        return cf.decompress_modules([m], [side])[0]

    def decompress_divide(self, orig_sideinfos_path: str, module_save_dir: str, sideinfos_save_dir: str) -> np.ndarray:
        """main.py:299-320: every chunk directory -> one grouped decode -> merge_divided_data."""
        with open(orig_sideinfos_path) as fh:
            data_shape = yaml.safe_load(fh)["data_shape"]
        names = sorted(os.listdir(module_save_dir))
        modules, sides = [], []
        for name in names:
            with open(os.path.join(sideinfos_save_dir, name, "sideinfos.yaml")) as fh:
                side = yaml.safe_load(fh)
            m = init_phi(dict(self.opt["Module"]["phi"], features=side["phi_features"], name=side["phi_name"]))
            load_model(m, os.path.join(module_save_dir, name, "module"))
            modules.append(m)
            sides.append(side)
        decoded = self.decompress_modules(modules, sides)
        chunks = []
        for name, a in zip(names, decoded):
            rng = [[int(x) for x in part.split("_")[1:]] for part in name.split("-")]
            chunks.append({"data": a, "name": name, "d": rng[0], "h": rng[1], "w": rng[2]})
        return misc.merge_divided_data(chunks, data_shape)

    # ---- whole-volume driver (compress_divide without the process farm; multi-GPU by block ownership) ------------
    def _preprocess_volume(self, data: np.ndarray):
        """main.py:518-519: Compress.preprocess on the WHOLE volume before it is divided (an opening applied block by
        block would treat block faces as borders).  Returns (host volume, device volume [d,h,w] or None)."""
        pre = self.opt["Compress"].get("preprocess")
        dev_volume = None
        if data.dtype in (np.uint8, np.uint16) and data.ndim == 4:
            dev_volume = _to_device_raw(data[..., 0], torch.device(self.device if not isinstance(self.device, int) else f"cuda:{self.device}"))
            if pre and not misc.preprocess_is_identity(data.dtype, pre["denoise"]["level"], pre["clip"]):
                misc._limits(data, pre["clip"][0], pre["clip"][1])
                preprocess_(dev_volume, pre["denoise"]["level"], pre["denoise"]["close"], pre["clip"], data.dtype.name)
                host = dev_volume.cpu().numpy()
                data = (host.view(np.uint16) if data.dtype == np.uint16 else host)[..., None]
        elif pre:
            blk = Block("volume", data, [0, 0], [0, 0], [0, 0], 0.0)
            misc.check_float_preprocess_is_identity([blk], pre)
        return data, dev_volume

    def compress_divide(self, data: np.ndarray, compressed_dir: Optional[str] = None, max_steps: Optional[int] = None,
                        seed: int = 42, orig_bytes: Optional[int] = None, rank: int = 0, world: int = 1):
        """main.py:509-581 without the process farm: preprocess the whole volume -> partition -> allocate the budget ->
        fit this rank's LPT share of the blocks (grouped launches; blocks whose Compress.divide.exception entry changes
        the configuration are fitted as their own groups) -> (optionally) write the reference's compressed/ directory.
        Returns (all blocks, my block indices)."""
        assert data.ndim == self.opt["Module"]["phi"]["coords_channel"] + 1
        assert data.shape[-1] == self.opt["Module"]["phi"]["data_channel"]
        param_size = self.parse_param_size(orig_bytes if orig_bytes is not None else data.nbytes)
        data, dev_volume = self._preprocess_volume(data)
        blocks = self.divide(data, param_size, dev_volume)
        del dev_volume
        steps = int(self.opt["Compress"]["max_steps"] if max_steps is None else max_steps)
        exception = self.opt["Compress"]["divide"].get("exception", "none")
        exception = {} if exception in ("none", None) else exception
        # per-block configuration: the reference merges exception[<chunk name>] over the block's task yaml
        # (main.py:535-537, 568-569); only its CompressFramework sub-tree reaches this path
        frameworks: Dict[str, NFGR] = {}
        keys = []
        for b in blocks:
            over = exception.get(b.name, {})
            over = over.get("CompressFramework", over) if isinstance(over, dict) else {}
            if over:
                eff = _merge(self.opt, over)
                key = yaml.safe_dump(eff, sort_keys=True)
                if key not in frameworks:
                    frameworks[key] = NFGR(eff, self.device, self.precision, self.reproducible)
            else:
                key = ""
                frameworks.setdefault(key, self)
            keys.append(key)
        costs = []
        for b, key in zip(blocks, keys):
            cf = frameworks[key]
            given = cf.opt["Compress"]["param"].get("given_size", 0) if key else 0
            if given:
                b.param_size = float(given)
            b.features, _ = cf.estimate_module_size(b.param_size)
            batch = cf._step_batch(b.shape)
            bsteps = steps if (not key or max_steps is not None) else int(cf.opt["Compress"]["max_steps"])
            costs.append(sharding.block_cost(b.features, cf.opt["Module"]["phi"]["layers"], batch, bsteps))
        owner = sharding.lpt_assign(costs, world)
        mine = sharding.my_blocks(owner, rank)
        for key, cf in frameworks.items():
            ids = [i for i in mine if keys[i] == key]
            if ids:
                cf.fit_blocks([blocks[i] for i in ids], steps if (not key or max_steps is not None) else None, seed,
                              stream_ids=ids, preprocessed=True).close()
        if compressed_dir is not None:
            self.save_compressed([blocks[i] for i in mine], data.shape, compressed_dir, chunks_total=self.chunks_total,
                                 write_top=(rank == 0))
        return blocks, mine

    # ---- single-task driver (main.py:322-454) ---------------------------------------------------------------------
    def compress(self, data_path: str, logdir: Optional[str] = None, seed: int = 42) -> str:
        """Same entry as the reference's NFGR.compress(data_path): fit ONE network to the volume in `data_path`
        (.npy [d,h,w] / [d,h,w,1], or a multi-page TIFF read with cv2) and, at every checkpoint, write
        <logdir>/steps{N}/compressed/{sideinfos.yaml, module/} (main.py:404-414); with Compress.decompress the block
        is decoded and mse / psnr / ssim / loss go to <logdir>/performance.csv (main.py:420-450).  Side effects only,
        like the reference; returns logdir.  (TensorBoard, MIP images and the *_preprocessed copy are logging, not
        part of this path.)"""
        import csv
        data = read_volume(data_path)
        assert data.ndim == self.opt["Module"]["phi"]["coords_channel"] + 1
        logdir = logdir or os.path.join("outputs", os.path.splitext(os.path.basename(data_path))[0])
        os.makedirs(logdir, exist_ok=True)
        orig_bytes = os.path.getsize(data_path)
        d, h, w = data.shape[:3]
        blk = Block(f"d_0_{d - 1}-h_0_{h - 1}-w_0_{w - 1}", data, [0, d - 1], [0, h - 1], [0, w - 1],
                    self.parse_param_size(orig_bytes))
        C = self.opt["Compress"]
        max_steps = int(C["max_steps"])

        def on_checkpoint(step, grp, blocks):
            b = blocks[0]
            out = os.path.join(logdir, f"steps{step}")
            comp = os.path.join(out, "compressed")
            os.makedirs(comp, exist_ok=True)
            with open(os.path.join(comp, "sideinfos.yaml"), "w") as fh:
                yaml.safe_dump(b.sideinfos, fh)
            save_model(b.module, os.path.join(comp, "module"))
            if C.get("decompress", False):
                dec = self.decompress_modules([b.module], [b.sideinfos])[0]
                D = self.opt["Decompress"]
                if D.get("keep_decompressed", False):
                    os.makedirs(os.path.join(out, "decompressed"), exist_ok=True)
                    np.save(os.path.join(out, "decompressed", os.path.splitext(os.path.basename(data_path))[0] + "_decompressed.npy"), dec)
                perf = misc.eval_performance(step, data, dec, None, D.get("mse", True), D.get("psnr", True), D.get("ssim", True))
                perf["loss"] = b.loss
                path = os.path.join(logdir, "performance.csv")
                new = not os.path.exists(path)
                with open(path, "a", newline="") as fh:
                    wr = csv.writer(fh, dialect="excel")
                    if new:
                        wr.writerow(perf.keys())
                    wr.writerow([perf[k] for k in perf])

        self.fit_blocks([blk], max_steps, seed, on_checkpoint=on_checkpoint).close()
        return logdir


def read_volume(path: str) -> np.ndarray:
    """read_img of utils/tool.py:73-79 for the formats this path needs: .npy, or TIFF via cv2 -> [d,h,w,1] / [h,w,1]."""
    if path.endswith(".npy"):
        a = np.load(path)
    else:
        import cv2
        ok, pages = cv2.imreadmulti(path, flags=cv2.IMREAD_UNCHANGED)
        if not ok or not pages:
            raise FileNotFoundError(path)
        a = np.stack(pages) if len(pages) > 1 else pages[0]
    return a if a.ndim == 4 or (a.ndim == 3 and a.shape[-1] == 1) else a[..., None]
