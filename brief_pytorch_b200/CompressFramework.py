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
from .io import get_type_max
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
    data: np.ndarray                       # [d,h,w,1] view of the (pre-processed) volume, original dtype
    d: List[int]
    h: List[int]
    w: List[int]
    param_size: float                      # byte budget (alloc_param)
    features: int = 0
    module: Optional[torch.nn.Module] = None
    sideinfos: Dict = field(default_factory=dict)
    loss: float = float("nan")

    @property
    def shape(self):
        return tuple(int(x) for x in self.data.shape[:3])


class NFGR:
    def __init__(self, opt: dict, device: int | str = 0, precision: str = "auto", reproducible: bool = False):
        """reproducible=True: a block's fitted parameters are bit-identical whichever rank owns it and whatever else
        shares its GPU (per-network slicing, SirenGroup.set_slicing); the default slicing is faster and reproducible
        run to run for a fixed sharding, but may differ in the last bits between shardings."""
        self.opt = copy.deepcopy(opt)
        self.device = device
        self.precision = precision
        self.reproducible = reproducible
        if self.opt["Compress"].get("half", False):
            raise NotImplementedError("Compress.half is not part of the fused SIREN path")
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
        f = ALL_CALC_PHI_FEATURES[name](param_count=ideal_bytes / 4.0, **phi)
        return f, ALL_CALC_PHI_PARAM_COUNT[name](features=f, **phi) * 4.0

    # ---- partition (main.py:484-532) ------------------------------------------------------------------------------
    def divide(self, data: np.ndarray, param_size: float) -> List[Block]:
        dv = self.opt["Compress"]["divide"]
        kind = dv["divide_type"]
        if kind == "none":
            d, h, w = data.shape[:3]
            chunks = [{"data": data, "d": [0, d - 1], "h": [0, h - 1], "w": [0, w - 1], "size": data.size,
                       "total_size": data.size, "name": f"d_0_{d - 1}-h_0_{h - 1}-w_0_{w - 1}", "param_size": param_size}]
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
            chunks = misc.alloc_param(chunks, param_size, dv["param_alloc"], dv["param_size_thres"])
        return [Block(c["name"], c["data"], c["d"], c["h"], c["w"], float(c["param_size"])) for c in chunks]

    # ---- fit (main.py:322-454 for every block at once) -----------------------------------------------------------
    def fit_blocks(self, blocks: Sequence[Block], max_steps: Optional[int] = None, seed: int = 42,
                   sampler_generator: str = "device", on_checkpoint=None,
                   stream_ids: Optional[Sequence[int]] = None) -> SirenGroup:
        """Fit every block's network (grouped launches).  Initial weights are drawn block by block from torch's CPU
        generator in the reference's order (seed -> init_phi).  Returns the live group (parameters on the device);
        block.module / block.sideinfos / block.loss are filled in."""
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
        raws = [np.ascontiguousarray(b.data[..., 0]) for b in blocks]
        if len({r.dtype for r in raws}) != 1:
            raise NotImplementedError("blocks of one volume share a dtype")
        dev_raw = [torch.from_numpy(r.view(np.int16) if r.dtype == np.uint16 else r).to(grp.device) for r in raws]
        pre = C.get("preprocess")
        if pre and raws and not misc.preprocess_is_identity(raws[0].dtype, pre["denoise"]["level"], pre["clip"]):
            # main.py:336-337, per block like the reference's per-block processes: threshold + opening + clip on the
            # device copy (brief_preprocess); min / max and the weights below are those of the preprocessed block
            misc._limits(raws[0], pre["clip"][0], pre["clip"][1])
            needs_host = any(r.split("_")[0] in ("quantile", "exp") for r in C["loss"]["weight"])
            for i, t in enumerate(dev_raw):
                preprocess_(t, pre["denoise"]["level"], pre["denoise"]["close"], pre["clip"], raws[i].dtype.name)
                if needs_host:
                    h = t.cpu().numpy()
                    raws[i] = h.view(np.uint16) if raws[i].dtype == np.uint16 else h
        stats = block_stats(dev_raw, raws[0].dtype.name)
        for i, b in enumerate(blocks):
            grp.set_axes(i, str(C["coords_mode"]))
            grp.load_module(i, b.module)
            raw = raws[i]
            vmin, vmax = float(stats[i, 0]), float(stats[i, 1])
            b.sideinfos = {"dtype": raw.dtype.name, "min": vmin, "max": vmax, "data_shape": list(b.data.shape),
                           "phi_features": int(b.features), "phi_name": self.opt["Module"]["phi"]["name"]}
            t = dev_raw[i]
            rules = misc.weight_rules_for_kernel(raw, C["loss"]["weight"])
            weight = None
            if rules is None:  # 'exp' rule or more than 4 rules: explicit per-voxel weights
                weight = torch.from_numpy(misc.parse_weight(raw, C["loss"]["weight"]).reshape(-1)).to(grp.device)
            thres = C["loss"].get("weight_thres", 0)
            assert thres <= get_type_max(raw)  # main.py:380
            tau = ((_f32(thres) - _f32(vmin)) / (_f32(vmax) - _f32(vmin))) * (hi - lo) + lo if thres else 0.0
            grp.bind_volume(i, t, vmin, vmax, lo, hi, weight=weight, rules=rules or (), tau=_f32(tau), np_dtype=raw.dtype.name)
            self._keep.append((t, weight))
            # main.py:325-334: the cube sampler (= whole block per step) only for blocks of at most 80^3 voxels
            name = C["sampler"]["name"]
            if name == "randomcube" and raw.size > 80 ** 3:
                name = "randompoint"
            if name not in ("randomcube", "randompoint"):
                raise NotImplementedError(name)  # main.py:371
            grp.set_sampler(i, name, int(C["sampler"]["sample_size"]))
        opt = misc.configure_lr_scheduler(misc.configure_optimizer(None, C["optimizer_name_phi"], C["lr_phi"]),
                                          C["lr_scheduler_phi"])
        done = 0
        for ck in checkpoints:
            hist = grp.fit_run(ck - done, opt.name, opt.lr, opt.betas, opt.eps, opt.milestones, opt.gamma, seed=seed,
                               loss_history=True)
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
    def save_compressed(blocks: Sequence[Block], data_shape: Sequence[int], compressed_dir: str) -> None:
        os.makedirs(compressed_dir, exist_ok=True)
        with open(os.path.join(compressed_dir, "sideinfos.yaml"), "w") as fh:
            yaml.safe_dump({"data_shape": [int(x) for x in data_shape], "chunks_numbers": len(blocks)}, fh)
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
    def compress_divide(self, data: np.ndarray, compressed_dir: Optional[str] = None, max_steps: Optional[int] = None,
                        seed: int = 42, orig_bytes: Optional[int] = None, rank: int = 0, world: int = 1):
        """Partition -> allocate budget -> fit this rank's LPT share of the blocks -> (optionally) write the
        reference's compressed/ directory.  Returns (all blocks, my block indices)."""
        assert data.ndim == self.opt["Module"]["phi"]["coords_channel"] + 1
        assert data.shape[-1] == self.opt["Module"]["phi"]["data_channel"]
        param_size = self.parse_param_size(orig_bytes if orig_bytes is not None else data.nbytes)
        blocks = self.divide(data, param_size)
        steps = int(self.opt["Compress"]["max_steps"] if max_steps is None else max_steps)
        L = self.opt["Module"]["phi"]["layers"]
        costs = []
        for b in blocks:
            b.features, _ = self.estimate_module_size(b.param_size)
            batch = b.data.size if b.data.size <= 80 ** 3 else int(self.opt["Compress"]["sampler"]["sample_size"])
            costs.append(sharding.block_cost(b.features, L, batch, steps))
        owner = sharding.lpt_assign(costs, world)
        mine = sharding.my_blocks(owner, rank)
        if mine:
            self.fit_blocks([blocks[i] for i in mine], steps, seed, stream_ids=mine).close()
        if compressed_dir is not None:
            self.save_compressed([blocks[i] for i in mine], data.shape, compressed_dir)
        return blocks, mine
