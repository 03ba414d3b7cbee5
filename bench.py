#!/usr/bin/env python
"""bench.py — the BRIEF hot path on B200: SIREN fit coord-samples/s (fwd+bwd+optimiser), with decompress voxels/s
beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload vessel|vessel64|config1]
                    [--precision auto|f16|fp32]

Workload at N=1 = BASELINE.json configs[1]: DivideTask opt/DivideTask/vessel.yaml on a synthetic 64x512x512 uint16
volume (ratio 128, `adaptotal_-1_-1_-1_4` -> 4 blocks of 64x256x256, one SIREN L=7 f=56 w0=10 per block,
RandompointSampler batch 100000 per block per step, Adamax lr 1e-3).  A "step" = one training step of every block's
network (gather + forward + weighted L2 + backward + Adamax).  For N>1 every rank owns the LPT share of 4N blocks
(one such volume per GPU: weak scaling), no collective on the fit path; per-block loss statistics are all-gathered
after the timed region.

Named shapes of BASELINE.json configs 3-5 (`--workload neuron1024 | neuron1024_nb4 | hipct2048 | hipct2048_equal |
decomp4096 | decomp4096_f39`): ONE volume of that shape, generated block by block on the device, its blocks sharded over
the `--gpus N` ranks by parameter-weighted LPT (strong scaling; no collective on the fit or decode path), with per-rank
busy time and LPT imbalance in the line.  The default run also carries a small strong-scaling leg
(`strong_scaling`) and one short measurement per block geometry (`workloads`).

One JSON line on stdout (rank 0); see README/DESIGN.md for the keys.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# Library banners (NCCL prints its version on fd 1) must not pollute the one-JSON-line contract: everything written to
# stdout goes to stderr, and emit() writes the result line to the real stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


METRIC = "siren_fit_coord_samples_per_s"
UNIT = "coord-samples/s"

WORKLOADS = {
    # name: (volume shape, ratio, Nb, layers, w0, batch (0 = whole block), description)
    "vessel": ((64, 512, 512), 128, 4, 7, 10.0, 100000,
               "DivideTask vessel.yaml as shipped: synthetic 64x512x512 u16, ratio 128, Nb=4 -> 4 blocks 64x256x256, "
               "SIREN L=7 f=56 w0=10, randompoint batch 100000/block, Adamax"),
    "vessel64": ((64, 512, 512), 128, 64, 7, 10.0, 0,
                 "DivideTask vessel.yaml with Nb=64: 64 blocks 64^3, SIREN L=7 f=13, whole-block batch 262144, Adamax"),
    "config1": ((64, 64, 64), 80, 1, 5, 20.0, 0,
                "SingleTask default.yaml shape: 64^3 u16 block, SIREN L=5 f=22 w0=20, whole-block batch, Adamax"),
    # sub-volumes with the block shape and width the big configs produce, 8 resp. 4 blocks per GPU
    "neuron128": ((256, 256, 256), 512, 8, 7, 10.0, 100000,
                  "DivideTask neuron.yaml geometry (1024^3 ratio 512, auto Nb=512): 128^3 blocks, SIREN L=7 f=19 w0=10, "
                  "randompoint batch 100000/block, Adamax; 8 blocks per GPU"),
    "hipct256": ((256, 512, 512), 128, 4, 7, 10.0, 100000,
                 "DivideTask hipct.yaml geometry (2048^3 ratio 128, Nb=512): 256^3 blocks, SIREN L=7 f=113 w0=10, "
                 "randompoint batch 100000/block, Adamax; 4 blocks per GPU (wide tcgen05 fit kernel: streamed weights, stashed activations)"),
}


# ONE volume of the named shape, blocks sharded over the ranks (strong scaling).  kind = synth_device generator.
NAMED = {
    "neuron1024": dict(kind="neuron", shape=(1024, 1024, 1024), ratio=512, nb=-1, alloc="by_size", thres=100, layers=7, w0=10.0,
                       rules=[(10001, 65535, 0.1)], fit=True,
                       desc="DivideTask neuron.yaml on a synthetic 1024^3 u16 volume, ratio 512, divide adaptotal_-1_-1_-1_-1 (auto Nb = "
                            "param bytes / (4*1361) = 770 -> grid 8x8x8 = 512 blocks 128^3, SIREN L=7 f=19), by_size budgets, weight "
                            "rule value_10001_65535_0.1, randompoint batch 100000/block, Adamax"),
    "neuron1024_nb4": dict(kind="neuron", shape=(1024, 1024, 1024), ratio=512, nb=4, alloc="by_size", thres=100, layers=7, w0=10.0,
                           rules=[(10001, 65535, 0.1)], fit=True,
                           desc="DivideTask neuron.yaml AS SHIPPED (Nb=4) on a synthetic 1024^3 u16 volume: 4 blocks 1024x512x512, "
                                "SIREN L=7 f=228 (F_PAD = 240: layer-wise tcgen05 kernels), randompoint batch 100000/block, Adamax"),
    "hipct2048": dict(kind="hipct", shape=(2048, 2048, 2048), ratio=128, nb=512, alloc="by_var", thres=26, layers=7, w0=10.0,
                      rules=[(65535, 65535, 1.0)], fit=True,
                      desc="DivideTask hipct.yaml on a synthetic 2048^3 u16 volume, ratio 128, Nb=512 -> 512 blocks 256^3, by_var "
                           "budgets (per-block widths around f=113), randompoint batch 100000/block, Adamax"),
    "hipct2048_equal": dict(kind="hipct", shape=(2048, 2048, 2048), ratio=128, nb=512, alloc="equal", thres=26, layers=7, w0=10.0,
                            rules=[(65535, 65535, 1.0)], fit=True,
                            desc="hipct.yaml geometry (2048^3, ratio 128, Nb=512 -> 512 blocks 256^3) with EQUAL budgets: every block "
                                 "f=113 (wide tcgen05 fit kernel), randompoint batch 100000/block, Adamax"),
    "decomp4096": dict(kind=None, shape=(4096, 4096, 4096), ratio=128, nb=4096, alloc="equal", thres=26, layers=7, w0=10.0,
                       rules=[], fit=False,
                       desc="Decompress-only sweep: 4096^3 u16 output grid from 4096 stored per-block SIRENs (256^3 blocks, L=7 "
                            "f=113, reference initialisation), output kept sharded by block owner"),
    "decomp4096_f39": dict(kind=None, shape=(4096, 4096, 4096), ratio=128, nb=32768, alloc="equal", thres=26, layers=7, w0=10.0,
                           rules=[], fit=False,
                           desc="Decompress-only sweep: 4096^3 u16 output grid from 32768 stored per-block SIRENs (128^3 blocks, "
                                "L=7 f=39), output kept sharded by block owner"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


def ncu_traffic(name):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/ncu_traffic.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[name]["bytes_per_launch"]
    except Exception:
        return None


def plan_blocks(name):
    """Block grid + per-block width exactly as the reference sizes them (cal_divide_num, alloc_param 'equal' with
    by-size-equal blocks, SIREN.calc_features)."""
    from brief_pytorch_b200 import misc
    from brief_pytorch_b200.Networks import SIREN
    shape, ratio, nb, layers, w0, batch, desc = WORKLOADS[name]
    raw_bytes = int(np.prod(shape)) * 2
    param_bytes = raw_bytes / ratio
    grid = [1, 1, 1] if nb == 1 else [int(x) for x in misc.cal_divide_num(*shape, nb, param_bytes)]
    n_blocks = grid[0] * grid[1] * grid[2]
    f = SIREN.calc_features(param_count=param_bytes / n_blocks / 4.0, coords_channel=3, data_channel=1, layers=layers)
    bshape = tuple(shape[i] // grid[i] for i in range(3))
    return dict(shape=shape, grid=grid, n_blocks=n_blocks, block_shape=bshape, features=f, layers=layers, w0=w0,
                batch=batch if batch else int(np.prod(bshape)), full_block=(batch == 0), desc=desc, name=name)


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, f"/tmp/brief_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def fit_config(plan, n_local, prec, world):
    """`config` of the default fit line; the reference arm prints the same object (it times the same workload, its own
    arithmetic is in `dtype` and what a step of it covers is in `cpu_baseline.sample`)."""
    return {"workload": plan["desc"], "blocks_per_gpu": n_local, "block_shape": list(plan["block_shape"]),
            "features": plan["features"], "layers": plan["layers"], "batch_per_block": plan["batch"],
            "precision": prec, "l2": "flushed (256 MiB write) between timed steps",
            "parallelism": f"blocks sharded by LPT over {world} GPU(s), no collective on the fit path"}


def oracle_fit_rate(plan, seconds, threads):
    """The reference's CPU PyTorch path (oracle restatement of main.py:385-400 around torch CPU ops) on ONE block of
    the workload; returns (coord-samples/s, steps timed)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import brief_oracle as O
    from brief_pytorch_b200 import synth
    torch.set_num_threads(threads)
    bs = plan["block_shape"]
    blk = synth.vessel(bs, seed=42)
    data_t, side = O.normalize_data(blk.copy(), "minmaxany_0_100")
    weight = O.parse_weight(blk.copy(), ["value_65535_65535_1"])
    thr = O.weight_thres_normalized(65535, "minmaxany_0_100", side["min"], side["max"])
    torch.manual_seed(42)
    phi = O.init_phi(dict(coords_channel=3, data_channel=1, name="SIREN", layers=plan["layers"], w0=plan["w0"],
                          features=plan["features"]))
    opt = O.configure_optimizer(phi.parameters(), "Adamax", 1e-3)
    sch = O.configure_lr_scheduler(opt, {"name": "MultiStepLR", "milestones": [50000, 60000, 70000], "gamma": 0.2})
    if plan["full_block"]:
        sampler = O.RandomCubeSampler(data_t, weight, "-1,1", 1, [10000000] * 3, 10 ** 9)
    else:
        sampler = O.RandompointSampler(data_t, weight, "-1,1", plan["batch"], 10 ** 9)
    it = iter(sampler)

    def step():
        c, d, w = next(it)
        return float(O.train_step(phi, opt, sch, c, d, w, thr).detach())  # .item() per step like main.py:401
    return step


def run_reference(args, plan):
    """--impl reference: the reference's own CPU path (oracle port; the Python reference cannot travel to the box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    step = oracle_fit_rate(plan, 0, threads)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = plan["batch"] * args.steps / dt
    sample = f"one block ({'x'.join(map(str, plan['block_shape']))}, f={plan['features']}) per step, batch {plan['batch']}"
    emit(({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": fit_config(plan, plan["n_blocks"], "fp32" if args.precision == "fp32" else "f16", args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))



# ======================================================================================================================
# named shapes: ONE volume, blocks sharded over the ranks by LPT (strong scaling)
# ======================================================================================================================
def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def plan_named(cfg, dev, rank, world):
    """Partition + budgets + widths of a named volume exactly as the reference sizes them (cal_divide_num, divide_data's
    block ranges, alloc_param, SIREN.calc_features).  by_var needs every block's variance: each rank generates the blocks
    of an even share, takes their sums with ONE brief_block_stats launch, and the table is all-gathered."""
    import torch.distributed as dist
    from brief_pytorch_b200 import misc, sharding, synth_device
    from brief_pytorch_b200.group import block_stats
    from brief_pytorch_b200.Networks import SIREN
    shape = cfg["shape"]
    param_bytes = int(np.prod(shape)) * 2 / cfg["ratio"]
    grid = [int(x) for x in misc.cal_divide_num(*shape, cfg["nb"], param_bytes)]
    ext = [shape[k] // grid[k] for k in range(3)]
    assert all(ext[k] * grid[k] == shape[k] for k in range(3))
    ranges = [((iz * ext[0], iy * ext[1], ix * ext[2]), ((iz + 1) * ext[0], (iy + 1) * ext[1], (ix + 1) * ext[2]))
              for iz in range(grid[0]) for iy in range(grid[1]) for ix in range(grid[2])]
    n_blocks = len(ranges)
    size = int(np.prod(ext))
    chunks = [{"size": size, "total_size": size * n_blocks, "name": i} for i in range(n_blocks)]
    t_var = 0.0
    if cfg["alloc"] == "by_var":
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        table = torch.zeros((n_blocks, 2), dtype=torch.float64, device=dev)
        for i in range(rank, n_blocks, world):
            blk = synth_device.block(cfg["kind"], shape, ranges[i][0], ranges[i][1], 42, i, dev)
            st = block_stats([blk], "uint16")
            table[i, 0], table[i, 1] = float(st[0, 2]), float(st[0, 3])
        if world > 1:
            dist.all_reduce(table)
        torch.cuda.synchronize()
        t_var = time.perf_counter() - t0
        for c, row in zip(chunks, table.cpu().numpy()):
            c["var"] = misc.variance_from_sums(float(row[0]), float(row[1]), size)
    chunks = misc.alloc_param(chunks, param_bytes, cfg["alloc"], cfg["thres"])
    feats = {}
    for c in chunks:
        feats[c["name"]] = SIREN.calc_features(param_count=c["param_size"] / 4.0, coords_channel=3, data_channel=1, layers=cfg["layers"])
    ids = sorted(feats)
    batch = size if size <= 80 ** 3 else 100000
    costs = [sharding.block_cost(feats[i], cfg["layers"], batch, 80000) for i in ids]
    owner = sharding.lpt_assign(costs, world)
    return dict(grid=grid, ext=tuple(ext), ranges=ranges, ids=ids, feats=feats, owner=owner, costs=costs, batch=batch,
                full_block=size <= 80 ** 3, n_blocks_total=n_blocks, t_var=t_var)


def run_named(args):
    """`--workload <named shape>`: see measure_named."""
    cfg = NAMED[args.workload]
    rank, world, local = _dist_env()
    if args.impl == "reference":
        return run_reference_named(args, cfg)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    if args.steps == 1000 and args.warmup == 50:
        args.steps, args.warmup = 10, 3
    line = measure_named(cfg, args, dev, rank, world, local)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def measure_named(cfg, args, dev, rank, world, local):
    """One volume of a named shape (BASELINE configs 3-5) sharded over the ranks: fit throughput (device-timed, max over
    ranks) and decompress throughput with the output kept sharded by block owner.  Returns the JSON line on rank 0."""
    import zlib
    import torch.distributed as dist
    from brief_pytorch_b200 import sharding, synth_device
    from brief_pytorch_b200.group import NetSpec, SirenGroup, launch_count, pack_module_params, reset_launch_count
    from brief_pytorch_b200.Networks import init_phi
    pk = peaks()
    P = plan_named(cfg, dev, rank, world)
    mine = [k for k, i in enumerate(P["ids"]) if P["owner"][k] == rank]
    feats = [P["feats"][P["ids"][k]] for k in mine]
    specs = [NetSpec(f, cfg["layers"], cfg["w0"], P["ext"]) for f in feats]
    t0 = time.perf_counter()
    grp = SirenGroup(specs, dev, args.precision) if specs else None
    if grp is not None:
        grp.set_slicing(args.reproducible)
    keep = []
    cache = {}
    for j, k in enumerate(mine):
        i = P["ids"][k]
        f = feats[j]
        if f not in cache:  # reference initialisation (seed 42 -> init_phi), one draw per width like the reference's per-block processes
            torch.manual_seed(42)
            cache[f] = pack_module_params(init_phi(dict(name="SIREN", coords_channel=3, data_channel=1, layers=cfg["layers"],
                                                        w0=cfg["w0"], features=f)))
        grp.set_params(j, cache[f])
        grp.set_stream(j, i)
        if not cfg["fit"]:
            grp.set_denorm(j, 1000.0, 40000.0, 0.0, 100.0)
        if cfg["fit"]:
            blk = synth_device.block(cfg["kind"], cfg["shape"], P["ranges"][i][0], P["ranges"][i][1], 42, i, dev)
            keep.append(blk)
    vox_local = len(mine) * int(np.prod(P["ext"]))
    fit = None
    precs = sorted({grp.precision(j) for j in range(len(mine))}) if grp is not None else []
    if cfg["fit"] and grp is not None:
        from brief_pytorch_b200.group import block_stats
        st = block_stats(keep, "uint16")
        for j in range(len(mine)):
            vmin, vmax = float(st[j, 0]), float(st[j, 1])
            tau = (65535.0 - vmin) / max(vmax - vmin, 1.0) * 100.0
            grp.bind_volume(j, keep[j], vmin, vmax if vmax > vmin else vmin + 1.0, 0.0, 100.0, rules=cfg["rules"], tau=tau, np_dtype="uint16")
            grp.set_sampler(j, "randomcube" if P["full_block"] else "randompoint", P["batch"])
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    opt_kw = dict(kind="Adamax", lr=1e-3, milestones=(50000, 60000, 70000), gamma=0.2, seed=42)
    clocks = ClockSampler(local)
    t_fit = 0.0
    checksum = 0
    if cfg["fit"]:
        if grp is not None:
            grp.fit_run(args.warmup, **opt_kw)
        barrier()
        clocks.start()
        reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        if grp is not None:
            grp.fit_run(args.steps, **opt_kw)  # blocks of a real volume stay resident across steps: no flush (inputs >> L2 anyway)
        e1.record()
        barrier()
        t_fit = e0.elapsed_time(e1) / 1e3
        if args.reproducible:
            for j, k in enumerate(mine):
                checksum ^= zlib.crc32(grp.get_params(j).tobytes() + int(P["ids"][k]).to_bytes(4, "little"))
    else:
        clocks.start()
    launches = launch_count() if cfg["fit"] else 0

    # ---- decompress: every own block, output kept on the owner (sharded volume) ----
    reset_launch_count()
    t_dec = 0.0
    out_bytes = vox_local * 2
    if grp is not None:
        free, _ = torch.cuda.mem_get_info(dev)
        slab = len(mine)
        while slab > 1 and slab * int(np.prod(P["ext"])) * 2 > free - (6 << 30):
            slab = (slab + 1) // 2
        outs = [torch.empty(P["ext"], dtype=torch.int16, device=dev) for _ in range(slab)]
        resident = slab == len(mine)
        grp.decompress("uint16", out=outs + [None] * (len(mine) - slab), nets=range(min(slab, 4)))  # warm-up (pack + a few blocks)
        barrier()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for lo in range(0, len(mine), slab):
            ids = list(range(lo, min(lo + slab, len(mine))))
            full = [None] * len(mine)
            for q, j in enumerate(ids):
                full[j] = outs[q]
            grp.decompress("uint16", out=full, nets=ids)
        d1.record()
        barrier()
        t_dec = d0.elapsed_time(d1) / 1e3
    else:
        barrier(); barrier()
        resident = True
    dec_launches = launch_count()
    clk = clocks.stop()

    samples_local = len(mine) * P["batch"] * args.steps if cfg["fit"] else 0
    fit_flops_local = sum(sharding.fit_flops_per_sample(f, cfg["layers"]) for f in feats) * P["batch"] * args.steps if cfg["fit"] else 0
    fwd_flops_local = sum(sharding.forward_flops_per_sample(f, cfg["layers"]) for f in feats) * int(np.prod(P["ext"]))
    mine_cost = sum(P["costs"][k] for k in mine)
    vals = torch.tensor([t_fit, t_dec, samples_local, vox_local, launches, fit_flops_local, fwd_flops_local, mine_cost, t_setup,
                         float(len(mine)), float(checksum)], dtype=torch.float64, device=dev)
    if world > 1:
        allv = [torch.empty_like(vals) for _ in range(world)]
        dist.all_gather(allv, vals)
        allv = torch.stack(allv).cpu().numpy()
    else:
        allv = vals.cpu().numpy()[None]
    if rank == 0:
        tf, td = allv[:, 0], allv[:, 1]
        tmax_fit, tmax_dec = float(tf.max()), float(td.max())
        total_samples, total_vox = float(allv[:, 2].sum()), float(allv[:, 3].sum())
        widths = sorted(P["feats"].values())
        cks = 0
        for c in allv[:, 10]:
            cks ^= int(c)
        line = {
            "metric": METRIC if cfg["fit"] else "siren_decompress_voxels_per_s",
            "value": (total_samples / tmax_fit) if cfg["fit"] else (total_vox / tmax_dec),
            "unit": UNIT if cfg["fit"] else "voxels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": (1e3 * tmax_fit / args.steps) if cfg["fit"] else 1e3 * tmax_dec,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16" if precs == ["f16"] else ("f32" if precs == ["fp32"] else "f16+f32"), "data": "synthetic",
            "config": {"workload": cfg["desc"], "volume": list(cfg["shape"]), "grid": P["grid"], "block_shape": list(P["ext"]),
                       "blocks": len(P["ids"]), "blocks_dropped_by_thres": P["n_blocks_total"] - len(P["ids"]),
                       "features_min_median_max": [int(widths[0]), int(widths[len(widths) // 2]), int(widths[-1])],
                       "layers": cfg["layers"], "batch_per_block": P["batch"], "param_alloc": cfg["alloc"],
                       "l2": "inputs larger than L2 (raw blocks of the rank >> 126 MB); no flush between steps",
                       "parallelism": f"blocks of ONE volume sharded by parameter-weighted LPT over {world} GPU(s); no collective on the fit or decode path",
                       "reproducible_slicing": bool(args.reproducible)},
            "gpu_launches": int(allv[:, 4].sum()) + dec_launches * world,
            "per_rank": {"blocks": [int(x) for x in allv[:, 9]], "fit_busy_ms_per_step": [round(1e3 * float(x) / max(args.steps, 1), 4) for x in tf],
                         "decompress_busy_ms": [round(1e3 * float(x), 3) for x in td],
                         "lpt_cost_share": [round(float(x) / max(float(allv[:, 7].sum()), 1e-30), 5) for x in allv[:, 7]],
                         "setup_s": [round(float(x), 2) for x in allv[:, 8]]},
            "imbalance": {"fit_max_over_mean": float(tf.max() / max(tf.mean(), 1e-30)) if cfg["fit"] else None,
                          "decompress_max_over_mean": float(td.max() / max(td.mean(), 1e-30)),
                          "lpt_cost_max_over_mean": float(allv[:, 7].max() / max(allv[:, 7].mean(), 1e-30))},
            "decompress": {"value": total_vox / tmax_dec, "unit": "voxels/s", "ms": 1e3 * tmax_dec,
                           "hbm_gbs": total_vox * 2 / tmax_dec / 1e9, "hbm_frac": total_vox * 2 / tmax_dec / 1e9 / (pk["hbm_gbs"] * world),
                           "tflops": float(allv[:, 6].sum()) / tmax_dec / 1e12,
                           "tensor_frac": float(allv[:, 6].sum()) / tmax_dec / 1e12 / (pk["bf16_tflops"] * world),
                           "output": "sharded by block owner" + ("" if resident else ", decoded slab by slab into a reused buffer"),
                           "output_bytes_total": int(total_vox * 2)},
            "clocks": clk,
        }
        if cfg["fit"]:
            tf_total = float(allv[:, 5].sum()) / tmax_fit / 1e12
            line["roofline"] = {"bound": "tensor", "achieved": tf_total, "peak": pk["bf16_tflops_sustained"] * world, "unit": "TFLOP/s",
                                "frac": tf_total / (pk["bf16_tflops_sustained"] * world), "traffic": None,
                                "peak_src": pk["src"] + " bf16 sustained x n_gpus (kernel timed inside a long step loop)",
                                "kernel": "whole step (fit kernels of every width bucket + optimiser), algorithmic FLOPs of all blocks"}
            line["e2e"] = None
            line["by_var_stats_s"] = P["t_var"]
        else:
            line["roofline"] = {"bound": "hbm", "achieved": total_vox * 2 / tmax_dec / 1e9, "peak": pk["hbm_gbs"] * world, "unit": "GB/s",
                                "frac": total_vox * 2 / tmax_dec / 1e9 / (pk["hbm_gbs"] * world), "traffic": None,
                                "peak_src": pk["src"] + " copy bandwidth x n_gpus",
                                "note": "asked for against HBM; the binding unit is the SFU ((L-1) f sines per voxel), see DESIGN.md 4.2"}
        if args.reproducible:
            line["param_checksum"] = f"{cks:08x}"
        return line
    return None


def run_reference_named(args, cfg):
    """--impl reference for a named shape: the oracle's CPU loop on ONE block of the volume's typical width."""
    rank, world, local = _dist_env()
    if rank != 0:
        return
    from brief_pytorch_b200 import misc
    from brief_pytorch_b200.Networks import SIREN
    shape = cfg["shape"]
    param_bytes = int(np.prod(shape)) * 2 / cfg["ratio"]
    grid = [int(x) for x in misc.cal_divide_num(*shape, cfg["nb"], param_bytes)]
    n_blocks = grid[0] * grid[1] * grid[2]
    ext = tuple(shape[k] // grid[k] for k in range(3))
    f = SIREN.calc_features(param_count=param_bytes / n_blocks / 4.0, coords_channel=3, data_channel=1, layers=cfg["layers"])
    sample_shape = tuple(min(e, 128) for e in ext)  # the oracle materialises 20 B/voxel of fp32 copies: bound the block
    plan = dict(block_shape=sample_shape, layers=cfg["layers"], w0=cfg["w0"], features=f, batch=100000, full_block=False,
                desc=cfg["desc"])
    if not cfg["fit"]:
        emit({"impl": "reference", "unavailable": "decompress-only workload: the reference arm times the fit metric"})
        return
    if args.steps == 1000 and args.warmup == 50:
        args.steps, args.warmup = 10, 3
    run_reference(args, plan)


def quick_workload(name, dev, precision, pk, flush, steps=60, warmup=10):
    """One short device-resident measurement of a block geometry (the per-GPU share of a WORKLOADS entry): step rate with
    the L2 flushed between steps, the fit kernel(s) alone for the roofline fraction, and the decode rate."""
    from brief_pytorch_b200 import sharding, synth_device
    from brief_pytorch_b200.group import NetSpec, SirenGroup, block_stats, pack_module_params
    from brief_pytorch_b200.Networks import init_phi
    plan = plan_blocks(name)
    bs, gd = plan["block_shape"], plan["grid"]
    n = plan["n_blocks"]
    grp = SirenGroup([NetSpec(plan["features"], plan["layers"], plan["w0"], bs) for _ in range(n)], dev, precision)
    torch.manual_seed(42)
    p0 = pack_module_params(init_phi(dict(name="SIREN", coords_channel=3, data_channel=1, layers=plan["layers"], w0=plan["w0"],
                                          features=plan["features"])))
    keep = []
    for k in range(n):
        iz, iy, ix = k // (gd[1] * gd[2]), (k // gd[2]) % gd[1], k % gd[2]
        lo = (iz * bs[0], iy * bs[1], ix * bs[2])
        keep.append(synth_device.block("vessel", plan["shape"], lo, tuple(lo[i] + bs[i] for i in range(3)), 42, k, dev))
    st = block_stats(keep, "uint16")
    for k in range(n):
        vmin, vmax = float(st[k, 0]), float(st[k, 1])
        grp.set_params(k, p0)
        grp.bind_volume(k, keep[k], vmin, vmax, 0.0, 100.0, rules=[(65535, 65535, 1.0)], tau=(65535.0 - vmin) / (vmax - vmin) * 100.0,
                        np_dtype="uint16")
        grp.set_sampler(k, "randomcube" if plan["full_block"] else "randompoint", plan["batch"])
    opt_kw = dict(kind="Adamax", lr=1e-3, milestones=(50000, 60000, 70000), gamma=0.2, seed=42)
    grp.fit_run(warmup, **opt_kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_step = t_kern = 0.0
    for s_ in range(steps):
        flush.fill_(s_ & 0xFF)
        e0.record(); grp.fit_run(1, **opt_kw); e1.record()
        torch.cuda.synchronize()
        t_step += e0.elapsed_time(e1) / 1e3
    n_k = min(steps, 30)
    for s_ in range(n_k):
        flush.fill_(s_ & 0xFF)
        e0.record(); grp.fit_kernel_only(seed=42, step=s_); e1.record()
        torch.cuda.synchronize()
        t_kern += e0.elapsed_time(e1) / 1e3
    outs = grp.decompress("uint16")
    torch.cuda.synchronize()
    t_dec = 0.0
    for s_ in range(3):
        flush.fill_(s_)
        e0.record(); grp.decompress("uint16", out=outs); e1.record()
        torch.cuda.synchronize()
        t_dec += e0.elapsed_time(e1) / 1e3
    samples = n * plan["batch"]
    flops = sharding.fit_flops_per_sample(plan["features"], plan["layers"])
    tf = samples * flops / (t_kern / n_k) / 1e12
    res = {"value": samples * steps / t_step, "unit": UNIT, "ms_per_step": 1e3 * t_step / steps, "blocks": n,
           "block_shape": list(bs), "features": plan["features"], "layers": plan["layers"], "batch_per_block": plan["batch"],
           "precision": grp.precision(0), "fit_kernel_ms": 1e3 * t_kern / n_k, "fit_tflops": tf,
           "roofline_frac": tf / pk["bf16_tflops"], "decompress_voxels_per_s": n * int(np.prod(bs)) / (t_dec / 3)}
    grp.close()
    return res


def gpu_eager_rate(plan, seconds, dev):
    """The reference's REAL GPU path beside the CPU one: the same oracle loop (main.py:385-401 around stock torch ops, fp32,
    eager) with module, data and sampler tensors on the B200 — cuBLAS sgemm + elementwise kernels, ~700 ATen ops per step,
    index draw on the CPU and loss.item() every step like the reference.  One block of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import brief_oracle as O
    from brief_pytorch_b200 import synth
    bs = plan["block_shape"]
    blk = synth.vessel(bs, seed=42)
    data_t, side = O.normalize_data(blk.copy(), "minmaxany_0_100")
    weight = O.parse_weight(blk.copy(), ["value_65535_65535_1"])
    thr = O.weight_thres_normalized(65535, "minmaxany_0_100", side["min"], side["max"])
    torch.manual_seed(42)
    phi = O.init_phi(dict(coords_channel=3, data_channel=1, name="SIREN", layers=plan["layers"], w0=plan["w0"],
                          features=plan["features"])).to(dev)
    opt = O.configure_optimizer(phi.parameters(), "Adamax", 1e-3)
    sch = O.configure_lr_scheduler(opt, {"name": "MultiStepLR", "milestones": [50000, 60000, 70000], "gamma": 0.2})
    if plan["full_block"]:
        sampler = O.RandomCubeSampler(data_t.to(dev), weight, "-1,1", 1, [10000000] * 3, 10 ** 9, device=str(dev), gpu_force=True)
    else:
        sampler = O.RandompointSampler(data_t.to(dev), weight, "-1,1", plan["batch"], 10 ** 9, device=str(dev))
    it = iter(sampler)

    def step():
        c, d, w = next(it)
        return float(O.train_step(phi, opt, sch, c, d, w, thr).detach())
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds and n < 2000:
        step(); n += 1
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"value": plan["batch"] * n / dt, "unit": UNIT, "kind": "reference's own GPU path: oracle loop on torch-CUDA eager fp32 "
            "(TF32 off), one block at a time, CPU index draw + loss.item() per step", "ms_per_step": 1e3 * dt / n,
            "sample": f"{n} steps of one block (f={plan['features']}, batch {plan['batch']})"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="vessel", choices=sorted(WORKLOADS) + sorted(NAMED))
    ap.add_argument("--reproducible", action="store_true",
                    help="per-network slicing: every block's fitted parameters are bit-identical for any number of GPUs "
                         "(the line carries a checksum of all fitted parameters to compare across N)")
    ap.add_argument("--no-side-legs", action="store_true", help="skip the HBM side legs, per-geometry table and strong-scaling leg")
    ap.add_argument("--precision", default="auto", choices=["auto", "f16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.workload in NAMED:
        return run_named(args)
    plan = plan_blocks(args.workload)
    if args.impl == "reference":
        if args.steps == 1000 and args.warmup == 50:
            args.steps, args.warmup = 20, 3
        return run_reference(args, plan)

    from brief_pytorch_b200 import sharding, synth
    from brief_pytorch_b200.group import (NetSpec, SirenGroup, launch_count, pack_module_params, reset_launch_count)
    from brief_pytorch_b200.Networks import init_phi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()

    # ---- blocks: N volumes' worth (weak scaling), LPT-assigned by parameter-weighted cost ----
    n_total = plan["n_blocks"] * world
    costs = [sharding.block_cost(plan["features"], plan["layers"], plan["batch"], 80000)] * n_total
    owner = sharding.lpt_assign(costs, world)
    mine = sharding.my_blocks(owner, rank)
    vol = synth.vessel(plan["shape"], seed=42 + rank)
    bs, gd = plan["block_shape"], plan["grid"]
    blocks = []
    for j, b in enumerate(mine):
        k = j % plan["n_blocks"]
        iz, iy, ix = k // (gd[1] * gd[2]), (k // gd[2]) % gd[1], k % gd[2]
        blocks.append(np.ascontiguousarray(vol[iz * bs[0]:(iz + 1) * bs[0], iy * bs[1]:(iy + 1) * bs[1],
                                               ix * bs[2]:(ix + 1) * bs[2], 0]))
    specs = [NetSpec(plan["features"], plan["layers"], plan["w0"], bs) for _ in mine]
    grp = SirenGroup(specs, dev, args.precision)
    prec = grp.precision(0)
    torch.manual_seed(42)
    pinned_blocks = [torch.from_numpy(b.view(np.int16)).pin_memory() for b in blocks]
    dev_blocks = []
    for j, b in enumerate(blocks):
        phi = init_phi(dict(name="SIREN", coords_channel=3, data_channel=1, layers=plan["layers"], w0=plan["w0"],
                            features=plan["features"]))
        grp.set_axes(j, "-1,1")
        grp.set_params(j, pack_module_params(phi))
        t = pinned_blocks[j].to(dev, non_blocking=True)
        dev_blocks.append(t)
        vmin, vmax = float(b.min()), float(b.max())
        tau = (65535.0 - vmin) / (vmax - vmin) * 100.0
        grp.bind_volume(j, t, vmin, vmax, 0.0, 100.0, rules=[(65535, 65535, 1.0)], tau=tau, np_dtype="uint16")
        grp.set_sampler(j, "randomcube" if plan["full_block"] else "randompoint", plan["batch"])
    n_local = len(mine)
    samples_per_step_local = n_local * plan["batch"]
    fit_flops = sharding.fit_flops_per_sample(plan["features"], plan["layers"])
    fwd_flops = sharding.forward_flops_per_sample(plan["features"], plan["layers"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    opt_kw = dict(kind="Adamax", lr=1e-3, milestones=(50000, 60000, 70000), gamma=0.2, seed=42)

    # ---- device-resident timing: K steps, L2 flushed between steps, CUDA events around each step ----
    grp.fit_run(args.warmup, **opt_kw)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    reset_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s in range(args.steps):
        flush.fill_(s & 0xFF)
        ev[s][0].record()
        grp.fit_run(1, **opt_kw)
        ev[s][1].record()
    barrier()
    launches = launch_count()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    t_dev = sum(step_ms) / 1e3
    # back-to-back (no flush; the volume stays L2-resident as it does in a real fit): one event pair around K steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    grp.fit_run(args.steps, **opt_kw)
    e1.record()
    barrier()
    t_b2b = e0.elapsed_time(e1) / 1e3
    clk = clocks.stop()

    # ---- the dominant kernel alone (fit kernel = fit step without the optimiser launch) for the roofline ----
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_k = min(args.steps, 50)
    torch.cuda.synchronize()
    kt = 0.0
    for s in range(n_k):
        flush.fill_(s & 0xFF)
        k0.record()
        grp.fit_kernel_only(seed=42, step=s)
        k1.record()
        torch.cuda.synchronize()
        kt += k0.elapsed_time(k1)
    t_kernel = kt / n_k / 1e3

    # ---- end to end through the Python API with HOST buffers: per step H2D of the sampler's indices (the reference
    #      draws them with the CPU generator, main.py:156) and D2H of the per-block loss (main.py:401) ----
    e2e = None
    if not plan["full_block"]:
        gen = torch.Generator().manual_seed(42)
        host_idx = [torch.randint(0, int(np.prod(bs)), (n_local * plan["batch"],), generator=gen).pin_memory()
                    for _ in range(8)]
        dev_idx = torch.empty(n_local * plan["batch"], dtype=torch.int64, device=dev)
        host_loss = torch.empty(n_local, dtype=torch.float32).pin_memory()
        h2d, d2h = dev_idx.numel() * 8, n_local * 4
    else:
        host_idx, dev_idx = None, None
        host_loss = torch.empty(n_local, dtype=torch.float32).pin_memory()
        h2d, d2h = 0, n_local * 4

    # One C call per step (brief_fit_step_host): the step is ONE CUDA-graph launch — [h2d: step scalars + this step's
    # pinned indices] -> fit -> optimiser -> [d2h: per-block loss] — and the host reads every step's loss, one step
    # behind the GPU (the reference's loss.item() stalls the GPU every step, main.py:401).  Four pinned buffer sets cycle.
    n_sets = 4
    if host_idx is not None:
        host_idx = host_idx[:n_sets]
    loss_buf = [torch.zeros(n_local, dtype=torch.float32).pin_memory() for _ in range(n_sets)]
    loss_done = [torch.cuda.Event() for _ in range(n_sets)]
    loss_seen = [0.0]
    e2e_kw = dict(kind="Adamax", lr=1e-3, milestones=(50000, 60000, 70000), gamma=0.2, seed=42)

    def e2e_step(s):
        k = s % n_sets
        grp.fit_step_host(host_idx[k] if host_idx is not None else None, loss_buf[k], **e2e_kw)
        loss_done[k].record()
        if s > 0:
            loss_done[(s - 1) % n_sets].synchronize()
            loss_seen[0] = float(loss_buf[(s - 1) % n_sets][0])  # the host really consumes the value
        return loss_buf[k]

    for s in range(args.warmup):
        e2e_step(s)
    barrier()
    reset_launch_count()
    t0 = time.perf_counter()
    for s in range(args.steps):
        e2e_step(s)
    barrier()
    t_e2e = time.perf_counter() - t0
    e2e_launches = launch_count()
    final_loss = loss_buf[(args.steps - 1) % n_sets].clone()

    # ---- decompress of every local block (secondary metric) ----
    outs = grp.decompress("uint16")
    torch.cuda.synchronize()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dt_dec = 0.0
    n_dec = 5
    for s in range(n_dec):
        flush.fill_(s)
        d0.record()
        grp.decompress("uint16", out=outs)
        d1.record()
        torch.cuda.synchronize()
        dt_dec += d0.elapsed_time(d1)
    t_dec = dt_dec / n_dec / 1e3
    vox_local = n_local * int(np.prod(bs))
    # decompress e2e: host module parameters in, host uint16 volume out (block i's d2h copy under block i+1's decode)
    host_out = [torch.empty(bs, dtype=torch.int16).pin_memory() for _ in range(n_local)]
    params_host = [grp.get_params(j) for j in range(n_local)]
    grp.decompress_to_host("uint16", host_out=host_out, dev_out=outs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for j in range(n_local):
        grp.set_params(j, params_host[j])
    grp.decompress_to_host("uint16", host_out=host_out, dev_out=outs)
    torch.cuda.synchronize()
    t_dec_e2e = time.perf_counter() - t0

    # ---- data prologue (HBM-bound): min / max / sum / sum^2 of raw voxels, one launch (brief_block_stats) ----
    from brief_pytorch_b200.group import block_stats
    big = torch.empty(1 << 30, dtype=torch.int16, device=dev)  # 2 GiB of uint16 voxels: 16x the L2
    big.random_(0, 30000)
    block_stats([big], "uint16")
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_stats = 0.0
    for _ in range(5):
        s0.record()
        block_stats([big], "uint16")
        s1.record()
        torch.cuda.synchronize()
        t_stats += s0.elapsed_time(s1) / 1e3
    t_stats /= 5
    stats_gbs = big.numel() * 2 / t_stats / 1e9

    # ---- preprocess (utils/misc.py:244-254) on the device: threshold + 2x2x2 opening + zeroing, in place, on the same
    #      2 GiB buffer seen as one 1024^3 block whose lower half is dark (below the level) ----
    from brief_pytorch_b200.group import preprocess_
    vol3 = big.view(1024, 1024, 1024)
    vol3[:512] >>= 6                      # values < 469: a solid region the opening keeps
    pre_scratch = torch.empty(2 * 1024 * 1024 * 32 * 4, dtype=torch.uint8, device=dev)
    preprocess_(vol3, 500, [2, 2, 2], [0, 65535], "uint16", scratch=pre_scratch)
    zeroed = int((vol3[:512] == 0).sum()) + int((vol3[512:] == 0).sum())
    t_pre = 0.0
    for _ in range(5):
        s0.record()
        preprocess_(vol3, 500, [2, 2, 2], [0, 65535], "uint16", scratch=pre_scratch)
        s1.record()
        torch.cuda.synchronize()
        t_pre += s0.elapsed_time(s1) / 1e3
    t_pre /= 5
    pre_stats = {"voxels_per_s": big.numel() / t_pre, "ms": 1e3 * t_pre,
                 "gbs": (big.numel() * 2 + zeroed * 2) / t_pre / 1e9, "zeroed_voxels": zeroed,
                 "note": "brief_preprocess on a 1024^3 uint16 block (2 GiB, 16x L2), level 500, close [2,2,2], clip = dtype "
                         "range: 3 launches (mask, opening, apply); algorithmic bytes = the block read once + 2 B per "
                         "zeroed voxel written"}
    pre_stats["hbm_frac"] = pre_stats["gbs"] / pk["hbm_gbs"]
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import brief_oracle as O_pre  # cpu_baseline leg: the reference's host path timed beside the kernel
        samp = vol3[448:512, :256, :256].cpu().numpy().view(np.uint16)[..., None].copy()  # 64 x 256 x 256
        t0 = time.perf_counter()
        O_pre.preprocess(samp, 500, [2, 2, 2], [0, 65535])
        pre_stats["cpu_voxels_per_s"] = samp.size / (time.perf_counter() - t0)
        pre_stats["cpu_sample"] = "the reference's scipy.ndimage call on one 64x256x256 block (1 thread)"
    del big, vol3, pre_scratch

    # ---- the sampler's random gather (main.py:154-163) standalone: voxel indices -> (coords, normalised value, weight)
    #      from a volume far larger than L2, so that every sample costs one DRAM sector ----
    gshape = (256, 1024, 1024)  # 512 MiB of uint16 voxels, 4x the L2
    gvol = torch.empty(gshape, dtype=torch.int16, device=dev)
    gvol.random_(0, 30000)
    ggrp = SirenGroup([NetSpec(plan["features"], plan["layers"], plan["w0"], gshape)], dev, args.precision)
    ggrp.bind_volume(0, gvol, 0.0, 30000.0, 0.0, 100.0, rules=[(10001, 65535, 0.1)], np_dtype="uint16")
    n_g = 1 << 24
    gidx = torch.randint(0, gvol.numel(), (n_g,), dtype=torch.int64, device=dev)
    ggrp.gather(0, gidx)
    torch.cuda.synchronize()
    g0e, g1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_g = 0.0
    for _ in range(5):
        g0e.record()
        ggrp.gather(0, gidx)
        g1e.record()
        torch.cuda.synchronize()
        t_g += g0e.elapsed_time(g1e) / 1e3
    t_g /= 5
    gather_stats = {"samples_per_s": n_g / t_g,
                    "useful_gbs": n_g * (8 + 2 + 20) / t_g / 1e9,          # index + voxel read, 3 coords + value + weight written
                    "sector_gbs": n_g * (8 + 32 + 20) / t_g / 1e9,         # a random 2-byte read moves a 32-byte sector
                    "note": "brief_gather on 16 Mi random voxels of a 512 MiB uint16 volume (4x L2): coords + normalised value "
                            "+ weight materialised like the reference sampler; the fit kernels fuse this gather instead"}
    gather_stats["hbm_frac_sector"] = gather_stats["sector_gbs"] / pk["hbm_gbs"]
    ggrp.close()
    del gvol, gidx

    # ---- deblocking post-filter (deblock.cpp:226-321) on a decoded 512^3 uint16 volume cut into 8x8x8 = 512 blocks ----
    deblock_stats = None
    if world == 1 and not args.no_side_legs:
        from brief_pytorch_b200.deblock import deblock_, seam_masks
        dvol = torch.empty((512, 512, 512), dtype=torch.int16, device=dev)
        dvol.random_(0, 30000)
        names = [f"d_{z}_{z + 63}-h_{y}_{y + 63}-w_{x}_{x + 63}" for z in range(0, 512, 64) for y in range(0, 512, 64) for x in range(0, 512, 64)]
        masks = seam_masks(names)
        seam_px = sum(64 * 64 * bin(m & 15).count("1") for m in masks)  # 64 slices x 64 pixels per listed seam of a block
        deblock_(dvol, names)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            deblock_(dvol, names)
        torch.cuda.synchronize()
        t_db = (time.perf_counter() - t0) / 3
        deblock_stats = {"ms": 1e3 * t_db, "blocks": len(names), "seam_pixels": seam_px, "seam_pixels_per_s": seam_px / t_db,
                         "gbs": seam_px * 20 / t_db / 1e9, "hbm_frac": seam_px * 20 / t_db / 1e9 / pk["hbm_gbs"],
                         "note": "brief_deblock on a 512^3 uint16 volume, 512 blocks of 64^3 in the reference's listing order; "
                                 "algorithmic bytes = 12 B read + 8 B written per seam pixel (6-tap read, 4-tap write); the launch is "
                                 "bound by the reference's seam ORDER (seams cross and are filtered in place: one CTA per z slice "
                                 "walks its seams with a barrier between them), not by HBM"}
        del dvol

    # ---- NCCL, after the hot path: decoded blocks -> rank 0 (variable-size send/recv), checked by a checksum table ----
    gather = None
    if world > 1:
        local_dec = {b: outs[j] for j, b in enumerate(mine)}
        shapes = [bs] * n_total
        sums = torch.stack([o.view(torch.int16).to(torch.int64).sum() for o in outs]).to(torch.float64).reshape(-1, 1)
        want = sharding.gather_block_stats(sums, owner)
        sharding.gather_blocks(local_dec, owner, shapes, dst=0)  # first call pays NCCL's lazy peer connections
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        got = sharding.gather_blocks(local_dec, owner, shapes, dst=0)
        g1.record()
        barrier()
        t_gather = g0.elapsed_time(g1) / 1e3
        if rank == 0:
            ok = all(float(got[b].view(torch.int16).to(torch.int64).sum()) == float(want[b, 0]) for b in range(n_total))
            nbytes = sum(int(np.prod(bs)) * 2 for b in range(n_total) if owner[b] != 0)
            gather = {"ms": 1e3 * t_gather, "bytes_into_rank0": nbytes, "gbs": nbytes / t_gather / 1e9, "checksums_ok": ok}
        # the same exchange at >= 1 GiB into the assembling rank (every block listed `rep` times): the NVLink-rate figure
        rep = max(1, int(np.ceil((1.15 * (1 << 30)) / max(1, (world - 1) * n_local * int(np.prod(bs)) * 2))))
        big_owner = [owner[b] for b in range(n_total) for _ in range(rep)]
        big_local = {b * rep + i: outs[j] for j, b in enumerate(mine) for i in range(rep)}
        big_shapes = [bs] * (n_total * rep)
        sharding.gather_blocks(big_local, big_owner, big_shapes, dst=0, dtype=torch.int16)
        barrier()
        g0.record()
        got_big = sharding.gather_blocks(big_local, big_owner, big_shapes, dst=0, dtype=torch.int16)
        g1.record()
        barrier()
        if rank == 0:
            nb_big = sum(int(np.prod(bs)) * 2 for b in big_owner if b != 0)
            gather["large"] = {"ms": g0.elapsed_time(g1), "bytes_into_rank0": nb_big, "gbs": nb_big / (g0.elapsed_time(g1) / 1e3) / 1e9,
                               "note": "one packed payload per sending rank, one grouped ncclSend/Recv; peer-copy reference 770 GB/s per direction"}
        del got_big
        del got

    # ---- one short measurement per block geometry of the BASELINE configs (rank 0's GPU; N = 1 run only) ----
    workloads = None
    if world == 1 and not args.no_side_legs:
        workloads = {}
        for name in sorted(WORKLOADS):
            if name == args.workload:
                continue
            try:
                workloads[name] = quick_workload(name, dev, args.precision, pk, flush)
            except Exception as e:  # a geometry that cannot run must show up in the line, not abort the headline
                workloads[name] = {"error": str(e)[:200]}

    # ---- strong scaling: ONE 1024^3 neuron volume (BASELINE config 3, 512 blocks 128^3) sharded over the ranks by LPT ----
    strong = None
    if not args.no_side_legs:
        from types import SimpleNamespace
        strong = measure_named(NAMED["neuron1024"], SimpleNamespace(steps=10, warmup=3, precision=args.precision,
                                                                    reproducible=args.reproducible), dev, rank, world, local)

    # ---- reduce over ranks (max time), gather per-block stats (the only collective; outside the timed region) ----
    times = torch.tensor([t_dev, t_b2b, t_e2e, t_kernel, t_dec, t_dec_e2e], dtype=torch.float64, device=dev)
    counts = torch.tensor([samples_per_step_local, vox_local, launches], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        table = sharding.gather_block_stats(final_loss.to(dev).reshape(-1, 1), owner)
    else:
        table = final_loss.reshape(-1, 1)
    t_dev, t_b2b, t_e2e, t_kernel, t_dec, t_dec_e2e = [float(x) for x in times.tolist()]
    samples_per_step, vox_total, launches_total = [float(x) for x in counts.tolist()]

    if rank == 0:
        value = samples_per_step * args.steps / t_dev
        tf_kernel = samples_per_step_local * fit_flops / t_kernel / 1e12
        peak_tf = pk["bf16_tflops"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": prec if prec == "f16" else "f32", "data": "synthetic",
            "config": fit_config(plan, n_local, prec, world),
            "back_to_back": {"value": samples_per_step * args.steps / t_b2b, "unit": UNIT,
                             "ms_per_step": 1e3 * t_b2b / args.steps, "note": "no L2 flush, one event pair around K steps"},
            "e2e": {"value": samples_per_step * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * t_e2e / args.steps,
                    "gpu_launches": int(e2e_launches),
                    "note": "per step ONE call SirenGroup.fit_step_host: pinned host sampler indices + step scalars -> device on "
                            "a copy stream (under the previous step's kernels), then one CUDA-graph launch: fit kernel -> optimiser "
                            "kernel -> per-block loss -> pinned host; the host reads every step's loss, one step behind the GPU"},
            "gpu_launches": int(launches_total),
            "roofline": {"bound": "tensor", "achieved": tf_kernel, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": tf_kernel / peak_tf,
                         "traffic": ncu_traffic(args.workload) if (world == 1 and prec == "f16") else None, "peak_src": pk["src"] + " bf16 burst",
                         "kernel": "fit (gather+fwd+loss+bwd)", "kernel_ms": 1e3 * t_kernel,
                         "flops_per_sample": fit_flops, "samples_per_launch": samples_per_step_local},
            "decompress": {"value": vox_total / t_dec, "unit": "voxels/s", "ms": 1e3 * t_dec,
                           "e2e_value": vox_total / t_dec_e2e,
                           "hbm_gbs": vox_total / world * 2 / t_dec / 1e9, "hbm_frac": vox_total / world * 2 / t_dec / 1e9 / pk["hbm_gbs"],
                           "tflops": vox_total / world * fwd_flops / t_dec / 1e12,
                           "tensor_frac": vox_total / world * fwd_flops / t_dec / 1e12 / peak_tf},
            "final_loss_mean": float(table.mean()), "clocks": clk,
        }
        line["block_stats"] = {"gbs": stats_gbs, "hbm_frac": stats_gbs / pk["hbm_gbs"], "bytes": 2 << 30,
                               "note": "min/max/sum/sum^2 of a 2 GiB uint16 buffer, one launch incl. its 2 tiny copies; "
                                       "rank 0's figure; peak = " + pk["src"] + " copy bandwidth (read+write)"}
        line["sampler_gather"] = gather_stats
        line["preprocess"] = pre_stats
        if deblock_stats is not None:
            line["deblock"] = deblock_stats
        if gather is not None:
            line["gather_decoded_blocks"] = gather
        if workloads is not None:
            workloads[args.workload] = {"value": value, "unit": UNIT, "ms_per_step": 1e3 * t_dev / args.steps, "blocks": n_local,
                                        "block_shape": list(bs), "features": plan["features"], "layers": plan["layers"],
                                        "batch_per_block": plan["batch"], "precision": prec, "fit_kernel_ms": 1e3 * t_kernel,
                                        "fit_tflops": tf_kernel, "roofline_frac": tf_kernel / peak_tf,
                                        "decompress_voxels_per_s": vox_total / t_dec}
            line["workloads"] = workloads
        if strong is not None:
            line["strong_scaling"] = {k: strong[k] for k in ("value", "unit", "ms_per_step", "n_gpus", "config", "per_rank", "imbalance",
                                                             "decompress", "roofline") if k in strong}
            if "param_checksum" in strong:
                line["strong_scaling"]["param_checksum"] = strong["param_checksum"]
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            step = oracle_fit_rate(plan, args.cpu_seconds, threads)
            step(); step(); step()
            t0 = time.perf_counter()
            n = 0
            while time.perf_counter() - t0 < args.cpu_seconds and n < 200:
                step(); n += 1
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": plan["batch"] * n / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{n} steps of one block (f={plan['features']}, batch {plan['batch']}) "
                                              f"with the oracle's torch-CPU restatement of main.py:385-400"}
            try:
                line["gpu_eager_baseline"] = gpu_eager_rate(plan, 4.0, dev)
            except Exception as e:
                line["gpu_eager_baseline"] = {"error": str(e)[:200]}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
