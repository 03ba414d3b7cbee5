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
        "config": {"workload": plan["desc"], "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="vessel", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="auto", choices=["auto", "f16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
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

    # step s+1's indices cross PCIe on a copy stream while step s computes (two device buffers); every step still moves
    # its own h2d bytes and reads its own loss back before the next one starts
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    idx_buf = [torch.empty_like(dev_idx) for _ in range(2)] if host_idx is not None else None
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(s):
        if host_idx is None:
            return
        k = s & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k])
            idx_buf[k].copy_(host_idx[s % 8], non_blocking=True)
            ready[k].record(copy_stream)

    loss_buf = [torch.empty(n_local, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_done = [torch.cuda.Event() for _ in range(2)]
    loss_seen = [0.0]

    def e2e_step(s):
        """Step s is enqueued (indices h2d -> fit -> optimiser -> loss d2h), THEN the host waits for and reads the loss
        of step s-1: every step's loss reaches the host, one step behind the GPU, so the launch latency of the next
        step is not exposed (the reference's loss.item() stalls the GPU every step, main.py:401)."""
        k = s & 1
        prefetch(s + 1)
        if host_idx is not None:
            main_stream.wait_event(ready[k])
        loss = grp.fit_step(idx_buf[k] if host_idx is not None else None, seed=42, step=s)
        grp.opt_step("Adamax", 1e-3)
        consumed[k].record(main_stream)
        loss_buf[k].copy_(loss, non_blocking=True)
        loss_done[k].record(main_stream)
        if s > 0:
            loss_done[k ^ 1].synchronize()
            loss_seen[0] = float(loss_buf[k ^ 1][0])  # the host really consumes the value
        return loss_buf[k]

    for k in range(2):
        consumed[k].record(main_stream)
    prefetch(0)
    for s in range(args.warmup):
        e2e_step(s)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        e2e_step(s)
    barrier()
    t_e2e = time.perf_counter() - t0
    final_loss = loss_buf[(args.steps - 1) & 1].clone()

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
        del got

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
            "config": {"workload": plan["desc"], "blocks_per_gpu": n_local, "block_shape": list(bs),
                       "features": plan["features"], "layers": plan["layers"], "batch_per_block": plan["batch"],
                       "precision": prec, "l2": "flushed (256 MiB write) between timed steps",
                       "parallelism": f"blocks sharded by LPT over {world} GPU(s), no collective on the fit path"},
            "back_to_back": {"value": samples_per_step * args.steps / t_b2b, "unit": UNIT,
                             "ms_per_step": 1e3 * t_b2b / args.steps, "note": "no L2 flush, one event pair around K steps"},
            "e2e": {"value": samples_per_step * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * t_e2e / args.steps,
                    "note": "per step: pinned host sampler indices -> device (copy stream, one step ahead), fit_step + "
                            "opt_step via the Python API, per-block loss -> pinned host buffer every step, read by the host one "
                            "step behind the GPU"},
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
        if gather is not None:
            line["gather_decoded_blocks"] = gather
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
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
