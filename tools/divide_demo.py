"""End-to-end DivideTask flow on 1..N GPUs (one process per GPU), the in-process replacement of the reference's
per-block subprocess farm (main.py:509-651):

    torchrun --nproc-per-node N tools/divide_demo.py [--nb 16] [--steps 300] [--shape 32,128,128] [--out /tmp/brief_demo]

Every rank partitions the same synthetic vessel volume, fits its LPT share of the blocks in one SirenGroup and writes
its part of the reference-layout compressed/ directory; rank 0 then decodes the whole directory (decompress_divide),
and reports PSNR / SSIM against the original plus the all-gathered per-block loss table."""
import argparse, json, os, shutil, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import yaml
from brief_pytorch_b200 import misc, sharding, synth
from brief_pytorch_b200.CompressFramework import NFGR

OPT = yaml.safe_load("""
Name: NFGR
Compress:
  divide: {divide_type: adaptotal_-1_-1_-1_16, param_alloc: by_size, param_size_thres: 26, exception: none}
  half: false
  sampler: {name: randomcube, cube_count: 1, cube_len: [10000000, 10000000, 10000000], sample_size: 100000}
  coords_mode: -1,1
  preprocess: {denoise: {level: 0, close: [2, 2, 2]}, clip: [0, 65535]}
  param: {init_net_path: none, filesize_ratio: 128, given_size: 0}
  loss: {name: datal2, beta: 0.01, weight: [value_65535_65535_1], weight_thres: 65535}
  max_steps: 80000
  checkpoints: none
  lr_phi: 0.001
  optimizer_name_phi: Adamax
  lr_scheduler_phi: {name: MultiStepLR, milestones: [50000, 60000, 70000], gamma: 0.2}
Decompress:
  sample_size: 10000
  postprocess: {denoise: {level: 0, close: [2, 2, 2]}, clip: [0, 65535]}
Module:
  phi: {coords_channel: 3, data_channel: 1, layers: 7, name: SIREN, w0: 10, output_act: false, res: false}
Normalize: {name: minmaxany_0_100}
""")

ap = argparse.ArgumentParser()
ap.add_argument("--nb", type=int, default=16)
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--shape", default="32,128,128")
ap.add_argument("--ratio", type=float, default=32)
ap.add_argument("--out", default="/tmp/brief_demo")
ap.add_argument("--data", default="vessel", choices=["vessel", "hipct", "neuron"])
ap.add_argument("--alloc", default="by_size", choices=["by_size", "by_var"])
ap.add_argument("--sampler", default="randomcube")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
OPT["Compress"]["divide"]["divide_type"] = f"adaptotal_-1_-1_-1_{args.nb}"
OPT["Compress"]["param"]["filesize_ratio"] = args.ratio
OPT["Compress"]["divide"]["param_alloc"] = args.alloc
OPT["Compress"]["sampler"]["name"] = args.sampler
shape = tuple(int(x) for x in args.shape.split(","))
vol = getattr(synth, args.data)(shape, seed=42)
cdir = os.path.join(args.out, "compressed")
if rank == 0:
    shutil.rmtree(args.out, ignore_errors=True)
    os.makedirs(cdir)
if world > 1:
    dist.barrier()
cf = NFGR(OPT, local, "auto")
torch.cuda.synchronize(); t0 = time.perf_counter()
blocks, mine = cf.compress_divide(vol, cdir, max_steps=args.steps, rank=rank, world=world)
torch.cuda.synchronize(); t_fit = time.perf_counter() - t0
owner = [None] * len(blocks)
for r in range(world):
    for i in sharding.my_blocks(sharding.lpt_assign([1.0] * len(blocks), world), r):
        owner[i] = r
local_loss = torch.tensor([[blocks[i].loss] for i in mine], dtype=torch.float32, device="cuda")
costs_owner = sharding.lpt_assign([1.0] * len(blocks), world)
table = sharding.gather_block_stats(local_loss, costs_owner) if world > 1 else local_loss
if world > 1:
    dist.barrier()
if rank == 0:
    t0 = time.perf_counter()
    dec = cf.decompress_divide(os.path.join(cdir, "sideinfos.yaml"), os.path.join(cdir, "module"), os.path.join(cdir, "sideinfos"))
    t_dec = time.perf_counter() - t0
    perf = misc.eval_performance(args.steps, vol, dec, device="cuda")  # utils/misc.py:477-499 on the device (brief_volume_quality)
    files = sum(len(f) for _, _, f in os.walk(cdir))
    nbytes = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(os.path.join(cdir, "module")) for f in fs)
    print(json.dumps({"world": world, "shape": shape, "blocks": len(blocks), "features": sorted({b.features for b in blocks}), "data": args.data, "alloc": args.alloc, "steps": args.steps,
                      "fit_s": round(t_fit, 3), "decode_s": round(t_dec, 3), "module_bytes": nbytes, "files": files,
                      "ratio_actual": round(vol.nbytes / nbytes, 2),
                      "psnr_db": round(float(perf["psnr"]), 3), "ssim": round(float(perf["ssim"]), 5),
                      "loss_mean": round(float(table.mean()), 4), "loss_table_rows": int(table.shape[0])}))
if world > 1:
    dist.destroy_process_group()
