"""profiles/r02_scaling.txt from the committed named-shape bench lines (profiles/r02_bench_<workload>_n<N>.json):
    python tools/scaling_table.py > profiles/r02_scaling.txt"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W = ["neuron1024", "neuron1024_nb4", "hipct2048_equal", "hipct2048", "decomp4096"]
print("Named shapes of BASELINE.json configs 3-5: ONE synthetic volume, blocks sharded by parameter-weighted LPT over N B200 of one box")
print("(strong scaling; no collective on the fit or decode path; device-timed, max over ranks; reproducible per-network slicing;")
print(" param_checksum = XOR of crc32(fitted parameters, block id) over all blocks after the same 13 steps: equal across N = bit-identical blocks)\n")
for w in W:
    base = None
    for n in (1, 2, 4, 8):
        p = os.path.join(ROOT, "profiles", f"r02_bench_{w}_n{n}.json")
        if not os.path.exists(p):
            continue
        d = json.load(open(p))
        if n == 1 or base is None:
            base = (n, d["value"], d["decompress"]["value"])
            print(f"{w}: {d['config']['workload']}")
            print(f"    blocks {d['config']['blocks']}, block {d['config']['block_shape']}, widths min/median/max {d['config']['features_min_median_max']}")
        fit = d["metric"].startswith("siren_fit")
        eff = d["value"] / base[1] / (n / base[0])
        deff = d["decompress"]["value"] / base[2] / (n / base[0])
        line = f"    N={n}: "
        if fit:
            line += f"fit {d['value'] / 1e9:7.3f} G coord-samples/s ({d['ms_per_step']:8.3f} ms/step, {d['roofline']['achieved']:7.1f} TFLOP/s = {d['roofline']['frac'] * 100:4.1f} % of sustained bf16 peak x N, eff {eff:4.2f}, rank max/mean {d['imbalance']['fit_max_over_mean']:.3f}); "
        line += f"decompress {d['decompress']['value'] / 1e9:7.2f} Gvox/s ({d['decompress']['ms']:9.1f} ms, {d['decompress']['hbm_gbs']:6.1f} GB/s = {d['decompress']['hbm_frac'] * 100:4.2f} % of HBM x N, eff {deff:4.2f}, rank max/mean {d['imbalance']['decompress_max_over_mean']:.3f})"
        if d.get("param_checksum") and fit:
            line += f"; param_checksum {d['param_checksum']}"
        print(line)
    print()
