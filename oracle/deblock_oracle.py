"""Deblocking oracle (TEST INFRASTRUCTURE ONLY): runs the reference's own deblock.cpp, compiled unmodified from
/root/reference by oracle/build_ref.sh against a stand-in tiffio.h (oracle/deblock_ref/), on a volume and a block
partition, and returns the filtered volume together with the order in which the reference saw the blocks
(`readdir` order of compressed/module — the filter is order dependent: seams are processed sequentially in place and
the reference's duplicate-seam flags are sticky, deblock.cpp:245-275).

Only tests/ and oracle/gen_golden_deblock.py import this module."""
import os
import shutil
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BINARY = os.path.join(HERE, "_ref", "deblock_ref")


def available() -> bool:
    return os.path.exists(BINARY)


def _write_raw(path, vol):
    d, h, w = vol.shape
    with open(path, "wb") as fh:
        fh.write(b"BRIEFRAW" + struct.pack("<III", d, h, w))
        fh.write(np.ascontiguousarray(vol, dtype="<u2").tobytes())


def _read_raw(path):
    with open(path, "rb") as fh:
        assert fh.read(8) == b"BRIEFRAW"
        d, h, w = struct.unpack("<III", fh.read(12))
        return np.frombuffer(fh.read(), dtype="<u2").reshape(d, h, w).copy()


def block_name(z1, z2, y1, y2, x1, x2) -> str:
    """Chunk directory name of main.py:589-607 / utils/misc.py:366 (inclusive ends)."""
    return f"d_{z1}_{z2}-h_{y1}_{y2}-w_{x1}_{x2}"


def run_reference(vol: np.ndarray, names):
    """vol: uint16 [D,H,W]; names: chunk directory names.  Returns (filtered volume, names in the reference's order)."""
    assert vol.dtype == np.uint16 and vol.ndim == 3
    d, h, w = vol.shape
    step = tempfile.mkdtemp(prefix="brief_deblock_")
    try:
        os.makedirs(os.path.join(step, "decompressed"))
        mod = os.path.join(step, "compressed", "module")
        os.makedirs(mod)
        for n in names:
            os.makedirs(os.path.join(mod, n))
        _write_raw(os.path.join(step, "decompressed", f"vol-0_{d}-0_{h}-0_{w}_decompressed.tif"), vol)
        order = os.listdir(mod)  # same getdents order the reference's readdir() loop sees
        subprocess.run([BINARY, step], check=True, stdout=subprocess.DEVNULL)
        assert os.listdir(mod) == order
        out = _read_raw(os.path.join(step, "deblock", f"vol-0_{d}-0_{h}-0_{w}_decompressed_deblocked_c++.tif"))
        return out, order
    finally:
        shutil.rmtree(step, ignore_errors=True)


# ------------------------------------------------------------------------------------------------------------------
# Restatement of deblock.cpp in numpy (each step cites the line it follows); pinned against the compiled reference by
# tests/golden/deblock.npz (tests/test_deblock.py, CPU).
# ------------------------------------------------------------------------------------------------------------------
def parse_block(name: str):
    """deblock.cpp:250-253: 'd_z1_z2-h_y1_y2-w_x1_x2' -> (z1, z2, y1, y2, x1, x2), inclusive ends."""
    parts = [p.split("_") for p in name.split("-")]
    return tuple(int(v) for p in parts for v in p[1:3])


def plan_lines(names):
    """deblock.cpp:244-276: the seam list, in order.  A block contributes its left / right / down / up seam for every z
    of its range unless the corresponding flag is set; a flag is raised when the block's seam AT z1 equals a seam already
    listed — and is NEVER lowered again (the flags live outside the block loop).  Returns per block the 4-bit mask of
    seams it contributes (bit 0 left, 1 right, 2 down, 3 up)."""
    seen = set()
    flags = [False, False, False, False]
    masks = []
    for name in names:
        z1, z2, y1, y2, x1, x2 = parse_block(name)
        cand = [(x1, x1, y1, y2), (x2, x2, y1, y2), (x1, x2, y1, y1), (x1, x2, y2, y2)]  # (l, r, d, u)
        for k in range(4):
            if (z1,) + cand[k] in seen:
                flags[k] = True
        mask = 0
        for k in range(4):
            if not flags[k]:
                mask |= 1 << k
                for z in range(z1, z2 + 1):
                    seen.add((z,) + cand[k])
        masks.append(mask)
    return masks


def _cdiv(a, b):
    """C integer division (truncation toward zero) on int64 arrays."""
    return np.sign(a) * (np.abs(a) // b)


def _alpha(x):  # deblock.cpp:13-16 (double arithmetic, float return)
    return np.float32(0.8 * (2.0 ** (np.float32(x) / np.float32(6)) - 1))


def _beta(x):   # deblock.cpp:18-21
    return np.float32(0.5 * x - 7)


def _filter_line(p2, p1, p0, q0, q1, q2, index_a, index_b, thres):
    """judge_filter (deblock.cpp:33-41) + filter (:43-71) for the six taps of every pixel of one seam (int64 arrays).
    Returns the new (p1, p0, q0, q1) as uint16 values (wrap-around like the compiled float -> uint16 conversion)."""
    al, be = _alpha(index_a), _beta(index_b)
    on = ((p1 + p0 + q0 + q1) // 4 <= thres) & (np.abs(p0 - q0).astype(np.float32) < al) & \
         (np.abs(p1 - p0).astype(np.float32) < be) & (np.abs(q1 - q0).astype(np.float32) < be)
    d0 = _cdiv(4 * (q0 - p0) + (p1 - q1) + 4, 8).astype(np.float32)
    dp1 = _cdiv(p2 + (p0 + q0 + 1) // 2 - 2 * p1, 2).astype(np.float32)
    dq1 = _cdiv(q2 + (q0 + p0 + 1) // 2 - 2 * q1, 2).astype(np.float32)
    c1 = np.float32(20)
    c0 = 20 + (np.abs(p2 - p0).astype(np.float32) < be).astype(np.int64) + (np.abs(q2 - q0).astype(np.float32) < be).astype(np.int64)
    c0 = c0.astype(np.float32)
    d0 = np.minimum(np.maximum(d0, -c0), c0)
    dp1 = np.minimum(np.maximum(dp1, -c1), c1)
    dq1 = np.minimum(np.maximum(dq1, -c1), c1)
    wrap = lambda v: (v.astype(np.float32).astype(np.int64)) & 0xFFFF   # noqa: E731  (uint16)(int)float
    n_p1 = wrap(p1.astype(np.float32) + dp1)
    n_p0 = wrap(p0.astype(np.float32) + d0)
    n_q0 = wrap(q0.astype(np.float32) - d0)
    n_q1 = wrap(q1.astype(np.float32) + dq1)
    return on, n_p1, n_p0, n_q0, n_q1


def deblock_restated(vol: np.ndarray, names, index_a=51, index_b=2000, thres=65535) -> np.ndarray:
    """deblock.cpp:226-321 on a uint16 [D,H,W] volume with the blocks in the given (readdir) order."""
    img = vol.astype(np.int64).copy()
    D_, H, W = img.shape
    blocks = [parse_block(n) for n in names]
    masks = plan_lines(names)
    # the reference walks its seam list in insertion order: block by block, z by z, (left, right, down, up)
    for (z1, z2, y1, y2, x1, x2), mask in zip(blocks, masks):
        cand = [(x1, x1, y1, y2), (x2, x2, y1, y2), (x1, x2, y1, y1), (x1, x2, y2, y2)]
        for z in range(z1, z2 + 1):
            for k in range(4):
                if not (mask >> k) & 1:
                    continue
                l, r, d, u = cand[k]
                if l == r and (l - 3 < 0 or l + 3 > W - 1):      # :284
                    continue
                elif d == u and (d - 3 < 0 or d + 3 > H - 1):    # :286
                    continue
                if l == r:                                        # :292 vertical seam at column l, rows d..u
                    ys, x = np.arange(d, u + 1), l
                    taps = [img[z, ys, x + o] for o in (-3, -2, -1, 0, 1, 2)]
                    on, a, b, c, e = _filter_line(*taps, index_a, index_b, thres)
                    for o, v in zip((-2, -1, 0, 1), (a, b, c, e)):
                        img[z, ys[on], x + o] = v[on]
                elif d == u:                                      # :304 horizontal seam at row d, columns l..r
                    xs, y = np.arange(l, r + 1), d
                    taps = [img[z, y + o, xs] for o in (-3, -2, -1, 0, 1, 2)]
                    on, a, b, c, e = _filter_line(*taps, index_a, index_b, thres)
                    for o, v in zip((-2, -1, 0, 1), (a, b, c, e)):
                        img[z, y + o, xs[on]] = v[on]
    return img.astype(np.uint16)
