"""Synthetic uint16 volumes of the shapes the BASELINE configs name (the example TIFFs of vessel/neuron/hipct are
missing from the reference checkout).  numpy default_rng(seed); cheap enough to build per run."""
from __future__ import annotations

import numpy as np


def _axes(shape):
    return np.meshgrid(*[np.linspace(-1, 1, n, dtype=np.float32) for n in shape], indexing="ij", sparse=True)


def vessel(shape=(64, 512, 512), seed=42) -> np.ndarray:
    """Dark noisy background with smooth bright tubular structures (peak ~30000)."""
    rng = np.random.default_rng(seed)
    z, y, x = _axes(shape)
    vol = np.full(shape, 200.0, dtype=np.float32)
    vol += rng.normal(0, 50, shape).astype(np.float32)
    for _ in range(12):
        a, b, c, d = rng.uniform(-1, 1, 4)
        fy, fx = rng.uniform(1.0, 4.0, 2)
        r = rng.uniform(0.02, 0.06)
        dist2 = (y - (a + 0.4 * np.sin(fx * x + 3 * z + b))) ** 2 + (z - (c * 0.6 + 0.3 * np.cos(fy * x + d))) ** 2
        vol += 30000.0 * np.exp(-dist2 / (2 * r * r))
    return np.clip(vol, 0, 65535).astype(np.uint16)[..., None]


def neuron(shape=(128, 128, 128), seed=42) -> np.ndarray:
    """Sparse thin bright filaments + a soma on a background below 10000 (both weight classes of
    `value_10001_65535_0.1`, opt/DivideTask/neuron.yaml:35)."""
    rng = np.random.default_rng(seed)
    z, y, x = _axes(shape)
    vol = 1500.0 + rng.normal(0, 300, shape).astype(np.float32)
    for _ in range(8):
        a, b, c, d = rng.uniform(-1, 1, 4)
        dist2 = (y - (a + 0.5 * np.sin(2.5 * x + b))) ** 2 + (z - (c + 0.5 * np.cos(2.0 * x + d))) ** 2
        vol += 45000.0 * np.exp(-dist2 / (2 * 0.015 ** 2))
    vol += 50000.0 * np.exp(-((x - 0.1) ** 2 + (y + 0.2) ** 2 + (z - 0.3) ** 2) / (2 * 0.08 ** 2))
    return np.clip(vol, 0, 65535).astype(np.uint16)[..., None]


def hipct(shape=(128, 128, 128), seed=42) -> np.ndarray:
    """Dense band-limited texture with region-to-region variance contrast (non-uniform `by_var` allocation)."""
    rng = np.random.default_rng(seed)
    z, y, x = _axes(shape)
    vol = np.zeros(shape, dtype=np.float32)
    for octave in range(4):
        k = 2.0 ** octave * np.pi
        pz, py, px = rng.uniform(0, 2 * np.pi, 3)
        vol += (0.5 ** octave) * np.sin(k * z + pz) * np.sin(k * y + py) * np.sin(k * x + px)
    contrast = 0.25 + 0.75 * (0.5 + 0.5 * np.tanh(3 * (x + y * 0.5)))
    vol = 30000.0 + 14000.0 * vol * contrast + rng.normal(0, 400, shape).astype(np.float32)
    return np.clip(vol, 0, 65535).astype(np.uint16)[..., None]
