"""Compressed-parameter layout — drop-in for the reference's utils/ModelSave.py:8-60.

A saved module is a directory with, per layer l, two header-less native-endian fp32 files:
`weight-{l}-{out}-{in}` (row-major [out][in]) and `bias-{l}-{n}`; shapes are recovered from the file
names alone.  Files written here are byte-identical to the reference's for equal parameters.
"""
import os
import shutil

import numpy as np
import torch


def save_model(model, save_path: str, devive: str = "cpu") -> None:  # 'devive' [sic]: reference keyword
    if not hasattr(model, "net"):
        torch.save(model.state_dict(), save_path)
        return
    if os.path.exists(save_path):
        shutil.rmtree(save_path)  # the reference also replaces the whole directory
    os.mkdir(save_path)
    for l, block in enumerate(model.net):
        lin = block[0]
        w = lin.weight.detach().to("cpu", torch.float32).contiguous().numpy()
        b = lin.bias.detach().to("cpu", torch.float32).contiguous().numpy()
        w.astype("=f4", copy=False).tofile(os.path.join(save_path, f"weight-{l}-{w.shape[0]}-{w.shape[1]}"))
        b.astype("=f4", copy=False).tofile(os.path.join(save_path, f"bias-{l}-{b.shape[0]}"))


def load_model(model, model_path: str, device: str = "cpu"):
    if not hasattr(model, "net"):
        model.load_state_dict(torch.load(model_path))
        return model
    for name in os.listdir(model_path):
        kind, *dims = name.split("-")
        if kind not in ("weight", "bias"):
            continue
        dims = [int(x) for x in dims]
        arr = np.fromfile(os.path.join(model_path, name), dtype="=f4")
        lin = model.net[dims[0]][0]
        with torch.no_grad():
            if kind == "weight":
                lin.weight.data = torch.from_numpy(arr.reshape(dims[1], dims[2]).copy()).to(device)
            else:
                lin.bias.data = torch.from_numpy(arr[:dims[1]].copy()).to(device)
    return model


def CopyDir(old_dir: str, new_dir: str) -> None:
    os.makedirs(new_dir, exist_ok=True)
    for name in os.listdir(old_dir):
        shutil.copy(os.path.join(old_dir, name), os.path.join(new_dir, name))
