"""Shared helpers of the parity tests (oracle side only - never imported by the product)."""
import hashlib
import os

import numpy as np
import torch

from cmr_agent_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED = 2023

ENV_CASES = {
    "env_small": (2, dict(num_pt=4096, img_h=64, img_w=256), 4, True),
    "env_ragged": (2, dict(num_pt=1531, img_h=36, img_w=100), 3, True),
    "env_kitti": (2, dict(num_pt=40960, img_h=160, img_w=512), 4, False),
    "env_nuscenes": (1, dict(num_pt=40960, img_h=160, img_w=320, unique=(26000, 34000)), 3, False),
}


def sha(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(t.numpy()).tobytes())
    return h.hexdigest()


def golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)


def env_inputs(name):
    """Regenerate the bulk inputs of a golden environment case and check them against its sha256."""
    batch, shape, iters, full = ENV_CASES[name]
    g = golden(name)
    data = synth.make_batch(batch, seed=SEED, **shape)
    got = sha(data["pc"], data["pc_geo_feat"], data["img_geo_feat"], data["pc_overlap_pred"], data["pc_mask"],
              data["pc_in_cam_space"], data["K"])
    assert got == str(g["inputs_sha"]), "synthetic inputs do not regenerate bit-identically on this host"
    return data, g, iters, full, shape


def to_device(data, device):
    """What the reference's feature network leaves on the device; K, P,
    pc_in_cam_space and pc_mask stay on the CPU as the dataset delivers them."""
    out = dict(data)
    for k in ("pc", "pc_overlap_pred", "pc_geo_feat", "img_geo_feat"):
        out[k] = data[k].to(device)
    return out


def rel_err(a, b, floor=1e-7):
    """max |a-b| / max(|a|,|b|) with an absolute floor."""
    a = a.double()
    b = b.double()
    denom = torch.maximum(a.abs(), b.abs()).clamp_min(floor)
    d = (a - b).abs()
    d = torch.where(d <= floor, torch.zeros_like(d), d)
    return float((d / denom).max()) if d.numel() else 0.0
