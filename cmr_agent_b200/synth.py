"""Synthetic KITTI-/NuScenes-shaped episode batches.

There is no network and no dataset in this environment, so every benchmark and
test runs on data of the *shape and distribution* the reference's datasets
produce (/root/reference/dataset/KittiDataset.py:238-253,290-334,352-422 and
/root/reference/dataset/NuScenesDataset.py:100-111), generated here with a CPU
``torch.Generator`` seeded ``seed + episode`` (config seed 2023,
/root/reference/config/KittiConfig.py:30).  Everything is returned on the CPU in
the layout the reference's ``data`` dict uses; callers
move to the device what the feature network would have produced there.
"""
import math

import torch

KITTI = dict(num_pt=40960, img_h=160, img_w=512, channels=64, unique=None)
NUSCENES = dict(num_pt=40960, img_h=160, img_w=320, channels=64, unique=(26000, 34000))

# KITTI odometry P2 intrinsics (standard published values; not in the reference)
# pushed through the dataset's pipeline: x0.5, centre crop to 512x160, x0.25
# (/root/reference/dataset/KittiDataset.py:290-309).
_FX, _FY, _CX, _CY = 718.856, 718.856, 607.1928, 185.2157
_RAW_W, _RAW_H = 1241, 376


def kitti_intrinsics(img_h=160, img_w=512):
    half_w = int(round(_RAW_W * 0.5))
    half_h = int(round(_RAW_H * 0.5))
    dx = int((half_w - img_w) / 2) if half_w >= img_w else 0
    dy = int((half_h - img_h) / 2) if half_h >= img_h else 0
    k = torch.tensor([[_FX * 0.5, 0.0, _CX * 0.5 - dx],
                      [0.0, _FY * 0.5, _CY * 0.5 - dy],
                      [0.0, 0.0, 1.0]], dtype=torch.float64)
    k = k * 0.25
    k[2, 2] = 1.0
    return k.to(torch.float32)


def _rot_y(a):
    c, s = math.cos(a), math.sin(a)
    return torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]], dtype=torch.float64)


def _unit(x, dim):
    return x / x.norm(dim=dim, keepdim=True).clamp_min(1e-12)


def make_episode(episode, seed=2023, num_pt=40960, img_h=160, img_w=512, channels=64,
                 unique=None, overlap_flip=0.05, with_features=True):
    """One episode as the reference's Dataset + MultiHeadModel would present it.

    Returns a dict of CPU tensors WITHOUT the batch dimension.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed) + int(episode))
    n_unique = num_pt
    if unique is not None:
        lo, hi = unique
        hi = min(hi, num_pt)
        lo = min(lo, hi)
        n_unique = int(torch.randint(lo, hi + 1, (1,), generator=g))

    def U(lo, hi, n):
        return torch.rand(n, generator=g, dtype=torch.float64) * (hi - lo) + lo

    r = U(2.0, 80.0, n_unique)
    th = U(-math.pi, math.pi, n_unique)
    cam = torch.stack([r * torch.sin(th), U(-2.5, 1.7, n_unique), r * torch.cos(th)], 0)
    if n_unique < num_pt:
        # duplication rule of downsample_pc (NuScenesDataset.py:100-111): whole
        # copies of the cloud, then a random subset without replacement
        reps = [torch.arange(n_unique)]
        while n_unique + sum(x.numel() for x in reps) < num_pt:
            reps.append(torch.arange(n_unique))
        fix = torch.cat(reps)
        extra = torch.randperm(n_unique, generator=g)[: num_pt - fix.numel()]
        cam = cam[:, torch.cat([fix, extra])]
    cam32 = cam.to(torch.float32)

    K = kitti_intrinsics(img_h, img_w)
    H, W = img_h // 4, img_w // 4
    # GT in-frustum mask the way the dataset computes it (KittiDataset.py:314-320)
    uvw = K.double() @ cam32.double()
    xy = torch.round(uvw[0:2] / uvw[2:3])
    pc_mask = ((xy[0] >= 0) & (xy[0] <= W - 1) & (xy[1] >= 0) & (xy[1] <= H - 1) & (uvw[2] > 0))

    ang = float(U(-math.pi, math.pi, 1))
    t = torch.stack([U(-10.0, 10.0, 1)[0], torch.tensor(0.0, dtype=torch.float64), U(-10.0, 10.0, 1)[0]])
    P_rand = torch.eye(4, dtype=torch.float64)
    P_rand[:3, :3] = _rot_y(ang)
    P_rand[:3, 3] = t
    pc = (P_rand[:3, :3] @ cam32.double() + P_rand[:3, 3:4]).to(torch.float32)
    P = torch.linalg.inv(P_rand).to(torch.float32)

    flips = torch.rand(num_pt, generator=g) < overlap_flip
    overlap_pred = pc_mask ^ flips

    out = {
        "pc": pc.contiguous(),
        "K": K,
        "P": P,
        "pc_mask": pc_mask.long(),
        "pc_in_cam_space": cam32.contiguous(),
        "pc_overlap_pred": overlap_pred,
        "angles": torch.tensor([0.0, ang, 0.0], dtype=torch.float64),
        "translation": t,
        "img_hw": (img_h, img_w),
    }
    if with_features:
        out["pc_geo_feat"] = _unit(torch.randn(channels, num_pt, generator=g), 0).contiguous()
        out["img_geo_feat"] = _unit(torch.randn(channels, H, W, generator=g), 0).contiguous()
    return out


def make_batch(batch, first_episode=0, seed=2023, **shape):
    """Stack episodes into the reference's batched ``data`` dict (CPU tensors)."""
    eps = [make_episode(first_episode + i, seed=seed, **shape) for i in range(batch)]
    img_h, img_w = eps[0]["img_hw"]
    data = {k: torch.stack([e[k] for e in eps], 0) for k in eps[0] if k != "img_hw"}
    # `img` is read for its shape only (environment.py:33-35); expand() keeps it at 3 floats
    data["img"] = torch.zeros(1, 1, 1, 1).expand(batch, 3, img_h, img_w)
    return data


def make_actions(batch, iterations, seed=2023, first_episode=0, dof6=False, bins=11):
    """Uniform int64 actions, [iterations, B, 1|3] rotation and [iterations, B, 2|3] translation."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed) * 7919 + int(first_episode))
    nr, nt = (3, 3) if dof6 else (1, 2)
    a_r = torch.randint(0, bins, (iterations, batch, nr), generator=g, dtype=torch.int64)
    a_t = torch.randint(0, bins, (iterations, batch, nt), generator=g, dtype=torch.int64)
    return a_r, a_t


def make_cloud_batch(batch, num_pt=40960, seed=2023, first=0, unique=None):
    """Point-major ``xyz [B,N,3]`` for the PointNN front-end microbench (config 4)."""
    clouds = [make_episode(first + i, seed=seed, num_pt=num_pt, unique=unique,
                           with_features=False)["pc"].t().contiguous() for i in range(batch)]
    return torch.stack(clouds, 0)


class StepConfig:
    """The three attributes environment.step/expert read from the reference's config
    (/root/reference/config/KittiConfig.py:101-109): float64 step tables and the DoF flag."""

    def __init__(self, device="cpu", is_6_DoF=False):
        r = torch.tensor([-62.5, -12.5, -2.5, -0.5, -0.1, 0.0, 0.1, 0.5, 2.5, 12.5, 62.5],
                         dtype=torch.float64) * math.pi / 180
        t = torch.tensor([-8.1, -2.7, -0.9, -0.3, -0.1, 0.0, 0.1, 0.3, 0.9, 2.7, 8.1], dtype=torch.float64)
        self.r_steps = r.to(device)
        self.t_steps = t.to(device)
        self.num_steps = 11
        self.is_6_DoF = is_6_DoF
        self.action_num = 10
