"""torch-CPU restatement ("port") of the reference environment.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows
/root/reference/environment/environment.py function by function; each function
cites the lines it restates.  It keeps the reference's operator sequence
(per-sample python loop, boolean-mask compaction, k=3 bmm, scatter-add + count
+ divide), because bench.py times it as the CPU baseline of the path
(cpu_baseline.kind == "port") and the arithmetic order decides the integer
outputs.

Parity pinning: the reference has no tests or golden vectors.  This port is
pinned against the real reference module, loaded by path in the build
container, by tests/test_oracle_vs_reference.py (bit-exact on every output),
and against the fixtures in tests/golden/ which were produced by the real
reference (tests/golden/make_golden.py).

The only deliberate extension: functions that take the cloud mean accept an
optional precomputed ``mean`` ([B,3,1]); when omitted they compute
``pc.mean(dim=2, keepdim=True)`` exactly like the reference.  torch's fp32 mean
is not correctly rounded and depends on device / ISA, so GPU parity tests pass
the device's mean in to compare everything downstream bit-for-bit.
"""
import functools
import math

import torch

from . import shims


def cloud_mean(pc):
    """[B,3,N] -> [B,3,1]; environment.py:46,91,274 (``pc.mean(dim=2, keepdim=True)``)."""
    return pc.mean(dim=2, keepdim=True)


def disentangled_transform(points, mean, RT):
    """environment.py:54-56 / :92-93: rotate about the cloud mean, then translate.

    points [b,3,n], mean [b,3,1], RT [b,4,4].  ((R @ (p - m)) + m) + t, evaluated left to right.
    """
    centred = points - mean
    rotated = torch.bmm(RT[:, 0:3, 0:3], centred)
    return rotated + mean + RT[:, 0:3, 3:4]


def project_pinhole(K, cam_points, H, W):
    """environment.py:58-65 / :95-101.  Returns (uvz [b,3,n] with u,v divided by z, in_cam bool [b,n])."""
    uvz = torch.bmm(K, cam_points)
    uvz[:, 0:2, :] = uvz[:, 0:2, :] / uvz[:, 2:3, :]
    u, v, z = uvz[:, 0, :], uvz[:, 1, :], uvz[:, 2, :]
    in_cam = (u >= 0) & (u <= (W - 1)) & (v >= 0) & (v <= (H - 1)) & (z > 0)
    return uvz, in_cam


def pixel_index(uvz, in_cam, H, W):
    """environment.py:67-72: round-half-even to int32, row-major pixel id, H*W for out-of-frustum."""
    uv_int = uvz[:, 0:2, :].round().int()
    idx = uv_int[:, 1, :] * W + uv_int[:, 0, :]
    idx[~in_cam] = H * W
    return idx


def scatter_mean_to_grid(point_feat, idx, H, W):
    """environment.py:74-82: pad one zero feature into the dump bin H*W, scatter-mean over
    pixel ids, drop the dump bin.  point_feat [1,C,m], idx [1,m] int32 -> [1,C,H,W]."""
    P = H * W
    C = point_feat.shape[1]
    pad_feat = torch.zeros_like(point_feat[:, :, 0:1])
    src = torch.cat([point_feat, pad_feat], dim=-1)
    pad_idx = torch.ones_like(idx[:, 0:1]).long() * P
    index = torch.cat([idx, pad_idx], dim=-1)
    # the reference hard-codes 64 channels in the index repeat (environment.py:79)
    binned = shims.scatter_mean(src, index.unsqueeze(1).repeat(1, C, 1), dim=2)
    return binned[:, :, :P].view(1, C, H, W)


@torch.no_grad()
def observation_from_a_pose(data, RT, mean=None):
    """environment.py:25-126 -> (obs2d [B,2C,H,W], obs3d [B,5,N])."""
    K = data["K"]
    pc = data["pc"]
    overlap_pred = data["pc_overlap_pred"]
    pc_feat = data["pc_geo_feat"]
    img_feat = data["img_geo_feat"]
    B = pc.shape[0]
    H = data["img"].shape[2] // 4
    W = data["img"].shape[3] // 4
    if mean is None:
        mean = cloud_mean(pc)

    planes = []
    for i in range(B):
        sel = overlap_pred[i]
        pts = pc[i:i + 1][:, :, sel]                       # :48 boolean-mask compaction
        feats = pc_feat[i:i + 1, :, sel]                   # :49
        cam = disentangled_transform(pts, mean[i:i + 1], RT[i:i + 1])
        uvz, in_cam = project_pinhole(K[i:i + 1], cam, H, W)
        idx = pixel_index(uvz, in_cam, H, W)
        grid = scatter_mean_to_grid(feats, idx, H, W)
        planes.append(torch.cat([img_feat[i:i + 1], grid], dim=1))   # :83
    obs2d = torch.cat(planes, dim=0)

    cam_all = disentangled_transform(pc, mean, RT)                       # :91-93
    _, in_cam_all = project_pinhole(K, cam_all, H, W)                    # :95-101
    obs3d = torch.cat([pc, overlap_pred.unsqueeze(1).float(), in_cam_all.unsqueeze(1).float()], dim=1)  # :121-124
    return obs2d, obs3d


@torch.no_grad()
def projected_pixels(data, RT, mean=None):
    """Integer by-products of environment.py:54-72 for ALL points (not only the predicted-overlap
    subset): (idx [B,N] int32 with H*W for out-of-frustum, in_cam [B,N] bool).  Column j of the
    reference's per-subset computation equals column j here."""
    pc = data["pc"]
    H = data["img"].shape[2] // 4
    W = data["img"].shape[3] // 4
    if mean is None:
        mean = cloud_mean(pc)
    cam = disentangled_transform(pc, mean, RT)
    uvz, in_cam = project_pinhole(data["K"], cam, H, W)
    return pixel_index(uvz, in_cam, H, W), in_cam


def init(data):
    """environment.py:129-140."""
    B = data["pc"].shape[0]
    pose_target = data["P"].clone()
    pose_source = torch.eye(4).repeat(B, 1, 1)
    return pose_source, pose_target


@torch.no_grad()
def to_disentangled(poses, pcd, mean=None):
    """environment.py:15-21 (in place): t <- (t - m) + R m."""
    m = pcd[:, 0:3, :].mean(dim=2) if mean is None else mean.reshape(-1, 3)
    poses[:, :3, 3] = poses[:, :3, 3] - m + (poses[:, :3, :3] @ m.unsqueeze(-1)).squeeze(-1)
    return poses


def axis_rotation(axis, angle):
    """environment.py:235-260."""
    c, s = torch.cos(angle), torch.sin(angle)
    one, zero = torch.ones_like(angle), torch.zeros_like(angle)
    flat = {"X": (one, zero, zero, zero, c, -s, zero, s, c),
            "Y": (c, zero, s, zero, one, zero, -s, zero, c),
            "Z": (c, -s, zero, s, c, zero, zero, zero, one)}[axis]
    return torch.stack(flat, -1).reshape(angle.shape + (3, 3))


def euler_angles_to_matrix(euler_angles, convention):
    """environment.py:210-232: left fold of matmul over the three axis rotations."""
    if euler_angles.dim() == 0 or euler_angles.shape[-1] != 3:
        raise ValueError("Invalid input euler angles.")
    if len(convention) != 3:
        raise ValueError("Convention must have 3 letters.")
    if convention[1] in (convention[0], convention[2]):
        raise ValueError(f"Invalid convention {convention}.")
    for letter in convention:
        if letter not in ("X", "Y", "Z"):
            raise ValueError(f"Invalid letter {letter} in convention string.")
    mats = [axis_rotation(a, e) for a, e in zip(convention, torch.unbind(euler_angles, -1))]
    return functools.reduce(torch.matmul, mats)


def step(action_r, action_t, pose_source, config):
    """environment.py:179-207 (in place, returns the same tensor)."""
    r_steps, t_steps = config.r_steps, config.t_steps
    B = action_r.shape[0]
    move_r = torch.zeros((B, 3))
    move_t = torch.zeros((B, 3))
    if config.is_6_DoF:
        for axis in range(3):
            move_r[:, axis] = r_steps[action_r[:, axis]]
            move_t[:, axis] = t_steps[action_t[:, axis]]
    else:
        move_r[:, 1] = r_steps[action_r[:, 0]]
        move_t[:, 0] = t_steps[action_t[:, 0]]
        move_t[:, 2] = t_steps[action_t[:, 1]]
    pose_source[:, :3, :3] = euler_angles_to_matrix(move_r, "XYZ") @ pose_source[:, :3, :3]
    pose_source[:, :3, 3] += move_t
    return pose_source


def reward(RT, data, prev_distance=None, mode="shipped", mean=None):
    """environment.py:263-302.

    mode "shipped": the code as shipped - RT is ignored, the distance is between
    pc_in_cam_space and the mean-centred, un-transformed cloud (:272-290).
    mode "intended": applies the disentangled transform of the commented line :273 (our
    extension)."""
    target = data["pc_in_cam_space"]
    mask = data["pc_mask"].bool()
    pc = data["pc"]
    B = pc.shape[0]
    if mean is None:
        mean = cloud_mean(pc)
    if mode == "shipped":
        moved = pc - mean
    elif mode == "intended":
        moved = disentangled_transform(pc, mean, RT)
    else:
        raise ValueError(mode)
    dist_out = torch.zeros(B)
    for i in range(B):
        a = target[i][:, mask[i]]
        b = moved[i][:, mask[i]]
        d = (a - b) * (a - b)
        dist_out[i] = d.sum(dim=0).mean()
    dist_out = dist_out.unsqueeze(-1).unsqueeze(-1)
    if prev_distance is None:
        return torch.zeros_like(dist_out), dist_out
    better = (dist_out < prev_distance).float() * 0.5
    same = (dist_out == prev_distance).float() * 0
    worse = (dist_out > prev_distance).float() * 0.5
    return better - worse - same, dist_out


def expert(pose_source, targets, config, data=None):
    """environment.py:143-176.  scipy Rotation.as_euler('xyz') (extrinsic xyz), float64."""
    from scipy.spatial.transform import Rotation

    delta_t = targets[:, :3, 3] - pose_source[:, :3, 3]
    delta_R = targets[:, :3, :3] @ pose_source[:, :3, :3].transpose(2, 1)
    delta_r = Rotation.from_matrix(delta_R.cpu().numpy()).as_euler("xyz")
    big = delta_r[:, 0] > 3
    delta_r[big, 0] = 0
    delta_r[big, 2] = 0
    pos = delta_r[:, 1] > 0
    delta_r[big & pos, 1] = math.pi - delta_r[big & pos, 1]
    neg = delta_r[:, 1] < 0
    delta_r[big & neg, 1] = -1 * math.pi - delta_r[big & neg, 1]
    delta_r = torch.from_numpy(delta_r)
    err_r = torch.abs(delta_r.unsqueeze(-1) - config.r_steps.unsqueeze(0).unsqueeze(0))
    action_r = err_r.argmin(dim=2)
    err_t = torch.abs(delta_t.unsqueeze(-1) - config.t_steps.unsqueeze(0).unsqueeze(0))
    action_t = err_t.argmin(dim=2)
    if not config.is_6_DoF:
        action_r = action_r[:, 1:2]
        action_t = torch.cat([action_t[:, 0:1], action_t[:, 2:3]], dim=1)
    return action_r, action_t
