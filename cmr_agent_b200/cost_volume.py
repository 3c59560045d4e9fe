"""The pose-sampling cost volume of models/IterModel.py:272-351 on the device.

``warp`` replaces the block between ``delta_RT = self.sample_poses(...)`` (:275) and the slicing at :350-351: every
cloud is transformed by its K candidate poses (X = R p + t, :281), projected (:283-286), tested against the frustum
(:293-297) and rounded (:304); the features of the masked points are scatter-MEANed and their in-camera scores
scatter-SUMmed onto the H x W grid (:341-343), 200 poses at a time in the reference, all at once here.

The same two kernels as the observation do the work (k_project with `share` poses per cloud, k_tile_gather with the
scores as a 65th, summed channel); sums run in point order, so the result equals the CPU evaluation of the reference's
expressions bit for bit.  CUDA tensors only - no CPU fallback.
"""
import torch

from . import _lib


def warp(pc, pc_mask, delta_RT, K, pc_geo_feat, pc_is_in_cam_scores, H, W):
    """pc [B,3,N] f32, pc_mask [N] bool (IterModel.py:272 uses the first sample's mask for every cloud),
    delta_RT [B,Kp,3,4] f32 (``sample_poses(...).view(B, -1, 3, 4)``, :276), K [B,3,3], pc_geo_feat [B,C,N],
    pc_is_in_cam_scores [B,N]  ->  (pc_warped_geo_feat [B,Kp,C,H*W], pc_warped_occupancy [B,Kp,H*W])."""
    pc = _lib.require_cuda(pc, "pc", torch.float32).contiguous()
    feat = _lib.require_cuda(pc_geo_feat, "pc_geo_feat", torch.float32)
    scores = _lib.require_cuda(pc_is_in_cam_scores, "pc_is_in_cam_scores", torch.float32)
    poses = _lib.require_cuda(delta_RT, "delta_RT", torch.float32)
    mask = _lib.require_cuda(pc_mask, "pc_mask")
    dev = pc.device
    B, three, N = pc.shape
    C = feat.shape[1]
    if three != 3 or tuple(feat.shape) != (B, C, N) or tuple(scores.shape) != (B, N) or tuple(mask.shape) != (N,):
        raise _lib.CmrError("cost_volume.warp: pc [B,3,N], pc_geo_feat [B,C,N], scores [B,N], pc_mask [N] expected")
    if poses.dim() != 4 or poses.shape[0] != B or tuple(poses.shape[2:]) != (3, 4):
        raise _lib.CmrError("cost_volume.warp: delta_RT must be [B,Kp,3,4]")
    if C % 64 != 0:
        raise _lib.CmrError("cost_volume.warp: the channel count must be a multiple of 64 (the reference uses 64)")
    Kp, P = poses.shape[1], H * W
    Kmat = K.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
    # features + the scores as one more (summed) channel; rows stay 16-byte multiples
    Cx = C + 4
    feat_x = torch.zeros(B, Cx, N, device=dev, dtype=torch.float32)
    feat_x[:, :C] = feat
    feat_x[:, C] = scores
    mask_u8 = (mask != 0).view(1, N).expand(B, N).contiguous().view(torch.uint8)
    poses44 = torch.zeros(B * Kp, 4, 4, device=dev, dtype=torch.float32)
    poses44[:, :3, :] = poses.reshape(B * Kp, 3, 4)
    poses44[:, 3, 3] = 1.0
    lib = _lib.load()
    nbytes = lib.cmr_cost_volume_workspace_bytes(B, Kp, N, Cx, P)
    if nbytes == 0:
        raise _lib.CmrError("cost_volume.warp: B * Kp must be at most 65535")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = torch.empty(B * Kp, Cx, P, device=dev, dtype=torch.float32)
    st = _lib.stream()
    _lib.call("cmr_cost_volume_prepare", _lib.ptr(mask_u8), _lib.ptr(feat_x), B, Kp, N, Cx, _lib.ptr(ws), st)
    _lib.call("cmr_cost_volume_warp", _lib.ptr(pc), _lib.ptr(mask_u8), _lib.ptr(Kmat), _lib.ptr(poses44), _lib.ptr(ws),
              B, Kp, N, Cx, H, W, C, _lib.ptr(out), st)
    out = out.view(B, Kp, Cx, P)
    return out[:, :, :C], out[:, :, C]
