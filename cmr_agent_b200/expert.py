"""expert() of the reference environment on the device (environment/environment.py:143-176).

The reference leaves the GPU every step here: delta_R -> .cpu().numpy() -> scipy Rotation.as_euler
-> torch.from_numpy().to(DEVICE).  cmr_expert does the same arithmetic in one tiny kernel (fp32
delta_R, then fp64 scipy's from_matrix/as_euler algorithm, the ">3 rad" fix-ups, first-minimum
argmin over the float64 step tables) and returns int64 actions on the device.
"""
import torch

from . import _lib


def expert(pose_source, targets, config, data=None):
    src = _lib.require_cuda(pose_source, "pose_source", torch.float32)
    tgt = _lib.require_cuda(targets, "targets", torch.float32)
    src = src if src.is_contiguous() else src.contiguous()
    tgt = tgt if tgt.is_contiguous() else tgt.contiguous()
    dev = src.device
    r_steps = config.r_steps.to(device=dev, dtype=torch.float64).contiguous()
    t_steps = config.t_steps.to(device=dev, dtype=torch.float64).contiguous()
    if r_steps.numel() != t_steps.numel():
        raise _lib.CmrError("r_steps and t_steps must have the same number of bins")
    B = src.shape[0]
    dof6 = bool(config.is_6_DoF)
    action_r = torch.empty(B, 3 if dof6 else 1, dtype=torch.int64, device=dev)
    action_t = torch.empty(B, 3 if dof6 else 2, dtype=torch.int64, device=dev)
    _lib.call("cmr_expert", _lib.ptr(src), _lib.ptr(tgt), _lib.ptr(r_steps), _lib.ptr(t_steps), int(r_steps.numel()),
              int(dof6), B, _lib.ptr(action_r), _lib.ptr(action_t), _lib.stream())
    return action_r, action_t
