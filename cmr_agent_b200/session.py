"""Pipelined registration rollouts from host memory (binding of the native rollout session, include/cmr_b200.h).

One rollout = what ``Test_Agent.py:150-170`` does for a batch with the actions given: ``init``, ``to_disentangled`` of
the target, then ``iters`` times ``observation_from_a_pose`` -> ``step`` -> ``reward``.  The session owns device
buffers for ``depth`` rollouts, a copy stream and a compute stream: ``submit`` enqueues the uploads of a batch (from
pinned host tensors) and its kernels and returns at once, ``wait`` hands back the per-iteration rewards / distances
and the final poses.  With ``depth >= 2`` the upload of batch k+1 runs under the kernels of batch k - the drop-in
functions cannot overlap the two because every batch starts with its own uploads on the caller's stream.

For callers whose policy is scripted, precomputed or (later) on the device; the reference's interactive loop uses the
drop-in module functions.
"""
import ctypes

import torch

from . import _lib
from . import environment as _env


class RolloutSession:
    def __init__(self, B, N, C, H, W, iters, config, depth=2, features_resident=False, reward_mode="shipped", device=None):
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.shape = (B, N, C, H, W)
        self.iters = iters
        self.dof6 = bool(config.is_6_DoF)
        rot, tt = _env.build_step_tables(config.r_steps, config.t_steps)          # host tables, as environment.step uses
        self._tabs = (rot.contiguous(), tt.contiguous())
        cfg = _lib.SessionConfig()
        cfg.B, cfg.N, cfg.C, cfg.H, cfg.W, cfg.iters = B, N, C, H, W, iters
        cfg.dof6 = int(self.dof6)
        cfg.reward_mode = 1 if reward_mode == "intended" else 0
        cfg.depth, cfg.features_resident, cfg.nbins = depth, int(bool(features_resident)), int(tt.shape[0])
        cfg.rot_tab, cfg.t_tab = rot.data_ptr(), tt.data_ptr()
        self.features_resident = bool(features_resident)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cmr_session_create(ctypes.byref(cfg), ctypes.byref(self._h)), "cmr_session_create")
        self._keep = {}

    def close(self):
        if self._h:
            self.lib.cmr_session_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def submit(self, data, action_r, action_t):
        """data: the reference's batch dict with HOST tensors (pinned for overlap); with ``features_resident`` the
        two feature tensors are CUDA tensors.  action_r / action_t: [iters, B, 1|3] / [iters, B, 2|3] int64 on the host."""
        B, N, C, H, W = self.shape
        nr, nt = (3, 3) if self.dof6 else (1, 2)

        def host(t, dtype, shape, name):
            if t.is_cuda or t.dtype != dtype or tuple(t.shape) != shape or not t.is_contiguous():
                raise _lib.CmrError(f"{name}: expected a contiguous host tensor {dtype} {list(shape)}")
            return t

        ov = data["pc_overlap_pred"]
        if ov.dtype == torch.bool:
            ov = ov.view(torch.uint8)
        feat, img = data["pc_geo_feat"], data["img_geo_feat"]
        if self.features_resident:
            feat = _lib.require_cuda(feat, "data['pc_geo_feat']", torch.float32).contiguous()
            img = _lib.require_cuda(img, "data['img_geo_feat']", torch.float32).contiguous()
        else:
            feat = host(feat, torch.float32, (B, C, N), "pc_geo_feat")
            img = host(img, torch.float32, (B, C, H, W), "img_geo_feat")
        keep = (host(data["pc"], torch.float32, (B, 3, N), "pc"), host(ov, torch.uint8, (B, N), "pc_overlap_pred"), feat, img,
                host(data["K"], torch.float32, (B, 3, 3), "K"), host(data["P"], torch.float32, (B, 4, 4), "P"),
                host(data["pc_in_cam_space"], torch.float32, (B, 3, N), "pc_in_cam_space"),
                host(data["pc_mask"], torch.int64, (B, N), "pc_mask"),
                host(action_r, torch.int64, (self.iters, B, nr), "action_r"),
                host(action_t, torch.int64, (self.iters, B, nt), "action_t"))
        inp = _lib.RolloutInputs(*[t.data_ptr() for t in keep])
        ticket = ctypes.c_longlong(-1)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cmr_session_submit(self._h, ctypes.byref(inp), ctypes.byref(ticket)), "cmr_session_submit")
        self._keep[ticket.value] = keep                      # host buffers stay alive until the rollout is collected
        return ticket.value

    def wait(self, ticket):
        """-> (rewards [iters,B], distances [iters,B], poses [B,4,4], disentangled target poses [B,4,4]) on the host."""
        B = self.shape[0]
        rew, dist = torch.empty(self.iters, B), torch.empty(self.iters, B)
        pose, target = torch.empty(B, 4, 4), torch.empty(B, 4, 4)
        _lib.check(self.lib.cmr_session_wait(self._h, ticket, rew.data_ptr(), dist.data_ptr(), pose.data_ptr(),
                                             target.data_ptr()), "cmr_session_wait")
        self._keep.pop(ticket, None)
        return rew, dist, pose, target

    def stats(self):
        gbs, nbytes = ctypes.c_double(0.0), ctypes.c_double(0.0)
        _lib.check(self.lib.cmr_session_stats(self._h, ctypes.byref(gbs), ctypes.byref(nbytes)), "cmr_session_stats")
        return {"h2d_gbs": gbs.value, "h2d_bytes_per_rollout": nbytes.value}

    def last_observation(self, ticket):
        """(obs2d [B,2C,H,W], obs3d [B,5,N]) of the last iteration of rollout `ticket`, as fresh CUDA tensors."""
        B, N, C, H, W = self.shape
        obs2d = torch.empty(B, 2 * C, H, W, device=self.device)
        obs3d = torch.empty(B, 5, N, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cmr_session_last_observation(self._h, ticket, obs2d.data_ptr(), obs3d.data_ptr(),
                                                             _lib.stream_handle(self.device.index)),
                       "cmr_session_last_observation")
        return obs2d, obs3d
