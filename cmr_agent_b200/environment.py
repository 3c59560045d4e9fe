"""Drop-in for the reference's ``environment/environment.py`` on B200.

Same module-level functions, argument meaning, in-place/aliasing behaviour and
return shapes as /root/reference/environment/environment.py
(``to_disentangled:15``, ``observation_from_a_pose:25``, ``init:129``,
``expert:143``, ``step:179``, ``euler_angles_to_matrix:210``, ``reward:263``),
so ``Train_Agent.py`` / ``Test_Agent.py`` / ``CMRAgent.py`` run unchanged after
``cmr_agent_b200.install()``.  All arithmetic on the path runs in the
hand-written sm_100a kernels of libcmr_b200.so (include/cmr_b200.h) on the
current CUDA stream, with no host synchronisation; PyTorch allocates memory and
hands over the stream.  The cloud mean (``pc.mean(dim=2)`` in the reference) is
computed once per episode batch by cmr_cloud_mean in fp64 - see DESIGN.md
"cloud mean" for why torch's own fp32 mean is not used by default.

Differences that are deliberate and documented:
  * CUDA only.  CPU tensors where the reference expects device tensors raise.
  * The per-episode work the reference redoes every iteration (boolean-mask
    compaction of the predicted-overlap points, cloud mean, H2D of K /
    pc_in_cam_space / pc_mask, the pose-independent distance of the shipped
    reward) is done once per batch and cached in a side table keyed by the
    identity of ``data`` and of its tensors (which the cache keeps alive);
    nothing is written into the caller's dict.
  * ``reward`` reproduces the shipped behaviour (the pose argument is ignored,
    environment.py:272-275) unless ``set_reward_mode("intended")`` is chosen.
"""
import collections
import functools
import math

import torch

from . import _lib

DEVICE = torch.device("cuda")

_reward_mode = "shipped"
_mean_provider = "kernel"
_validate = False


def set_reward_mode(mode):
    """"shipped" (default; environment.py:272-290 verbatim semantics) or "intended"
    (applies the disentangled transform of the commented line :273)."""
    global _reward_mode
    if mode not in ("shipped", "intended"):
        raise ValueError(mode)
    _reward_mode = mode


def set_mean_provider(name):
    """"kernel" (default): cmr_cloud_mean - fp64 accumulation in a fixed order: correctly rounded,
    deterministic, and independent of the batch shape (so sharding episodes over GPUs cannot change it).
    "torch": ``pc.mean(dim=2)`` on the device - the reference's own expression, but torch's fp32 mean is
    neither correctly rounded nor shape invariant (the same cloud gives different bits at B=2 and B=16)."""
    global _mean_provider
    if name not in ("torch", "kernel"):
        raise ValueError(name)
    _mean_provider = name
    _episodes.clear()


def set_validation(on):
    """Validation mode: after every call the library's sticky device fault word is read back (this SYNCHRONISES the
    stream) and a fault - an action or gather index out of range, an activation of the 3-D tower outside the fp16
    range - raises CmrError, where torch would have raised a device assert.  Off by default: the hot path never syncs."""
    global _validate
    _validate = bool(on)


def check_fault(what="call"):
    code = _lib.take_fault()
    if code:
        raise _lib.CmrError(f"{what}: device fault {code} "
                            "(1 = gather index out of range, 2 = action out of range, 3 = tower activation beyond fp16 range)")


def _dev_f32(t, name):
    t = _lib.require_cuda(t, name, torch.float32)
    return t if t.is_contiguous() else t.contiguous()


def cloud_mean(pc):
    """[B,3,N] -> [B,3] with the selected provider."""
    if _mean_provider == "torch":
        return pc.mean(dim=2).contiguous()
    pc = pc if pc.is_contiguous() else pc.contiguous()
    B, _, N = pc.shape
    out = torch.empty(B, 3, device=pc.device, dtype=torch.float32)
    _lib.call("cmr_cloud_mean", _lib.ptr(pc), B, N, _lib.ptr(out), _lib.stream())
    return out


# ---- per-batch derived state ------------------------------------------------------------------------------------
# The reference recomputes everything from `data` on every call; the drop-in hoists the per-episode work (overlap
# compaction, cloud mean, H2D of K / pc_in_cam_space / pc_mask, the pose-independent reward distance) out of the
# iteration loop.  That state lives in a small side table keyed by id(data) - nothing is written into the caller's
# dict - and every entry HOLDS the tensors it was derived from: a hit requires the very same tensor objects at the same
# in-place version, and because the entry keeps them alive their storage cannot be recycled under another tensor.
_EP_KEYS = ("pc", "pc_overlap_pred", "pc_geo_feat", "img_geo_feat", "K")
_RW_KEYS = ("pc", "pc_in_cam_space", "pc_mask")
_CACHE_ENTRIES = 2          # the training batch and a validation batch (Train_Agent.py:160-213)


class _Table(collections.OrderedDict):
    def lookup(self, data, keys, extra):
        ent = self.get(id(data))
        if ent is not None:
            src = ent.src
            for k, (t, v) in zip(keys, src):
                d = data.get(k)
                if d is not t or d._version != v:
                    return None
            if ent.extra != extra:
                return None
        return ent

    def store(self, data, ent):
        self[id(data)] = ent
        self.move_to_end(id(data))
        while len(self) > _CACHE_ENTRIES:
            self.popitem(last=False)


_episodes = _Table()
_rewards = _Table()


def clear_cache():
    """Drop every cached per-batch state (and the tensors it keeps alive)."""
    _episodes.clear()
    _rewards.clear()


class _Episode:
    """Per-batch derived state of observation_from_a_pose (never visible to callers)."""

    def __init__(self, data):
        pc = _dev_f32(data["pc"], "data['pc']")
        feat = _dev_f32(data["pc_geo_feat"], "data['pc_geo_feat']")
        img_feat = _dev_f32(data["img_geo_feat"], "data['img_geo_feat']")
        overlap = _lib.require_cuda(data["pc_overlap_pred"], "data['pc_overlap_pred']")
        if overlap.dtype != torch.bool:
            overlap = overlap != 0
        overlap = overlap.contiguous()
        B, three, N = pc.shape
        if three != 3:
            raise _lib.CmrError("data['pc'] must be [B,3,N]")
        C = feat.shape[1]
        img = data["img"]
        H, W = img.shape[2] // 4, img.shape[3] // 4                      # environment.py:33-35
        if tuple(img_feat.shape) != (B, C, H, W):
            raise _lib.CmrError(f"img_geo_feat {tuple(img_feat.shape)} does not match (B,C,H/4,W/4)={(B, C, H, W)}")
        if tuple(feat.shape) != (B, C, N) or tuple(overlap.shape) != (B, N):
            raise _lib.CmrError("pc_geo_feat / pc_overlap_pred do not match pc")
        self.src = tuple((data[k], data[k]._version) for k in _EP_KEYS)
        self.extra = (tuple(img.shape), _mean_provider, id(data.get("_cmr_b200_mean_override")))
        self.B, self.N, self.C, self.H, self.W = B, N, C, H, W
        self.pc, self.img_feat, self.feat = pc, img_feat, feat
        self.overlap = overlap.view(torch.uint8)
        self.K = data["K"].to(device=pc.device, dtype=torch.float32, non_blocking=True).contiguous()  # :26, once
        self.mean = data.get("_cmr_b200_mean_override")
        if self.mean is None:
            self.mean = cloud_mean(pc)                                   # :46,91 once per episode
        else:
            self.mean = self.mean.to(pc.device, torch.float32).reshape(B, 3).contiguous()
        lib = _lib.load()
        nbytes = lib.cmr_workspace_bytes(B, N, C, H * W)
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=pc.device)
        self.dev_index = pc.device.index if pc.device.index is not None else torch.cuda.current_device()
        _lib.call("cmr_episode_prepare", _lib.ptr(self.overlap), _lib.ptr(feat), B, N, C, _lib.ptr(self.ws),
                  _lib.stream())
        # everything cmr_observe needs that does not change between iterations, ready for ctypes
        self.img_rows = None     # sample_image_features: the image features pixel-major, built on first use
        self.observe_head = (self.pc.data_ptr(), self.overlap.data_ptr(), self.img_feat.data_ptr(), self.K.data_ptr())
        self.observe_mid = (self.mean.data_ptr(), self.ws.data_ptr(), B, N, C, H, W)
        self.obs2d_shape = (B, 2 * C, H, W)
        self.obs3d_shape = (B, 5, N)
        self.pose_shape = (B, 4, 4)
        self.device = pc.device
        self.iter_args = None


def _episode(data):
    ep = _episodes.lookup(data, _EP_KEYS, (tuple(data["img"].shape), _mean_provider, id(data.get("_cmr_b200_mean_override"))))
    if ep is None:
        ep = _Episode(data)
        _episodes.store(data, ep)
    return ep


def episode_state(data):
    """The cached per-batch state of `data` (tests and benchmarks read the cloud mean / workspace from it)."""
    return _episode(data)


_observe_fn = None


@torch.no_grad()
def observation_from_a_pose(data, RT, return_pixels=False):
    """environment.py:25-126 -> (observation_2d [B,2C,H,W], observation_3d [B,5,N]), fresh tensors.

    ``return_pixels`` (extension) also returns the int32 pixel id of every point ([B,N], H*W when
    outside the frustum) and the per-episode count of visible predicted-overlap points."""
    global _observe_fn
    ep = _episode(data)
    if not (RT.is_cuda and RT.dtype is torch.float32 and RT.is_contiguous()):
        RT = _dev_f32(RT, "RT")
    if tuple(RT.shape) != ep.pose_shape:
        raise _lib.CmrError(f"RT must be {list(ep.pose_shape)}")
    obs2d = torch.empty(ep.obs2d_shape, device=ep.device, dtype=torch.float32)
    obs3d = torch.empty(ep.obs3d_shape, device=ep.device, dtype=torch.float32)
    pix = mvis = None
    if return_pixels:
        pix = torch.empty(ep.B, ep.N, device=ep.device, dtype=torch.int32)
        mvis = torch.empty(ep.B, device=ep.device, dtype=torch.int32)
    if _observe_fn is None:
        _observe_fn = _lib.bind("cmr_observe")
    rc = _observe_fn(*ep.observe_head, RT.data_ptr(), *ep.observe_mid, obs2d.data_ptr(), obs3d.data_ptr(),
                     pix.data_ptr() if return_pixels else None, mvis.data_ptr() if return_pixels else None,
                     _lib.stream_handle(ep.dev_index))
    if rc:
        _lib.fail(rc, "cmr_observe")
    if return_pixels:
        return obs2d, obs3d, pix, mvis
    return obs2d, obs3d


@torch.no_grad()
def sample_image_features(data, RT):
    """Image features sampled bilinearly at every point's projection (BASELINE.json north_star: "bilinearly sample
    image features onto visible points").  An EXTENSION - the reference has no point-side gather; the
    operator is ``F.grid_sample(img_geo_feat, uv, mode="bilinear", padding_mode="zeros", align_corners=True)`` at
    the pixel coordinates of environment.py:54-59, zero for points outside the frustum of :61-65
    (oracle/sample_oracle.py).

    -> (features [B,C,N] f32 - the layout of data["pc_geo_feat"] -, in_cam [B,N] bool)."""
    ep = _episode(data)
    if not (RT.is_cuda and RT.dtype is torch.float32 and RT.is_contiguous()):
        RT = _dev_f32(RT, "RT")
    if tuple(RT.shape) != ep.pose_shape:
        raise _lib.CmrError(f"RT must be {list(ep.pose_shape)}")
    if ep.C % 2:
        raise _lib.CmrError("sample_image_features: an even number of channels is required")
    st = _lib.stream()
    if ep.img_rows is None:      # [B, H*W, C]: a pixel's channels as one row; does not depend on the pose
        nbytes = _lib.load().cmr_sample_workspace_bytes(ep.B, ep.C, ep.H * ep.W)
        ep.img_rows = torch.empty(nbytes, dtype=torch.uint8, device=ep.device)
        _lib.call("cmr_sample_prepare", _lib.ptr(ep.img_feat), ep.B, ep.C, ep.H * ep.W, _lib.ptr(ep.img_rows), st)
    feats = torch.empty(ep.B, ep.C, ep.N, device=ep.device, dtype=torch.float32)
    in_cam = torch.empty(ep.B, ep.N, device=ep.device, dtype=torch.uint8)
    _lib.call("cmr_sample_image_features", _lib.ptr(ep.pc), _lib.ptr(ep.K), _lib.ptr(RT), _lib.ptr(ep.mean),
              _lib.ptr(ep.img_rows), ep.B, ep.N, ep.C, ep.H, ep.W, _lib.ptr(feats), _lib.ptr(in_cam), st)
    return feats, in_cam.view(torch.bool)


def init(data):
    """environment.py:129-140: identity source poses, ground-truth target pose on the device."""
    B = data["pc"].shape[0]
    pose_target = data["P"].to(DEVICE, non_blocking=True)    # pinned sources copy asynchronously (graph-capturable)
    pose_source = torch.eye(4, device=DEVICE).repeat(B, 1, 1)
    return pose_source, pose_target


reset = init  # the reference has no `reset`; alias kept for callers written against north_star's wording


@torch.no_grad()
def to_disentangled(poses, pcd):
    """environment.py:15-21, in place on ``poses`` (returned): t <- (t - mean) + R mean."""
    _lib.require_cuda(poses, "poses", torch.float32)
    pcd = _lib.require_cuda(pcd, "pcd", torch.float32)
    mean = cloud_mean(pcd[:, 0:3, :] if pcd.shape[1] != 3 else pcd)
    target = poses if poses.is_contiguous() else poses.contiguous()
    _lib.call("cmr_to_disentangled", _lib.ptr(target), _lib.ptr(mean), poses.shape[0], _lib.stream())
    if target is not poses:
        poses.copy_(target)
    return poses


def _axis_angle_rotation(axis, angle):
    """environment.py:235-260."""
    cos = torch.cos(angle)
    sin = torch.sin(angle)
    one = torch.ones_like(angle)
    zero = torch.zeros_like(angle)
    if axis == "X":
        flat = (one, zero, zero, zero, cos, -sin, zero, sin, cos)
    elif axis == "Y":
        flat = (cos, zero, sin, zero, one, zero, -sin, zero, cos)
    elif axis == "Z":
        flat = (cos, -sin, zero, sin, cos, zero, zero, zero, one)
    else:
        raise ValueError(f"Invalid letter {axis} in convention string.")
    return torch.stack(flat, -1).reshape(angle.shape + (3, 3))


def euler_angles_to_matrix(euler_angles, convention):
    """environment.py:210-232 (host-side helper; the step kernel uses tables built from
    ``_axis_angle_rotation``, not this function)."""
    if euler_angles.dim() == 0 or euler_angles.shape[-1] != 3:
        raise ValueError("Invalid input euler angles.")
    if len(convention) != 3:
        raise ValueError("Convention must have 3 letters.")
    if convention[1] in (convention[0], convention[2]):
        raise ValueError(f"Invalid convention {convention}.")
    for letter in convention:
        if letter not in ("X", "Y", "Z"):
            raise ValueError(f"Invalid letter {letter} in convention string.")
    mats = map(_axis_angle_rotation, convention, torch.unbind(euler_angles, -1))
    return functools.reduce(torch.matmul, mats)


def build_step_tables(r_steps, t_steps):
    """Host logic of ``step`` (environment.py:183-201): the float64 step tables become
    rot_tab [3, nbins+1, 3, 3] f32 (per-axis rotation matrix of every bin; the extra last entry is
    angle 0.0, what the 3-DoF branch assigns to the x and z axes) and t_tab [nbins] f32.
    Angles are rounded to fp32 first, exactly as ``move_r[:, i] = r_steps[r]`` does, and cos/sin are
    evaluated by torch on the CPU so the tables are identical on every host."""
    r32 = torch.cat([r_steps.detach().to("cpu", torch.float64).to(torch.float32), torch.zeros(1)])
    t32 = t_steps.detach().to("cpu", torch.float64).to(torch.float32)
    rot = torch.stack([_axis_angle_rotation(a, r32) for a in "XYZ"], 0).contiguous()
    return rot, t32.contiguous()


_table_cache = {}


def _step_tables(config, device):
    r, t = config.r_steps, config.t_steps
    key = (id(r), r._version, id(t), t._version, str(device))
    hit = _table_cache.get(key)
    if hit is None or hit[3] is not r or hit[4] is not t:
        rot, tt = build_step_tables(r, t)
        hit = (rot.to(device), tt.to(device), int(t.shape[0]), r, t)      # r, t kept alive: their ids stay theirs
        if len(_table_cache) > 16:
            _table_cache.clear()
        _table_cache[key] = hit
    return hit


def _actions(a, cols, name):
    if a.is_cuda and a.dtype is torch.int64 and a.dim() == 2 and a.shape[1] == cols and a.is_contiguous():
        return a
    a = _lib.require_cuda(a, name)
    if a.dtype != torch.int64:
        a = a.long()
    if a.dim() != 2 or a.shape[1] < cols:
        raise _lib.CmrError(f"{name} must be [B,{cols}]")
    if a.shape[1] != cols:
        a = a[:, :cols]
    return a.contiguous()


_step_fn = None


def step(action_r, action_t, pose_source, config):
    """environment.py:179-207: in-place pose update, returns ``pose_source``."""
    global _step_fn
    _lib.require_cuda(pose_source, "pose_source", torch.float32)
    dof6 = bool(config.is_6_DoF)
    rot, tt, nbins, _, _ = _step_tables(config, pose_source.device)
    ar = _actions(action_r, 3 if dof6 else 1, "action_r")
    at = _actions(action_t, 3 if dof6 else 2, "action_t")
    B = pose_source.shape[0]
    if ar.shape[0] != B or at.shape[0] != B:
        raise _lib.CmrError("actions and pose_source disagree on the batch size")
    target = pose_source if pose_source.is_contiguous() else pose_source.contiguous()
    if _step_fn is None:
        _step_fn = _lib.bind("cmr_step")
    dev_index = pose_source.device.index if pose_source.device.index is not None else torch.cuda.current_device()
    rc = _step_fn(target.data_ptr(), ar.data_ptr(), at.data_ptr(), rot.data_ptr(), tt.data_ptr(), nbins, int(dof6), B,
                  _lib.stream_handle(dev_index))
    if rc:
        _lib.fail(rc, "cmr_step")
    if target is not pose_source:
        pose_source.copy_(target)
    if _validate:
        check_fault("step")
    return pose_source


class _RewardState:
    def __init__(self, data):
        pc = _dev_f32(data["pc"], "data['pc']")
        dev = pc.device
        self.src = tuple((data[k], data[k]._version) for k in _RW_KEYS)
        self.extra = (_mean_provider, id(data.get("_cmr_b200_mean_override")))
        self.pc = pc
        self.target = data["pc_in_cam_space"].to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()  # :267
        self.mask = (data["pc_mask"].to(dev, non_blocking=True) != 0).contiguous().view(torch.uint8)          # :268
        ep = _episodes.get(id(data))
        if ep is not None and ep.src[0][0] is data["pc"] and ep.src[0][1] == data["pc"]._version:
            self.mean = ep.mean                                          # the observation's mean of the same cloud
        elif data.get("_cmr_b200_mean_override") is not None:
            self.mean = data["_cmr_b200_mean_override"]
        else:
            self.mean = cloud_mean(pc)
        self.mean = self.mean.to(dev, torch.float32).reshape(pc.shape[0], 3).contiguous()
        nbytes = _lib.load().cmr_reward_scratch_bytes(pc.shape[0])
        self.scratch = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        self.dist_cached = None          # the shipped reward's distance: a constant of the batch (environment.py:272-275)
        self.dev_index = dev.index if dev.index is not None else torch.cuda.current_device()


def _reward_state(data):
    st = _rewards.lookup(data, _RW_KEYS, (_mean_provider, id(data.get("_cmr_b200_mean_override"))))
    if st is None:
        st = _RewardState(data)
        _rewards.store(data, st)
    return st


_reward_fn = _reward_cmp_fn = None


def reward(RT, data, prev_distance=None):
    """environment.py:263-302 -> (reward [B,1,1], p2p_distance [B,1,1]).

    In the shipped mode the distance does not depend on the pose (:272-275), so it is computed once per batch and
    memoised; later calls compare the memoised value with ``prev_distance`` (:293-298) and return fresh tensors."""
    global _reward_fn, _reward_cmp_fn
    st = _reward_state(data)
    B, _, N = st.pc.shape
    dev = st.pc.device
    prev = None
    if prev_distance is not None:
        prev = prev_distance
        if not (prev.is_cuda and prev.dtype is torch.float32 and prev.is_contiguous() and prev.numel() == B):
            prev = _lib.require_cuda(prev_distance, "prev_distance", torch.float32).reshape(B).contiguous()
    rew = torch.empty(B, 1, 1, device=dev, dtype=torch.float32)
    dist = torch.empty(B, 1, 1, device=dev, dtype=torch.float32)
    sh = _lib.stream_handle(st.dev_index)
    if _reward_mode == "shipped" and st.dist_cached is not None:
        if _reward_cmp_fn is None:
            _reward_cmp_fn = _lib.bind("cmr_reward_compare")
        rc = _reward_cmp_fn(st.dist_cached.data_ptr(), prev.data_ptr() if prev is not None else None, B, rew.data_ptr(),
                            dist.data_ptr(), sh)
        if rc:
            _lib.fail(rc, "cmr_reward_compare")
        return rew, dist
    mode = 1 if _reward_mode == "intended" else 0
    pose = _dev_f32(RT, "RT") if mode == 1 else None
    if _reward_fn is None:
        _reward_fn = _lib.bind("cmr_reward")
    rc = _reward_fn(st.target.data_ptr(), st.pc.data_ptr(), st.mask.data_ptr(), st.mean.data_ptr(),
                    pose.data_ptr() if pose is not None else None, prev.data_ptr() if prev is not None else None, mode, B, N,
                    st.scratch.data_ptr(), rew.data_ptr(), dist.data_ptr(), sh)
    if rc:
        _lib.fail(rc, "cmr_reward")
    if mode == 0:
        st.dist_cached = dist.clone()    # our own copy: callers may do anything with the tensor they were handed
    return rew, dist


_iteration_fn = None


def _iteration_args(ep, st, config):
    dof6 = bool(config.is_6_DoF)
    rot, tt, nbins, _, _ = _step_tables(config, ep.device)
    mode = 1 if _reward_mode == "intended" else 0
    cached = st.dist_cached if (st is not None and mode == 0) else None
    key = (id(rot), id(st), mode, id(cached), dof6)
    if ep.iter_args is None or ep.iter_args[0] != key:
        a = _lib.IterationArgs()
        a.pc, a.overlap, a.img_feat, a.K = ep.observe_head
        a.mean, a.workspace = ep.mean.data_ptr(), ep.ws.data_ptr()
        a.rot_tab, a.t_tab = rot.data_ptr(), tt.data_ptr()
        if st is not None:
            a.target, a.mask, a.reward_scratch = st.target.data_ptr(), st.mask.data_ptr(), st.scratch.data_ptr()
            a.dist_cached = cached.data_ptr() if cached is not None else None
        a.B, a.N, a.C, a.H, a.W = ep.B, ep.N, ep.C, ep.H, ep.W
        a.nbins, a.dof6, a.reward_mode = nbins, int(dof6), mode
        ep.iter_args = (key, a, (rot, tt, cached))
    return ep.iter_args[1]


@torch.no_grad()
def iterate(data, pose_source, action_r, action_t, config, prev_distance=None, with_reward=True, observe=True):
    """Extension: the loop body around the agent's forward pass as ONE library call (cmr_iteration) -
    ``step`` (:179-207) -> ``reward`` of the new pose (:263-302) -> ``observation_from_a_pose`` of the new pose (:25-126).
    Returns (pose_source, reward, p2p_distance, observation_2d, observation_3d); parts that were not asked for are None.
    Bit-identical to the three separate calls."""
    global _iteration_fn
    ep = _episode(data)
    st = _reward_state(data) if with_reward else None
    if with_reward and _reward_mode == "shipped" and st.dist_cached is None:
        reward(pose_source, data, None)                   # memoises the distance of this batch
    args = _iteration_args(ep, st, config)
    B = ep.B
    _lib.require_cuda(pose_source, "pose_source", torch.float32)
    if not pose_source.is_contiguous():
        raise _lib.CmrError("iterate() updates pose_source in place: it must be contiguous")
    dof6 = bool(config.is_6_DoF)
    ar = _actions(action_r, 3 if dof6 else 1, "action_r") if action_r is not None else None
    at = _actions(action_t, 3 if dof6 else 2, "action_t") if action_t is not None else None
    rew = dist = obs2d = obs3d = prev = None
    if with_reward:
        rew = torch.empty(B, 1, 1, device=ep.device, dtype=torch.float32)
        dist = torch.empty(B, 1, 1, device=ep.device, dtype=torch.float32)
        if prev_distance is not None:
            prev = _lib.require_cuda(prev_distance, "prev_distance", torch.float32).reshape(B).contiguous()
    if observe:
        obs2d = torch.empty(ep.obs2d_shape, device=ep.device, dtype=torch.float32)
        obs3d = torch.empty(ep.obs3d_shape, device=ep.device, dtype=torch.float32)
    if _iteration_fn is None:
        _iteration_fn = _lib.bind("cmr_iteration")
    import ctypes
    p = lambda t: t.data_ptr() if t is not None else None   # noqa: E731
    rc = _iteration_fn(ctypes.addressof(args), pose_source.data_ptr(), p(ar), p(at), p(prev), p(rew), p(dist), p(obs2d),
                       p(obs3d), _lib.stream_handle(ep.dev_index))
    if rc:
        _lib.fail(rc, "cmr_iteration")
    if _validate:
        check_fault("iterate")
    return pose_source, rew, dist, obs2d, obs3d


def _refresh_in_place(ep):
    """The per-batch state of ``ep`` recomputed from the tensors it already points at, into the buffers it already owns
    (no allocation, no host read: this is what a reusable captured rollout replays before its first observation)."""
    if _mean_provider == "torch":
        ep.mean.copy_(ep.pc.mean(dim=2))
    else:
        _lib.call("cmr_cloud_mean", _lib.ptr(ep.pc), ep.B, ep.N, _lib.ptr(ep.mean), _lib.stream())
    _lib.call("cmr_episode_prepare", _lib.ptr(ep.overlap), _lib.ptr(ep.feat), ep.B, ep.N, ep.C, _lib.ptr(ep.ws), _lib.stream())


class CapturedRollout:
    """A whole rollout as one CUDA graph (see ``capture_rollout``)."""

    def __init__(self, graph, pose, rewards, distances, obs2d, obs3d, keep, actions_r=None, actions_t=None, reusable=None):
        self.graph, self.pose, self.rewards, self.distances = graph, pose, rewards, distances
        self.observation_2d, self.observation_3d = obs2d, obs3d
        self.actions_r, self.actions_t = actions_r, actions_t
        self._keep = keep
        self._reusable = reusable          # (data, episode state) of a capture made with reusable=True

    def load(self, batch):
        """Another batch of the same shapes (a ``data`` dict as the reference's loader yields it) copied into the
        tensors the graph was captured on; the next ``replay`` prepares and registers it.  Only for captures made
        with ``reusable=True``."""
        if self._reusable is None:
            raise _lib.CmrError("CapturedRollout.load needs a capture made with reusable=True")
        data, ep = self._reusable
        for k in ("pc", "pc_overlap_pred", "pc_geo_feat", "img_geo_feat"):
            if batch[k] is not data[k]:
                data[k].copy_(batch[k], non_blocking=True)
        ep.K.copy_(batch["K"].to(torch.float32), non_blocking=True)
        if torch.is_tensor(data.get("K")) and batch["K"] is not data["K"]:
            data["K"].copy_(batch["K"])
        return self

    def replay(self):
        """Run the rollout again (same inputs, whatever they hold now); the result tensors are rewritten in place:
        ``pose`` [B,4,4] final poses, ``rewards`` / ``distances`` [iters,B,1,1], ``observation_2d/3d`` of the last pose
        observed, ``actions_r`` / ``actions_t`` [iters,B,*] the actions taken."""
        self.graph.replay()
        return self


@torch.no_grad()
def capture_rollout(data, config, actions_r=None, actions_t=None, with_reward=True, policy=None, iters=None,
                    reusable=False):
    """Extension: ``init`` + iterations of observe -> (policy) -> step -> reward (Test_Agent.py:150-170) captured ONCE
    as a CUDA graph - a replay costs one launch on the host instead of every call of every iteration.  Either

    * ``actions_r`` / ``actions_t``: scripted actions, [iters, B, 1|3] / [iters, B, 2|3] int64 on the device (their
      CONTENT may change between replays, like every input tensor's), or
    * ``policy(observation_2d, observation_3d) -> (action_r, action_t)``: an on-device policy, e.g. the reference's
      agent in eval mode (``lambda s2, s3: agent.action_from_logits(*agent(s2, s3)[:2], deterministic=True)``), run
      inside the capture ``iters`` (default ``config.action_num``) times.  It must not synchronise with the host
      (``torch.distributions`` validates its arguments with a host read unless
      ``Distribution.set_default_validate_args(False)``).

    The per-batch state of ``data`` (cloud mean, compacted features, intrinsics) is prepared before the capture and is
    NOT rebuilt by a replay - unless ``reusable=True``: then the graph itself begins with that preparation, and
    ``CapturedRollout.load(batch)`` copies the next batch (same shapes) into the captured tensors - one capture serves
    every batch of a run.  ``reusable`` needs the inputs in the layout the kernels read (float32 contiguous CUDA
    tensors, a bool mask) and ``with_reward=False`` (the inference loop has no reward)."""
    ep = _episode(data)
    dev = ep.device
    if reusable:
        if with_reward:
            raise ValueError("capture_rollout: reusable=True needs with_reward=False")
        same = (ep.pc.data_ptr() == data["pc"].data_ptr() and ep.feat.data_ptr() == data["pc_geo_feat"].data_ptr() and
                ep.img_feat.data_ptr() == data["img_geo_feat"].data_ptr() and
                ep.overlap.data_ptr() == data["pc_overlap_pred"].data_ptr())
        if not same or data.get("_cmr_b200_mean_override") is not None:
            raise _lib.CmrError("capture_rollout(reusable=True): inputs must be float32 contiguous CUDA tensors and a bool "
                                "mask, without a mean override (the graph reads them in place)")
    if (policy is None) == (actions_r is None or actions_t is None):
        raise ValueError("capture_rollout: give either actions_r and actions_t or a policy")
    if policy is None:
        iters = int(actions_r.shape[0])
        actions_r = _lib.require_cuda(actions_r, "actions_r", torch.int64).contiguous()
        actions_t = _lib.require_cuda(actions_t, "actions_t", torch.int64).contiguous()
        taken_r, taken_t = actions_r, actions_t
    else:
        iters = int(config.action_num if iters is None else iters)
        taken_r = taken_t = None
    pose0, _ = init(data)
    if with_reward:
        reward(pose0, data, None)                               # per-batch reward state (and the memoised distance)
    pose = pose0.clone()
    rewards = torch.zeros(iters, ep.B, 1, 1, device=dev)
    distances = torch.zeros(iters, ep.B, 1, 1, device=dev)
    torch.cuda.synchronize(dev)

    def body():
        nonlocal taken_r, taken_t
        if reusable:
            _refresh_in_place(ep)
        pose.copy_(pose0)
        prev = None
        o2 = o3 = None
        for it in range(iters):
            o2, o3 = observation_from_a_pose(data, pose)
            if policy is None:
                a_r, a_t = actions_r[it], actions_t[it]
            else:
                a_r, a_t = policy(o2, o3)
                if taken_r is None:                             # shapes are the policy's: allocated on the first pass
                    taken_r = torch.zeros((iters,) + tuple(a_r.shape), dtype=a_r.dtype, device=dev)
                    taken_t = torch.zeros((iters,) + tuple(a_t.shape), dtype=a_t.dtype, device=dev)
                taken_r[it].copy_(a_r)
                taken_t[it].copy_(a_t)
            step(a_r, a_t, pose, config)
            if with_reward:
                r, prev = reward(pose, data, prev)
                rewards[it].copy_(r)
                distances[it].copy_(prev)
        return o2, o3

    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3 if policy is not None else 1):         # warm-up outside the capture (cuDNN picks its kernels)
            body()
    torch.cuda.current_stream(dev).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        o2, o3 = body()
    return CapturedRollout(graph, pose, rewards, distances, o2, o3, (data, actions_r, actions_t, pose0, policy, ep),
                           taken_r, taken_t, (data, ep) if reusable else None)


def expert(pose_source, targets, config, data=None):
    """environment.py:143-176.  Implemented on the device in
    cmr_expert; see cmr_agent_b200/expert.py."""
    from . import expert as _expert
    return _expert.expert(pose_source, targets, config, data)
