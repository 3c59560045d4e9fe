"""Cost-volume warp (models/IterModel.py:272-351): the CUDA path (cmr_cost_volume_* through
cmr_agent_b200.cost_volume.warp) against the CPU restatement of the reference's torch expressions
(oracle/cost_volume_oracle.py) and against tests/golden/cost_volume.npz - the outputs of the reference's OWN
statements (IterModel.py:96-172 and :272-351, compiled from its syntax tree by tests/golden/make_golden.py with
``Tensor.cuda`` as the identity), which is what pins the oracle.
Bar: warped features and occupancy BIT-EXACT (ordered sums), which implies identical pixel indices and masks."""
import hashlib
import math
import os
import sys

import numpy as np
import pytest
import torch

from cmr_agent_b200 import synth
from oracle import cost_volume_oracle as cvo

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden  # noqa: E402  (its cost_volume_inputs() regenerates the seeded inputs)

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cost_volume.npz"))
GOLDEN_CASES = tuple(make_golden.COST_VOLUME_CASES)


def _sha(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(t.numpy()).tobytes())
    return h.hexdigest().encode()


def _golden_inputs(case):
    data, pc_i, mask, scores, nlabel, _ = make_golden.cost_volume_inputs(case)
    assert _sha(data["pc"], data["pc_geo_feat"], mask, data["K"], pc_i, scores) == \
        GOLDEN[case + "_inputs_sha"].tobytes(), "seeded inputs differ from the fixture's"
    return data, pc_i, mask[0], scores, torch.from_numpy(GOLDEN[case + "_poses"])


def _check_golden(case, wf, occ):
    assert torch.equal(occ, torch.from_numpy(GOLDEN[case + "_occupancy"])), "occupancy differs from the reference's"
    rows = torch.from_numpy(GOLDEN[case + "_sample_rows"])
    got = wf.permute(0, 1, 3, 2).reshape(-1, wf.shape[2])[rows]
    assert torch.equal(got, torch.from_numpy(GOLDEN[case + "_sample_features"])), "sampled pixels differ"
    assert _sha(wf.contiguous()) == GOLDEN[case + "_features_sha"].tobytes(), "warped features differ from the reference's"


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_matches_reference_golden(case):
    """Pins the oracle: the restatement reproduces, bit for bit, what the reference's own lines computed."""
    data, pc_i, mask, scores, poses = _golden_inputs(case)
    wf, occ = cvo.warp(pc_i, mask, poses, data["K"], data["pc_geo_feat"], scores, 40, 128)
    _check_golden(case, wf, occ)


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_gpu_matches_reference_golden(cuda, case):
    from cmr_agent_b200 import cost_volume
    data, pc_i, mask, scores, poses = _golden_inputs(case)
    wf, occ = cost_volume.warp(pc_i.to(cuda), mask.to(cuda), poses.to(cuda), data["K"], data["pc_geo_feat"].to(cuda),
                               scores.to(cuda), 40, 128)
    torch.cuda.synchronize()
    _check_golden(case, wf.cpu(), occ.cpu())


def _poses(B, nlabel, r_amp, t_amp, seed):
    """delta_RT_inv[:, :, 0:3, :] of IterModel.sample_poses (:133-172): nlabel^3 poses, rotation about y, shifts in x, z."""
    g = torch.Generator().manual_seed(seed)
    base = torch.linspace(-(nlabel - 1) / 2, (nlabel - 1) / 2, nlabel)
    out = torch.zeros(B, nlabel, nlabel, nlabel, 4, 4)
    for b in range(B):
        jitter = torch.rand(3, generator=g) * 0.3
        for i, ry in enumerate(base * (2 * r_amp / (nlabel - 1)) + jitter[0] * 0.01):
            c, s = math.cos(float(ry)), math.sin(float(ry))
            R = torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])
            for j, tx in enumerate(base * (2 * t_amp / (nlabel - 1)) + jitter[1]):
                for k, tz in enumerate(base * (2 * t_amp / (nlabel - 1)) + jitter[2]):
                    m = torch.eye(4)
                    m[:3, :3] = R
                    m[0, 3], m[2, 3] = tx, tz
                    out[b, i, j, k] = torch.linalg.inv(m)
    return out.view(B, -1, 4, 4)[:, :, 0:3, :].contiguous()


def _case(B, N, img_h, img_w, nlabel, seed, dense=False):
    data = synth.make_batch(B, seed=seed, num_pt=N, img_h=img_h, img_w=img_w)
    g = torch.Generator().manual_seed(seed + 1)
    scores = torch.rand(B, N, generator=g)
    mask = data["pc_overlap_pred"][0].clone()
    if dense:
        mask[:] = True
    # the sampled poses perturb the GROUND-TRUTH registration, so that the cloud is in view
    gt = data["P"][:, 0:3, :]
    d = _poses(B, nlabel, 0.15, 2.0, seed)
    R = d[:, :, :, 0:3] @ gt[:, None, :, 0:3]
    t = d[:, :, :, 0:3] @ gt[:, None, :, 3:4] + d[:, :, :, 3:4]
    poses = torch.cat([R, t], dim=-1).contiguous()
    return data, mask, poses, scores


@pytest.mark.parametrize("B,N,img_h,img_w,nlabel,dense", [
    (1, 4096, 160, 512, 3, False),        # KITTI grid, 27 poses
    (2, 3001, 64, 256, 2, False),         # two clouds share the first one's mask; N % 4 != 0
    (1, 8192, 160, 320, 3, True),         # NuScenes grid, every point masked in: dense buckets
])
def test_oracle_shapes_and_occupancy(B, N, img_h, img_w, nlabel, dense):
    data, mask, poses, scores = _case(B, N, img_h, img_w, nlabel, 3, dense)
    H, W = img_h // 4, img_w // 4
    wf, occ = cvo.warp(data["pc"], mask, poses, data["K"], data["pc_geo_feat"], scores, H, W)
    assert tuple(wf.shape) == (B, nlabel ** 3, 64, H * W) and tuple(occ.shape) == (B, nlabel ** 3, H * W)
    assert float(occ.sum()) > 0 and float(wf.abs().sum()) > 0          # the cloud is in view
    assert bool(((occ > 0) == (wf.abs().sum(dim=2) > 0)).all())        # a pixel has features iff it has points


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,img_h,img_w,nlabel,dense", [
    (1, 4096, 160, 512, 3, False),
    (2, 3001, 64, 256, 2, False),
    (1, 8192, 160, 320, 3, True),
    (1, 40960, 160, 512, 3, False),       # the reference's sizes (27 of its 729 poses)
    (2, 40960, 64, 64, 2, True),          # 8 buckets for thousands of visible points: buffers overflow, shared clouds
    (1, 2048, 36, 100, 3, True),          # 9 x 25 grid: P % 4 != 0, results leave without TMA
])
def test_gpu_matches_oracle(cuda, B, N, img_h, img_w, nlabel, dense):
    from cmr_agent_b200 import cost_volume
    data, mask, poses, scores = _case(B, N, img_h, img_w, nlabel, 3, dense)
    H, W = img_h // 4, img_w // 4
    want_f, want_o = cvo.warp(data["pc"], mask, poses, data["K"], data["pc_geo_feat"], scores, H, W)
    got_f, got_o = cost_volume.warp(data["pc"].to(cuda), mask.to(cuda), poses.to(cuda), data["K"],
                                    data["pc_geo_feat"].to(cuda), scores.to(cuda), H, W)
    torch.cuda.synchronize()
    assert torch.equal(got_o.cpu(), want_o), "occupancy (ordered sum of the in-camera scores) differs"
    assert torch.equal(got_f.cpu(), want_f), "warped features (ordered scatter-mean) differ"
    # a second call on fresh buffers gives the same bits (the counters are left clean)
    again_f, again_o = cost_volume.warp(data["pc"].to(cuda), mask.to(cuda), poses.to(cuda), data["K"],
                                        data["pc_geo_feat"].to(cuda), scores.to(cuda), H, W)
    assert torch.equal(again_f, got_f) and torch.equal(again_o, got_o)


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,img_h,img_w,nlabel,dense", [
    (1, 40960, 160, 512, 3, False),
    (2, 40960, 64, 64, 2, True),
])
def test_gpu_bucket_path_matches_oracle(cuda, monkeypatch, B, N, img_h, img_w, nlabel, dense):
    """The reference's shape takes the one-CTA-per-pose kernel (cost_volume_kernels.cuh); every other shape - more
    channels, a grid that is not whole 32-pixel buckets - takes the observation's bucket kernels.  CMR_B200_CV=buckets
    sends the same cases down that path: the two must agree with the oracle, hence with each other, bit for bit."""
    from cmr_agent_b200 import cost_volume
    monkeypatch.setenv("CMR_B200_CV", "buckets")
    data, mask, poses, scores = _case(B, N, img_h, img_w, nlabel, 3, dense)
    H, W = img_h // 4, img_w // 4
    want_f, want_o = cvo.warp(data["pc"], mask, poses, data["K"], data["pc_geo_feat"], scores, H, W)
    got_f, got_o = cost_volume.warp(data["pc"].to(cuda), mask.to(cuda), poses.to(cuda), data["K"],
                                    data["pc_geo_feat"].to(cuda), scores.to(cuda), H, W)
    assert torch.equal(got_o.cpu(), want_o) and torch.equal(got_f.cpu(), want_f)


@pytest.mark.gpu
@pytest.mark.parametrize("N", [40960, 8192])     # lists in global memory / in shared memory
def test_gpu_cloud_collapsed_onto_few_pixels(cuda, N):
    """A cloud seen from far away: thousands of points on a handful of pixels (the bitmap ordering of
    k_cost_volume_sort; a quadratic ranking would take seconds here)."""
    from cmr_agent_b200 import cost_volume
    data, mask, poses, scores = _case(1, N, 160, 512, 2, 5, True)
    poses = poses.clone()
    poses[:, :, 2, 3] += 4000.0          # push the cloud 4 km down the optical axis
    H, W = 40, 128
    want_f, want_o = cvo.warp(data["pc"], mask, poses, data["K"], data["pc_geo_feat"], scores, H, W)
    assert int((want_o > 0).sum(dim=-1).max()) <= 64 and float(want_o.sum()) > 1000    # a few pixels hold everything
    got_f, got_o = cost_volume.warp(data["pc"].to(cuda), mask.to(cuda), poses.to(cuda), data["K"],
                                    data["pc_geo_feat"].to(cuda), scores.to(cuda), H, W)
    assert torch.equal(got_o.cpu(), want_o) and torch.equal(got_f.cpu(), want_f)


def test_cost_volume_rejects_cpu_tensors():
    from cmr_agent_b200 import _lib, cost_volume
    data, mask, poses, scores = _case(1, 256, 64, 64, 2, 1)
    with pytest.raises(_lib.CmrError):
        cost_volume.warp(data["pc"], mask, poses, data["K"], data["pc_geo_feat"], scores, 16, 16)
