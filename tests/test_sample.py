"""``environment.sample_image_features`` - image features sampled bilinearly at the points' projections (BASELINE.json
north_star; an extension, SURVEY.md D1: the reference has no point-side gather and names ``F.grid_sample`` as the oracle).

* CPU: the operation-by-operation statement (oracle/sample_oracle.sample_image_features) against torch's own
  ``F.grid_sample(mode="bilinear", padding_mode="zeros", align_corners=True)`` - tolerance 2e-5 of the features' scale
  (grid_sample re-derives the pixel coordinate from the normalised one: two more roundings of a coordinate up to 127
  move a weight by ~1e-5), identical frustum masks, exact zeros outside.
* GPU: the CUDA kernel against that statement BIT FOR BIT, on KITTI-size and ragged shapes, plus the properties that do
  not need an oracle (a constant image samples to the constant; an image linear in x and y samples to the same linear
  function of (u, v))."""
import pytest
import torch

from cmr_agent_b200 import synth
from oracle import env_oracle, sample_oracle

from tests.helpers import to_device


def _case(B, N, img_h, img_w, seed, at_gt=True, C=64):
    data = synth.make_batch(B, seed=seed, num_pt=N, img_h=img_h, img_w=img_w)
    if C != 64:
        g = torch.Generator().manual_seed(seed + 9)
        data["img_geo_feat"] = torch.randn(B, C, img_h // 4, img_w // 4, generator=g)
        data["pc_geo_feat"] = torch.randn(B, C, N, generator=g)
    pose = torch.eye(4).repeat(B, 1, 1)
    if at_gt:      # the ground-truth registration puts the cloud in view (identity leaves a few hundred points)
        pose = env_oracle.to_disentangled(data["P"].clone(), data["pc"])
    return data, pose.contiguous()


@pytest.mark.parametrize("B,N,img_h,img_w,at_gt", [
    (2, 4096, 160, 512, True),
    (1, 3001, 64, 256, True),
    (2, 2048, 160, 512, False),
])
def test_statement_matches_grid_sample(B, N, img_h, img_w, at_gt):
    data, pose = _case(B, N, img_h, img_w, 11, at_gt)
    got, cam = sample_oracle.sample_image_features(data, pose)
    want, cam2 = sample_oracle.grid_sample_reference(data, pose)
    assert torch.equal(cam, cam2)
    assert int(cam.sum()) > 50, "the case must put points inside the frustum"
    scale = float(data["img_geo_feat"].abs().max())
    assert float((got - want).abs().max()) <= 2e-5 * scale
    assert bool((got[~cam.unsqueeze(1).expand_as(got)] == 0).all())


def test_statement_at_integer_pixels_returns_the_pixel():
    """(u, v) exactly on a pixel centre - also on the last row / column, where the right / lower neighbours do not
    exist - samples that pixel's features exactly."""
    B, C, H, W = 1, 4, 5, 7
    img = torch.arange(B * C * H * W, dtype=torch.float32).reshape(B, C, H, W)
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    # points straight in front of a camera whose intrinsics are the identity: u = x / z, v = y / z with z = 1
    pc = torch.stack([xs.reshape(-1).float(), ys.reshape(-1).float(), torch.ones(H * W)]).unsqueeze(0)
    data = {"pc": pc, "K": torch.eye(3).unsqueeze(0), "img": torch.zeros(1, 3, 4 * H, 4 * W), "img_geo_feat": img}
    got, cam = sample_oracle.sample_image_features(data, torch.eye(4).unsqueeze(0), mean=torch.zeros(1, 3, 1))
    assert bool(cam.all())
    assert torch.equal(got, img.reshape(B, C, H * W))


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,img_h,img_w,at_gt,C", [
    (2, 40960, 160, 512, True, 64),       # KITTI
    (2, 40960, 160, 512, False, 64),      # identity pose: most points outside the frustum
    (3, 3001, 64, 256, True, 64),         # N % 4 != 0: the scalar store path, a ragged last warp
    (1, 8192, 160, 320, True, 128),       # two channel slabs
    (2, 1000, 36, 100, True, 8),          # a partial slab, a 9 x 25 grid
    (1, 40, 64, 64, True, 64),            # fewer than 45 points: the unfused products of a small bmm
])
def test_gpu_matches_statement_bit_for_bit(cuda, B, N, img_h, img_w, at_gt, C):
    from cmr_agent_b200 import environment as env
    data, pose = _case(B, N, img_h, img_w, 5, at_gt, C)
    want, want_cam = sample_oracle.sample_image_features(data, pose)
    dd = to_device(data, cuda)
    # the same fp32 cloud mean on both sides (the kernel's own mean is an fp64 sum in a fixed order, torch's CPU mean is
    # not: one ulp of the mean moves u by an ulp - invisible to the rounded pixel ids, visible to bilinear weights)
    dd["_cmr_b200_mean_override"] = env_oracle.cloud_mean(data["pc"]).reshape(B, 3)
    got, cam = env.sample_image_features(dd, pose.to(cuda))
    torch.cuda.synchronize()
    assert torch.equal(cam.cpu(), want_cam)
    assert torch.equal(got.cpu(), want), f"max difference {float((got.cpu() - want).abs().max())}"
    again, _ = env.sample_image_features(dd, pose.to(cuda))     # the cached pixel-major image is reused
    assert torch.equal(again, got)


@pytest.mark.gpu
def test_gpu_constant_and_linear_images(cuda):
    from cmr_agent_b200 import environment as env
    data, pose = _case(2, 40960, 160, 512, 7, True)
    H, W = 40, 128
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    img = torch.zeros(2, 64, H, W)
    img[:, 0] = 3.25                       # constant
    img[:, 1] = xs                         # linear in x
    img[:, 2] = ys                         # linear in y
    img[:, 3] = 2.0 * xs - 0.5 * ys + 1.0
    data["img_geo_feat"] = img
    dd = to_device(data, cuda)
    dd["_cmr_b200_mean_override"] = env_oracle.cloud_mean(data["pc"]).reshape(2, 3)
    got, cam = env.sample_image_features(dd, pose.to(cuda))
    u, v, cam_ref = sample_oracle.project_all(data, pose)
    got, cam = got.cpu(), cam.cpu()
    assert torch.equal(cam, cam_ref) and int(cam.sum()) > 1000
    m = cam
    assert float((got[:, 0][m] - 3.25).abs().max()) <= 1e-6, "the weights of a point add up to one (to a few ulps)"
    assert float((got[:, 1][m] - u[m]).abs().max()) <= 2e-5 * W
    assert float((got[:, 2][m] - v[m]).abs().max()) <= 2e-5 * H
    assert float((got[:, 3][m] - (2.0 * u[m] - 0.5 * v[m] + 1.0)).abs().max()) <= 1e-4 * W
    assert bool((got[:, :4][~m.unsqueeze(1).expand(-1, 4, -1)] == 0).all())


@pytest.mark.gpu
def test_gpu_sample_rejects_bad_arguments(cuda):
    from cmr_agent_b200 import _lib
    from cmr_agent_b200 import environment as env
    data, pose = _case(1, 512, 64, 64, 3)
    dd = to_device(data, cuda)
    with pytest.raises(_lib.CmrError):
        env.sample_image_features(dd, torch.eye(4, device=cuda).repeat(2, 1, 1))      # batch mismatch
