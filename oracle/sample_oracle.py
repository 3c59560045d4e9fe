"""CPU restatement of ``environment.sample_image_features`` (TEST INFRASTRUCTURE ONLY).

The reference has no point-side image gather (its observation scatters point features onto the pixel grid,
environment/environment.py:67-83); BASELINE.json's north_star asks for one ("bilinearly sample image features onto
visible points"), and SURVEY.md D1 names its oracle: ``F.grid_sample(mode="bilinear", padding_mode="zeros",
align_corners=True)``.  The operator is therefore specified here, in two forms that tests/test_sample.py holds against
each other and against the CUDA kernel:

* ``sample_image_features``: the statement the kernel follows operation by operation - pixel coordinates and frustum
  mask exactly as the reference projects points (environment.py:54-65 through oracle/env_oracle.py, itself pinned to
  the reference), then x0 = floor(u), dx = u - x0, the four weights (1-dx)(1-dy), dx(1-dy), (1-dx)dy, dx dy, and the
  four products added left to right, every operation rounded to fp32.  The kernel must equal it BIT FOR BIT.
* ``grid_sample_reference``: torch's own ``F.grid_sample`` on the same pixel coordinates (normalised for
  align_corners=True).  Its un-normalisation re-derives the pixel coordinate with two more roundings, so the two forms
  agree to ~1e-6 of the features' scale, not to the bit; the test states the tolerance.
parity unpinned against the reference (there is nothing in it to pin to); pinned against torch's grid_sample.
"""
import torch
import torch.nn.functional as F

from . import env_oracle


@torch.no_grad()
def project_all(data, RT, mean=None):
    """(u [B,N], v [B,N], in_cam [B,N]) of ALL points: environment.py:91-101."""
    pc = data["pc"]
    H = data["img"].shape[2] // 4
    W = data["img"].shape[3] // 4
    if mean is None:
        mean = env_oracle.cloud_mean(pc)
    cam = env_oracle.disentangled_transform(pc, mean, RT)
    uvz, in_cam = env_oracle.project_pinhole(data["K"], cam, H, W)
    return uvz[:, 0, :], uvz[:, 1, :], in_cam


@torch.no_grad()
def sample_image_features(data, RT, mean=None):
    """-> (feats [B,C,N] f32, in_cam [B,N] bool)."""
    img = data["img_geo_feat"]
    B, C, H, W = img.shape
    u, v, in_cam = project_all(data, RT, mean)
    u = torch.where(in_cam, u, torch.zeros_like(u))      # outside: any valid address, the result is zeroed below
    v = torch.where(in_cam, v, torch.zeros_like(v))
    fx, fy = torch.floor(u), torch.floor(v)
    x0, y0 = fx.long(), fy.long()
    dx, dy = u - fx, v - fy
    ex, ey = 1.0 - dx, 1.0 - dy
    w00, w01, w10, w11 = ex * ey, dx * ey, ex * dy, dx * dy
    right, below = (x0 + 1 < W), (y0 + 1 < H)
    flat = img.reshape(B, C, H * W)

    def take(yy, xx, ok):
        idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).unsqueeze(1).expand(B, C, -1)
        return torch.gather(flat, 2, idx) * ok.unsqueeze(1).to(flat.dtype)

    a = take(y0, x0, torch.ones_like(right))
    b = take(y0, x0 + 1, right)
    c = take(y0 + 1, x0, below)
    d = take(y0 + 1, x0 + 1, right & below)
    out = ((a * w00.unsqueeze(1) + b * w01.unsqueeze(1)) + c * w10.unsqueeze(1)) + d * w11.unsqueeze(1)
    out = torch.where(in_cam.unsqueeze(1), out, torch.zeros_like(out))
    return out, in_cam


@torch.no_grad()
def grid_sample_reference(data, RT, mean=None):
    """The same operator through torch.nn.functional.grid_sample (SURVEY.md D1's oracle)."""
    img = data["img_geo_feat"]
    B, C, H, W = img.shape
    u, v, in_cam = project_all(data, RT, mean)
    gx = 2.0 * u / (W - 1) - 1.0
    gy = 2.0 * v / (H - 1) - 1.0
    grid = torch.stack([gx, gy], dim=-1).unsqueeze(1)            # [B,1,N,2]
    grid = torch.where(in_cam.unsqueeze(1).unsqueeze(-1), grid, torch.full_like(grid, -2.0))   # outside: zero padding
    out = F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=True)[:, :, 0, :]
    return out * in_cam.unsqueeze(1).to(out.dtype), in_cam
