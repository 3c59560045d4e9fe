"""CPU restatement of the cost-volume warp of models/IterModel.py:272-351 (TEST INFRASTRUCTURE ONLY).

The reference evaluates these lines inline in ``IterModel.forward`` with hard-coded ``.cuda()`` calls and there is no
callable boundary to import; this file restates the same torch expressions, line by line, on the CPU (torch_scatter
through oracle/shims.py, the same stand-in that is pinned for the environment path).  PINNED by
tests/golden/cost_volume.npz: the outputs of the reference's own statements (taken from its syntax tree and run on the
CPU with ``Tensor.cuda`` as the identity by tests/golden/make_golden.py), reproduced bit for bit by ``warp`` below
(tests/test_cost_volume.py::test_oracle_matches_reference_golden).  Differences from the reference text: tensors stay where
they are; the dump bin is H*W instead of the literal 5120 (:311, equal for KITTI's 40 x 128); poses are processed in
one piece instead of chunks of 200 (:324-346; the chunks are independent).
"""
import torch

from . import shims


def warp(pc, pc_mask, delta_RT, K, pc_geo_feat, scores, H, W):
    ts = shims   # scatter_mean / scatter_sum: the stand-ins for torch_scatter (behaviour restated in oracle/shims.py)
    pc = pc.unsqueeze(1)                                                       # :279
    pc_RT = delta_RT[:, :, 0:3, 0:3] @ pc + delta_RT[:, :, 0:3, 3:4]           # :281
    Km = K.unsqueeze(1)                                                        # :283
    pc_RT_K = Km @ pc_RT                                                       # :285
    pc_RT_K[:, :, 0:2, :] = pc_RT_K[:, :, 0:2, :] / pc_RT_K[:, :, 2:3, :]      # :286
    in_cam = (pc_RT_K[:, :, 0, :] >= 0) & (pc_RT_K[:, :, 0, :] <= (W - 1)) & \
             (pc_RT_K[:, :, 1, :] >= 0) & (pc_RT_K[:, :, 1, :] <= (H - 1)) & (pc_RT_K[:, :, 2, :] > 0)   # :293-297
    in_cam = in_cam[:, :, pc_mask]                                             # :302
    pix = pc_RT_K[:, :, 0:2, :].round().int()                                  # :304
    feat = pc_geo_feat[:, :, pc_mask].unsqueeze(1).repeat(1, delta_RT.shape[1], 1, 1)   # :306-311
    pix = pix[:, :, :, pc_mask].permute(0, 1, 3, 2)                            # :313-314
    idx = pix[:, :, :, 1] * W + pix[:, :, :, 0]                                # :316
    idx[~in_cam] = H * W                                                       # :318
    sc = scores[:, pc_mask].unsqueeze(1).repeat(1, in_cam.shape[1], 1)         # :321-323
    sc[~in_cam] = 0.0                                                          # :325
    feat = torch.cat([feat, torch.zeros_like(feat[:, :, :, 0:1])], dim=-1)     # :331-332
    sc = torch.cat([sc, torch.zeros_like(sc[:, :, 0:1])], dim=-1)              # :334-336
    idx = torch.cat([idx.long(), torch.ones_like(idx[:, :, 0:1]).long() * (H * W)], dim=-1)   # :338-340
    wf = ts.scatter_mean(feat, idx.unsqueeze(2).repeat(1, 1, feat.shape[2], 1), dim=3)         # :341
    occ = ts.scatter_sum(sc, idx, dim=2)                                       # :343
    return wf[:, :, :, :H * W], occ[:, :, :H * W]                              # :350-351
