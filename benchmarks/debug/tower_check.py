"""Block-by-block check of the tcgen05 tower against oracle/tower_oracle.py (debugging aid, GPU box).
Reads the kernel's own intermediates back from the workspace: max keys of every block, feat2 (planes 2,3),
feat3 (planes 0,1).   python benchmarks/debug/tower_check.py [B] [N]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cmr_agent_b200 import agent_tower, synth  # noqa: E402
from oracle import tower_oracle as to  # noqa: E402


def key2f(k):
    k = k.astype(np.uint32)
    bits = np.where(k & 0x80000000, k ^ 0x80000000, ~k).astype(np.uint32)
    return torch.from_numpy(bits.view(np.float32).copy())


def piece_plane(ws, idx, plane_bytes, B, N):
    raw = ws[idx * plane_bytes: idx * plane_bytes + B * N * 128].cpu().numpy()
    return torch.from_numpy(raw.view(np.float16).astype(np.float32)).reshape(B, N, 64)      # fp16 pieces (CMR_TOWER_FMT=1)


def scaled(got, want):
    return float((got - want).abs().max() / want.abs().max())


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    dev = torch.device("cuda:0")
    states = [to.make_state(2023 + 40 + i, cin, cout) for i, (cin, cout) in enumerate(to.TOWER)]
    data = synth.make_batch(B, seed=2044, num_pt=N, img_h=160, img_w=512, with_features=False) if False else None
    g = torch.Generator().manual_seed(5)
    xyz = (torch.rand(B, 3, N, generator=g) - 0.5) * 120
    flags = (torch.rand(B, 2, N, generator=g) < 0.3).float()
    obs3d = torch.cat([xyz, flags], dim=1).contiguous()
    tower = agent_tower.Tower3D(states, dev)
    out = tower(obs3d.to(dev))
    torch.cuda.synchronize()
    ws = tower._ws
    plane_bytes = (B * N * 128 + 1023) // 1024 * 1024
    keys = ws[4 * plane_bytes: 4 * plane_bytes + B * 320 * 4].cpu().numpy().view(np.uint32)
    k1, k2, k3 = (key2f(keys[i * B * 64:(i + 1) * B * 64]).reshape(B, 64) for i in range(3))
    k4 = key2f(keys[3 * B * 64:]).reshape(B, 128)
    # oracle, block by block
    f1 = to.block(states[0], obs3d, None); m1 = f1.max(dim=2)[0]
    f2 = to.block(states[1], f1, m1); m2 = f2.max(dim=2)[0]
    f3 = to.block(states[2], f2, m2); m3 = f3.max(dim=2)[0]
    f4 = to.block(states[3], f3, m3); m4 = f4.max(dim=2)[0]
    print("max1", scaled(k1, m1), "max2", scaled(k2, m2), "max3", scaled(k3, m3), "max4", scaled(k4, m4))
    feat2 = piece_plane(ws, 2, plane_bytes, B, N) + piece_plane(ws, 3, plane_bytes, B, N)
    feat3 = piece_plane(ws, 0, plane_bytes, B, N) + piece_plane(ws, 1, plane_bytes, B, N)
    print("feat2", scaled(feat2, f2.permute(0, 2, 1)), "feat3", scaled(feat3, f3.permute(0, 2, 1)))
    print("out", scaled(out.cpu(), m4), "finite", bool(torch.isfinite(out).all()))
    bad = (feat2 - f2.permute(0, 2, 1)).abs().max(dim=2)[0]
    print("feat2 worst rows per episode:", [int(bad[b].argmax()) for b in range(B)], float(bad.max()))


if __name__ == "__main__":
    main()
