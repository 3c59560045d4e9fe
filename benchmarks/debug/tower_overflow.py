"""Debug: what the tower's intermediates look like when an activation leaves the fp16 range."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cmr_agent_b200 import agent_tower, _lib
from oracle import tower_oracle as to
dev = torch.device("cuda:0")
B, N = 1, 256
states = [to.make_state(321 + i, ci, co) for i, (ci, co) in enumerate(to.TOWER)]
g = torch.Generator().manual_seed(5)
xyz = (torch.rand(B, 3, N, generator=g) - 0.5) * 160 * float(sys.argv[1] if len(sys.argv) > 1 else 1e5)
obs3d = torch.cat([xyz, (torch.rand(B, 2, N, generator=g) < 0.3).float()], 1).contiguous()
tower = agent_tower.Tower3D(states, dev)
out = tower(obs3d.to(dev)); torch.cuda.synchronize()
ws = tower._ws
pb = (B * N * 128 + 1023) // 1024 * 1024
keys = ws[5 * pb: 5 * pb + B * 320 * 4].cpu().numpy().view(np.uint32)
print("keys1", [hex(k) for k in keys[:4]], "keys2", [hex(k) for k in keys[64:68]], "keys3", [hex(k) for k in keys[128:132]], "keys4", [hex(k) for k in keys[192:196]])
def plane(i): return ws[i * pb: i * pb + B * N * 128].cpu().numpy().view(np.float16).reshape(B, N, 64)
print("feat3 hi row0", plane(0)[0, 0, :6], "lo", plane(1)[0, 0, :6], "lo2", plane(4)[0, 0, :6])
print("feat2 hi row0", plane(2)[0, 0, :6], "lo", plane(3)[0, 0, :6])
print("out", out[0, :6].cpu(), "fault", _lib.take_fault())
