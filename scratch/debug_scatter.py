import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cmr_agent_b200 import synth, environment as env
from tests import helpers as hp
from tests.test_gpu_env import _oracle_obs, _random_poses
dev = torch.device('cuda:0')
shape = dict(num_pt=40960, img_h=160, img_w=512); B = 4; H, W = 40, 128
data_cpu = synth.make_batch(B, seed=77, **shape)
data = hp.to_device(data_cpu, dev)
for trial in range(3):
    pose = _random_poses(B, 100 + trial, scale_t=2.0 if trial else 0.0)
    o2, o3, pix, mvis = env.observation_from_a_pose(data, pose.to(dev), return_pixels=True)
    torch.cuda.synchronize()
    mean = data["_cmr_b200_episode"][1].mean.cpu()
    wpix, winc, wproj = _oracle_obs(data_cpu, pose, mean, H, W)
    got = o2[:, 64:].cpu()
    diff = (got != wproj)
    print('trial', trial, 'pix equal', torch.equal(pix.cpu(), wpix), 'mismatch elems', int(diff.sum()), 'mvis', mvis.cpu().tolist())
    for b in range(B):
        ov = data_cpu['pc_overlap_pred'][b]
        ids = wpix[b][ov]
        cnt = torch.bincount(ids[ids < H*W], minlength=H*W).view(H, W)
        d = diff[b].any(dim=0)  # [H,W]
        rows = d.any(dim=1).nonzero().flatten().tolist()
        print('  ep', b, 'bad rows', rows, 'row point counts', [int(cnt[r].sum()) for r in rows], 'max row count', int(cnt.sum(1).max()))
        for r in rows[:2]:
            cols = d[r].nonzero().flatten().tolist()
            print('    row', r, 'bad cols', cols[:20], 'cnt', [int(cnt[r, c]) for c in cols[:20]])
            c0 = cols[0]
            print('    got', got[b, :4, r, c0].tolist(), 'want', wproj[b, :4, r, c0].tolist())
