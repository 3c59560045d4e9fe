import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cmr_agent_b200 import synth, environment as env
from tests import helpers as hp
from tests.test_gpu_env import _oracle_obs, _random_poses
dev = torch.device('cuda:0')
for shape, B in ((dict(num_pt=4096, img_h=64, img_w=256), 2), (dict(num_pt=40960, img_h=160, img_w=512), 2)):
    H, W = shape['img_h']//4, shape['img_w']//4
    data_cpu = synth.make_batch(B, seed=2023, **shape)
    data = hp.to_device(data_cpu, dev)
    pose = torch.eye(4).repeat(B,1,1)
    o2, o3, pix, mvis = env.observation_from_a_pose(data, pose.to(dev), return_pixels=True)
    torch.cuda.synchronize()
    mean = data["_cmr_b200_episode"][1].mean.cpu()
    wpix, winc, wproj = _oracle_obs(data_cpu, pose, mean, H, W)
    got = o2[:, 64:].cpu()
    print(shape, 'pix equal', torch.equal(pix.cpu(), wpix), 'mvis', mvis.cpu().tolist())
    for b in range(B):
        ov = data_cpu['pc_overlap_pred'][b]
        ids = wpix[b][ov]
        cnt = torch.bincount(ids[ids < H*W], minlength=H*W)
        d = (got[b] != wproj[b]).any(dim=0).flatten()
        bad = d.nonzero().flatten()
        occupied = (cnt > 0).nonzero().flatten()
        print(' ep', b, 'occupied pixels', len(occupied), 'bad pixels', len(bad), 'bad tiles', sorted(set((bad // 128).tolist())))
        gotnz = (got[b].abs().amax(dim=0).flatten() > 0)
        print('   got nonzero pixels', int(gotnz.sum()), 'of which in occupied', int((gotnz & (cnt > 0)).sum()))
        for p in bad[:6].tolist():
            g = got[b].reshape(64, -1)[:3, p].tolist(); w = wproj[b].reshape(64, -1)[:3, p].tolist()
            print('   pixel', p, 'tile', p // 128, 'pl&7', (p % 128) & 7, 'cnt', int(cnt[p]), 'got', g, 'want', w)
