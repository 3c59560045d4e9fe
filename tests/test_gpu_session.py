"""The native rollout session (cmr_session_*, include/cmr_b200.h) against the drop-in functions: same kernels,
so rewards, distances, poses and the last observation must agree BIT FOR BIT - with one slot, with the uploads of
the next rollout overlapping the kernels of the previous one (depth 2, 3), and with the features left on the device."""
import pytest
import torch

from cmr_agent_b200 import synth
from tests import helpers as hp

pytestmark = pytest.mark.gpu


def _drop_in_rollout(env, data_cpu, a_r, a_t, cfg, dev, iters):
    data = hp.to_device(data_cpu, dev)
    pose, target = env.init(data)
    env.to_disentangled(target, data["pc"])
    rews, dists, prev = [], [], None
    for it in range(iters):
        o2, o3 = env.observation_from_a_pose(data, pose)
        env.step(a_r[it].to(dev), a_t[it].to(dev), pose, cfg)
        rew, prev = env.reward(pose, data, prev)
        rews.append(rew.flatten())
        dists.append(prev.flatten())
    torch.cuda.synchronize()
    return torch.stack(rews).cpu(), torch.stack(dists).cpu(), pose.cpu(), target.cpu(), o2, o3


@pytest.mark.parametrize("depth,resident,dof6", [(1, False, False), (2, False, False), (3, True, False), (2, False, True)])
def test_session_equals_the_drop_in_loop(cuda, depth, resident, dof6):
    from cmr_agent_b200 import environment as env, session
    shape = dict(num_pt=8192, img_h=160, img_w=512)
    B, iters = 3, 4
    cfg_h, cfg_d = synth.StepConfig(is_6_DoF=dof6), synth.StepConfig(device=cuda, is_6_DoF=dof6)
    batches = [synth.make_batch(B, first_episode=10 * k, seed=hp.SEED, **shape) for k in range(4)]
    actions = [synth.make_actions(B, iters, seed=77 + k, dof6=dof6) for k in range(4)]
    want = [_drop_in_rollout(env, b, a[0], a[1], cfg_d, cuda, iters) for b, a in zip(batches, actions)]
    ses = session.RolloutSession(B, 8192, 64, 40, 128, iters, cfg_h, depth=depth, features_resident=resident, device=cuda)
    pinned = []
    for b in batches:
        d = {k: (v.pin_memory() if torch.is_tensor(v) and k != "img" else v) for k, v in b.items()}
        if resident:
            d["pc_geo_feat"], d["img_geo_feat"] = b["pc_geo_feat"].to(cuda), b["img_geo_feat"].to(cuda)
        pinned.append(d)
    tickets = []
    got = {}
    for k, (d, a) in enumerate(zip(pinned, actions)):
        tickets.append(ses.submit(d, a[0].pin_memory(), a[1].pin_memory()))
        if k >= depth - 1:                                   # keep `depth` rollouts in flight
            j = k - (depth - 1)
            got[j] = ses.wait(tickets[j])
    for j in range(len(batches)):
        if j not in got:
            got[j] = ses.wait(tickets[j])
    last = len(batches) - 1
    o2, o3 = ses.last_observation(tickets[last])
    torch.cuda.synchronize()
    for j, w in enumerate(want):
        g = got[j]
        assert torch.equal(g[0], w[0]) and torch.equal(g[2], w[2]) and torch.equal(g[3], w[3]), f"rollout {j}"
        assert torch.equal(g[1], w[1]), f"rollout {j}: distances"
    assert torch.equal(o2, want[last][4]) and torch.equal(o3, want[last][5])
    st = ses.stats()
    assert st["h2d_gbs"] > 0 and st["h2d_bytes_per_rollout"] > 0
    ses.close()


def test_session_rejects_bad_input(cuda):
    from cmr_agent_b200 import _lib, session
    ses = session.RolloutSession(2, 1024, 64, 8, 16, 2, synth.StepConfig(), depth=1, device=cuda)
    b = synth.make_batch(2, seed=1, num_pt=1024, img_h=32, img_w=64)
    a_r, a_t = synth.make_actions(2, 2)
    with pytest.raises(_lib.CmrError):
        ses.submit(dict(b, pc=b["pc"][:, :, :100].contiguous()), a_r, a_t)
    with pytest.raises(_lib.CmrError):
        ses.wait(5)
    ses.close()
