"""The reference's OWN callers, unchanged, on the drop-ins, on a B200 (BASELINE configs[0]).

`oracle/make_ref.py` stages the reference's python files into the git-ignored `oracle/_ref/`, which travels to
the GPU box with the snapshot.  After `cmr_agent_b200.install()` the reference's import lines resolve to the
drop-ins and the following run exactly as the reference wrote them:

  * `models.CMRAgent` (random init, seed fixed, eval) in the `Test_Agent.py:150-170` inference loop,
  * `environment.buffer.Buffer` + `agent.action_logprob_and_entropy` + the PPO/BC update of `Train_Agent.py:216-300`,
  * `models.PointNN.KnnPointTransformer` (`models/PointNN.py:188-232`) forward and backward,
  * `models.pointnet_util.PointNetSetAbstraction` (the module's own class, reaching the drop-ins through its globals),
  * the two driver scripts THEMSELVES - `Test_Agent.py` and `Train_Agent.py`, executed with runpy as `__main__` - with
    stand-ins only for what lies outside the path (the dataset on disk, the feature network and its checkpoints,
    tensorboardX) and, for training, one epoch instead of 64.

The comparator is the reference's `environment/environment.py` / `models/pointnet_util.py` themselves on CUDA tensors
(what a CMR-Agent user runs today), loaded under private module names.  On the GPU the reference is not bit-stable
against itself - cuBLAS rounds the k=3 products differently from the CPU, `pc.mean` differs by an ulp between the
per-sample and the batched call (environment.py:46 vs :91), the scatter uses float atomics - so the environment
comparison is made in lock-step (both environments see the drop-in's action sequence) with these bars:
obs3d rows 0-3 bit-exact; at most 4 in-frustum flags differ; >= 99.98 % of obs2d within 1e-5
relative; logits within 1e-4; the action equal unless the two best logits are a near-tie; poses within 1e-5.
The bit-exact bars stay where they are: against the CPU reference (tests/test_gpu_env.py, golden fixtures).
"""
import copy

import pytest
import torch

from cmr_agent_b200 import synth
from oracle import reference_loader as rl
from tests import helpers as hp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world(cuda):
    if not rl.available():
        pytest.skip("no reference tree (oracle/_ref is staged by oracle/make_ref.py in the build container)")
    rl.put_on_path()
    import cmr_agent_b200
    from cmr_agent_b200 import environment as drop_env, pointnet_util as drop_pn
    cmr_agent_b200.install()                         # before the reference's imports
    from config import KittiConfiguration            # noqa: E402  (reference packages, by their own names)
    from models import CMRAgent
    from environment import environment as env
    from environment.buffer import Buffer
    import models.PointNN as pnn
    import models.pointnet_util as pu
    assert env is drop_env and pnn.index_points is drop_pn.index_points and pu.sample_and_group is drop_pn.sample_and_group
    w = dict(cfg_cls=KittiConfiguration, CMRAgent=CMRAgent, env=env, Buffer=Buffer, pnn=pnn, pu=pu,
             ref_env=rl.environment(), ref_pn=rl.pointnet_util(), drop_pn=drop_pn, dev=cuda)
    yield w
    cmr_agent_b200.uninstall()


def _agent(w, config, seed=2023):
    torch.manual_seed(seed)
    return w["CMRAgent"](config).to(w["dev"]).eval()


def _batch(B, dev, first=0):
    data = hp.to_device(synth.make_batch(B, first_episode=first, seed=hp.SEED), dev)
    return data


def _inference_loop(env, agent, data, config, hook=None):
    """Test_Agent.py:150-170, verbatim control flow."""
    pose_source, pose_target = env.init(data)
    pose_target = env.to_disentangled(pose_target, data['pc'])
    for step in range(config.action_num):
        current_state_2d, current_state_3d = env.observation_from_a_pose(data, pose_source)
        action_r_logits, action_t_logits, state_value = agent(current_state_2d, current_state_3d)
        action_r, action_t = agent.action_from_logits(action_r_logits, action_t_logits, deterministic=True)
        if hook is not None:
            action_r, action_t = hook(step, pose_source, current_state_2d, current_state_3d, action_r_logits,
                                      action_t_logits, action_r, action_t)
        pose_source = env.step(action_r, action_t, pose_source, config)
    return pose_source, pose_target


@pytest.mark.parametrize("B", [1, 8])
def test_test_agent_loop_on_the_drop_in_matches_the_reference_environment_on_cuda(world, B):
    w = world
    config = w["cfg_cls"]()
    agent = _agent(w, config)
    data = _batch(B, w["dev"], first=3)
    N = data["pc"].shape[2]
    ours = []

    def record(step, pose, s2, s3, rl_, tl_, ar, at):
        ours.append(dict(pose=pose.clone(), s2=s2, s3=s3, rl=rl_.clone(), tl=tl_.clone(), ar=ar.clone(), at=at.clone()))
        return ar, at

    with torch.no_grad():
        pose_ours, target_ours = _inference_loop(w["env"], agent, data, config, record)
    assert len(ours) == config.action_num
    assert pose_ours.shape == (B, 4, 4) and torch.isfinite(pose_ours).all()
    # observations handed out earlier are still intact (fresh tensors per call; buffer.py:105-106 keeps them)
    for rec in ours:
        assert torch.equal(rec["s3"][:, :3], data["pc"])

    stats = dict(flag_flips=0, obs2d_bad=0.0, logit_err=0.0, near_ties=0, pose_err=0.0)

    def compare(step, pose, s2, s3, rl_, tl_, ar, at):
        rec = ours[step]
        stats["pose_err"] = max(stats["pose_err"], float((pose - rec["pose"]).abs().max()))
        assert torch.equal(s3[:, :4], rec["s3"][:, :4])
        flips = int((s3[:, 4] != rec["s3"][:, 4]).sum())
        stats["flag_flips"] = max(stats["flag_flips"], flips)
        a, b = s2.double(), rec["s2"].double()
        bad = ((a - b).abs() > 1e-5 * torch.maximum(a.abs(), b.abs()).clamp_min(1e-2)).double().mean()
        stats["obs2d_bad"] = max(stats["obs2d_bad"], float(bad))
        assert torch.equal(s2[:, :64], rec["s2"][:, :64])            # image half: a copy on both sides
        for mine, theirs, a_mine, a_theirs in ((rec["rl"], rl_, rec["ar"], ar), (rec["tl"], tl_, rec["at"], at)):
            stats["logit_err"] = max(stats["logit_err"], float((mine - theirs).abs().max()))
            top2 = theirs.topk(2, dim=-1).values
            gap = top2[..., 0] - top2[..., 1]
            differ = a_mine != a_theirs
            stats["near_ties"] += int(differ.sum())
            assert bool((gap[differ] < 1e-2).all()), "a different action although the logits were not a near-tie"
        return rec["ar"], rec["at"]          # lock-step: the reference environment follows the drop-in's actions

    with torch.no_grad():
        pose_ref, target_ref = _inference_loop(w["ref_env"], agent, data, config, compare)
    print(f"\n[g1 B={B}] {stats}")
    # measured on a B200 (round 2): 0 flips, 2.4e-5 of obs2d, logits 6e-7 apart, no near-tie
    assert stats["flag_flips"] <= 4, stats
    assert stats["obs2d_bad"] <= 2e-4, stats
    assert stats["logit_err"] <= 1e-4, stats
    assert stats["pose_err"] <= 1e-5, stats
    assert float((pose_ref - pose_ours).abs().max()) <= 1e-5
    assert float((target_ref - target_ours).abs().max()) <= 1e-4


@pytest.mark.parametrize("B", [1, 4])
def test_test_agent_loop_captured_as_one_graph_takes_the_eager_loop_s_actions(world, B):
    """environment.capture_rollout(policy=...): the Test_Agent.py:150-170 loop with the reference's unchanged agent as
    the on-device policy, captured once as a CUDA graph; every replay takes the eager loop's actions and ends at its
    pose."""
    from cmr_agent_b200 import agent_tower
    w = world
    config = w["cfg_cls"]()
    agent = agent_tower.accelerate_agent(_agent(w, config))
    data = _batch(B, w["dev"], first=11)
    taken = []

    def record(step, pose, s2, s3, rl_, tl_, ar, at):
        taken.append((ar.clone(), at.clone()))
        return ar, at

    torch.distributions.Distribution.set_default_validate_args(False)   # the check synchronises with the host
    try:
        with torch.no_grad():
            roll = w["env"].capture_rollout(
                data, config, with_reward=False, reusable=True,
                policy=lambda s2, s3: agent.action_from_logits(*agent(s2, s3)[:2], deterministic=True))
            # the capture serves every batch of the run: the eager loop on a batch, then the same batch loaded into
            # the captured tensors and replayed
            for first in (11, 40, 11):
                batch = _batch(B, w["dev"], first=first)
                del taken[:]
                pose_eager, _ = _inference_loop(w["env"], agent, batch, config, record)
                roll.load(batch).replay()
                torch.cuda.synchronize()
                assert roll.actions_r.shape[0] == config.action_num == len(taken)
                for it, (ar, at) in enumerate(taken):
                    assert torch.equal(roll.actions_r[it], ar) and torch.equal(roll.actions_t[it], at), (first, it)
                assert float((roll.pose - pose_eager).abs().max()) <= 1e-6
        with pytest.raises(ValueError):
            w["env"].capture_rollout(data, config)
        with pytest.raises(ValueError):
            w["env"].capture_rollout(data, config, policy=lambda a, b: None, reusable=True)
    finally:
        torch.distributions.Distribution.set_default_validate_args(True)


def test_train_agent_trajectory_buffer_and_update_run_unchanged(world):
    """Train_Agent.py:216-300 on the drop-in: expert, stochastic actions, reward, the reference's Buffer (which keeps
    every observation by reference), returns/advantages, one BC+PPO update through the unchanged CMRAgent."""
    w = world
    env, Buffer = w["env"], w["Buffer"]
    config = w["cfg_cls"]()
    agent = _agent(w, config, seed=7)
    B = 2
    data = _batch(B, w["dev"], first=11)
    # pc_in_cam_space / pc_mask / K / P stay on the CPU as the DataLoader delivers them
    buffer = Buffer(config)
    buffer.start_trajectory()
    kept = []
    torch.manual_seed(5)
    with torch.no_grad():
        pose_source, pose_target = env.init(data)
        pose_target = env.to_disentangled(pose_target, data['pc'])
        _, prev_p2p_distance = env.reward(pose_source, data)
        d0 = prev_p2p_distance.clone()
        for _ in range(config.action_num):
            expert_action_r, expert_action_t = env.expert(pose_source, pose_target, config, data)
            current_state_2d, current_state_3d = env.observation_from_a_pose(data, pose_source)
            action_r_logits, action_t_logits, state_value = agent(current_state_2d, current_state_3d)
            action_r, action_t = agent.action_from_logits(action_r_logits, action_t_logits, deterministic=False)
            action_logprob, action_entropy = agent.action_logprob_and_entropy(action_r_logits, action_t_logits, action_r,
                                                                              action_t)
            pose_source = env.step(action_r, action_t, pose_source, config)
            reward, prev_p2p_distance = env.reward(pose_source, data, prev_distance=prev_p2p_distance)
            buffer.log_step(current_state_2d, current_state_3d, state_value, reward, expert_action_r, expert_action_t,
                            action_r, action_t, action_logprob)
            kept.append((current_state_2d.clone(), current_state_3d.clone(), reward.clone()))
    # aliasing contract: what the buffer kept by reference was never overwritten by later calls
    for i, (s2, s3, r) in enumerate(kept):
        assert torch.equal(buffer.states_2d[-1][i], s2), f"obs2d of iteration {i} was overwritten"
        assert torch.equal(buffer.states_3d[-1][i], s3), f"obs3d of iteration {i} was overwritten"
        assert torch.equal(buffer.rewards[-1][i], r)
        assert r.shape == (B, 1, 1) and bool(((r == 0.5) | (r == -0.5) | (r == 0)).all())
    # the shipped reward ignores the pose (environment.py:272-275): the distance never changes, so every
    # step is "same" and the reward is better - worse - same = 0 (:293-298)
    assert torch.equal(prev_p2p_distance, d0)
    assert all(bool((r == 0).all()) for _, _, r in kept)
    # the reference's own reward on the same (CPU-resident) inputs gives the same distance within 1e-5
    _, d_ref = w["ref_env"].reward(pose_source, data)
    assert hp.rel_err(d0.cpu(), d_ref.cpu()) <= 1e-5

    agent.train()
    samples = buffer.get_samples()
    ppo_dataset = torch.utils.data.TensorDataset(*samples)
    loader = torch.utils.data.DataLoader(ppo_dataset, batch_size=10, shuffle=False, drop_last=False)
    optimizer = torch.optim.Adam(agent.parameters(), lr=1e-4)
    cross_entropy = torch.nn.CrossEntropyLoss()
    batch = next(iter(loader))
    states_2d, states_3d, state_values, expert_actions_r, expert_actions_t, action_r, action_t, action_logprob, \
        state_value_ref, advantages = batch
    new_r, new_t, new_value = agent(states_2d, states_3d)
    new_logprob, new_entropy = agent.action_logprob_and_entropy(new_r, new_t, action_r, action_t)
    loss_r = cross_entropy(new_r.view(-1, new_r.shape[2]), expert_actions_r.view(-1))
    loss_t = cross_entropy(new_t.view(-1, new_t.shape[2]), expert_actions_t.view(-1))
    ratio = torch.exp(new_logprob - action_logprob)
    policy_loss = -torch.min(ratio * advantages, ratio.clamp(1 - config.CLIP_EPS, 1 + config.CLIP_EPS) * advantages).mean()
    value_loss = (new_value.view(-1, 1) - state_value_ref).pow(2).mean()
    loss = loss_r + loss_t + (policy_loss + value_loss * config.W_VALUE - new_entropy.mean() * config.W_ENTROPY) * config.alpha
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    assert torch.isfinite(loss)


def test_knn_point_transformer_runs_unchanged_forward_and_backward(world):
    """models/PointNN.py:188-232 on the drop-in's square_distance / index_points against the same module on the
    reference's own functions (CUDA tensors)."""
    w = world
    pnn = w["pnn"]
    torch.manual_seed(11)
    layer = pnn.KnnPointTransformer(64, 64, k=16).to(w["dev"])
    B, n = 2, 1280
    xyz = (torch.rand(B, 3, n, device=w["dev"]) * 40 - 20).contiguous()
    feats = torch.randn(B, 64, n, device=w["dev"])

    def run():
        f = feats.clone().requires_grad_(True)
        layer.zero_grad()
        out = layer(xyz, f)
        out.square().mean().backward()
        return out.detach(), f.grad.detach(), [p.grad.detach().clone() for p in layer.parameters()]

    out_a, gin_a, gp_a = run()
    mine = (pnn.square_distance, pnn.index_points)
    try:
        pnn.square_distance, pnn.index_points = w["ref_pn"].square_distance, w["ref_pn"].index_points
        out_b, gin_b, gp_b = run()
    finally:
        pnn.square_distance, pnn.index_points = mine
    # the distances themselves are bit-identical to the reference expression on this device
    pts = xyz.permute(0, 2, 1)
    assert torch.equal(w["drop_pn"].square_distance(pts, pts), w["ref_pn"].square_distance(pts, pts))
    assert hp.rel_err(out_a.cpu(), out_b.cpu(), floor=1e-6) <= 1e-4
    assert torch.allclose(gin_a, gin_b, rtol=1e-4, atol=1e-7)
    for a, b in zip(gp_a, gp_b):
        assert torch.allclose(a, b, rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("knn", [False, True])
def test_set_abstraction_class_of_the_reference_module_reaches_the_drop_ins(world, knn):
    """models/pointnet_util.py:156-199: the nn.Module the reference ships, kept in place by install() and calling
    sample_and_group through its module globals; compared with the untouched module (private copy)."""
    w = world
    pu, ref_pn = w["pu"], w["ref_pn"]
    torch.manual_seed(3)
    sa = pu.PointNetSetAbstraction(npoint=64, radius=6.0, nsample=16, in_channel=3 + 8, mlp=[16, 32], group_all=False,
                                   knn=knn).to(w["dev"])
    sa_ref = ref_pn.PointNetSetAbstraction(npoint=64, radius=6.0, nsample=16, in_channel=3 + 8, mlp=[16, 32],
                                           group_all=False, knn=knn).to(w["dev"])
    sa_ref.load_state_dict(copy.deepcopy(sa.state_dict()))
    B, N = 3, 2048
    xyz = (torch.rand(B, N, 3, device=w["dev"]) * 30).contiguous()
    pts = torch.randn(B, N, 8, device=w["dev"])

    def run(mod):
        torch.manual_seed(99)                        # farthest_point_sample draws its start on the CPU generator
        p = pts.clone().requires_grad_(True)
        new_xyz, new_points = mod(xyz, p)
        new_points.square().mean().backward()
        return new_xyz.detach(), new_points.detach(), p.grad.detach()

    a = run(sa)
    b = run(sa_ref)
    assert torch.equal(a[0], b[0])                   # same centroids: FPS indices are bit-exact
    assert torch.allclose(a[1], b[1], rtol=1e-4, atol=1e-6)
    assert torch.allclose(a[2], b[2], rtol=1e-3, atol=1e-7)


# ---- the driver script itself ----------------------------------------------------------------------------------------
class _SynthKitti(torch.utils.data.Dataset):
    """Stand-in for dataset/KittiDataset.py (files on disk): what its loader yields per sample, from the synthetic
    generator.  The feature network's outputs ride along under private names for `_FeatureNetStandIn`."""
    first, count = 200, 3
    counts = {"test": 3, "train": 32, "val": 8}

    def __init__(self, config, mode="test"):
        self.config = config
        self.count = self.counts[mode]
        self.first = {"test": 200, "train": 300, "val": 400}[mode]

    def __len__(self):
        return self.count

    def __getitem__(self, i):
        b = synth.make_batch(1, first_episode=self.first + i, seed=hp.SEED)
        item = {k: v[0].clone() for k, v in b.items() if k != "img"}
        item["img"] = torch.zeros(3, b["img"].shape[2], b["img"].shape[3])
        for k in ("pc_overlap_pred", "pc_geo_feat", "img_geo_feat"):
            item["_net_" + k] = item.pop(k)
        return item


class _FeatureNetStandIn(torch.nn.Module):
    """Stand-in for models.MultiHeadModel (ViT + point transformer + checkpoint, outside the path): leaves on the
    device what the real network leaves there."""

    def __init__(self, config):
        super().__init__()

    def load_state_dict(self, state_dict, strict=True):
        return None

    def forward(self, data):
        data["pc"] = data["pc"].cuda()
        for k in ("pc_overlap_pred", "pc_geo_feat", "img_geo_feat"):
            data[k] = data.pop("_net_" + k).cuda()
        return data


def test_the_reference_s_test_agent_script_runs_unchanged_on_the_drop_in(world, tmp_path, monkeypatch, capsys):
    """`python Test_Agent.py --dataset kitti` (the file itself, through runpy, not a restatement of its loop) with
    `environment.environment` = the drop-in: its imports, its loop (:150-170), its error metrics (:98-105,181-206).
    Outside the path and replaced by stand-ins: the dataset on disk, the feature network, the two checkpoint files.
    The errors it prints equal the same loop run here, sample by sample."""
    import os
    import runpy
    import sys
    import numpy as np
    import dataset as ref_dataset
    import models as ref_models
    w = world
    config = w["cfg_cls"]()
    torch.manual_seed(99)
    weights = w["CMRAgent"](config).state_dict()
    monkeypatch.setattr(ref_dataset, "KittiDataset", _SynthKitti)
    monkeypatch.setattr(ref_models, "MultiHeadModel", _FeatureNetStandIn)
    monkeypatch.setattr(torch, "load", lambda path, *a, **k: weights if "agent" in str(path) else {})
    monkeypatch.setattr(sys, "argv", ["Test_Agent.py", "--dataset", "kitti"])
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", os.environ.get("CUDA_VISIBLE_DEVICES", "0"))
    monkeypatch.chdir(tmp_path)
    script = os.path.join(rl.put_on_path(), "Test_Agent.py")
    with np.errstate(all="ignore"):
        ns = runpy.run_path(script, run_name="__main__")
    out = capsys.readouterr().out
    assert ns["env"] is w["env"], "the script's `env` is not the drop-in"
    assert "Registration Recall:" in out and "RRE Std:" in out
    printed = [tuple(float(x) for x in ln.split()) for ln in out.splitlines()
               if len(ln.split()) == 2 and ln.split()[0][0].isdigit()]
    assert len(printed) == _SynthKitti.count, out
    # the same samples through the loop as restated in this file, with the script's own metric
    agent = w["CMRAgent"](config)
    agent.load_state_dict(weights)
    agent = agent.to(w["dev"]).eval()
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = False, False   # the script's set_seed set them
    with torch.no_grad():
        for i, (t_script, r_script) in enumerate(printed):
            data = hp.to_device(synth.make_batch(1, first_episode=_SynthKitti.first + i, seed=hp.SEED), w["dev"])
            pose, target = _inference_loop(w["env"], agent, data, config)
            t_here, r_here = ns["get_P_diff"](pose[0].cpu().numpy(), target[0].cpu().numpy())
            assert abs(t_here - t_script) <= 1e-3 * max(1.0, abs(t_here)), (i, t_here, t_script)
            assert abs(r_here - r_script) <= 1e-3 * max(1.0, abs(r_here)), (i, r_here, r_script)


def test_the_reference_s_train_agent_script_runs_unchanged_on_the_drop_in(world, tmp_path, monkeypatch, capsys):
    """`python Train_Agent.py --dataset kitti` (the file itself, through runpy) on the drop-in: its validation pass
    (:160-213), the trajectory loop with expert / stochastic actions / reward / Buffer (:216-248), the BC + PPO update
    (:252-312), checkpointing and logging.  Stand-ins for what is outside the path: the dataset on disk, the feature
    network and its checkpoint, tensorboardX; the configuration is the reference's class with ONE epoch (its 64 would
    run for days) - 4 global steps of 8 episodes = one full buffer = one update."""
    import os
    import runpy
    import sys
    import time
    import numpy as np
    import config as ref_config
    import dataset as ref_dataset
    import models as ref_models
    import tensorboardX
    w = world
    scalars = []

    class OneEpoch(w["cfg_cls"]):
        def __init__(self):
            super().__init__()
            self.epoch = 1
            self.num_workers = 4

    class Writer:
        def __init__(self, *a, **k):
            pass

        def add_scalar(self, tag, value, global_step=None):
            scalars.append((tag, float(value), global_step))

    monkeypatch.setattr(ref_config, "KittiConfiguration", OneEpoch)
    monkeypatch.setattr(ref_dataset, "KittiDataset", _SynthKitti)
    monkeypatch.setattr(ref_models, "MultiHeadModel", _FeatureNetStandIn)
    monkeypatch.setattr(tensorboardX, "SummaryWriter", Writer)
    monkeypatch.setattr(torch, "load", lambda path, *a, **k: {})
    monkeypatch.setattr(time, "sleep", lambda s: None)
    if not hasattr(np, "Inf"):
        monkeypatch.setattr(np, "Inf", np.inf, raising=False)      # Train_Agent.py:157-158 predates numpy 2
    monkeypatch.setattr(sys, "argv", ["Train_Agent.py", "--dataset", "kitti"])
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", os.environ.get("CUDA_VISIBLE_DEVICES", "0"))
    monkeypatch.chdir(tmp_path)
    script = os.path.join(rl.put_on_path(), "Train_Agent.py")
    before = {k: v.clone() for k, v in w["CMRAgent"](OneEpoch()).state_dict().items()}   # shapes only
    try:
        ns = runpy.run_path(script, run_name="__main__")
    finally:
        torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = False, False
    out = capsys.readouterr().out
    assert ns["env"] is w["env"], "the script's `env` is not the drop-in"
    assert ns["Buffer"] is w["Buffer"]
    assert "New Training!" in out and "0-th epoch end." in out
    assert ns["global_step"] == _SynthKitti.counts["train"] // ns["config"].train_batch_size == 4
    tags = [t for t, _, _ in scalars]
    assert tags.count("val_error/error_r") == 1 and tags.count("train_loss/BC_Loss") == 1, tags
    for tag, value, _ in scalars:
        assert np.isfinite(value), (tag, value)
    saved = [f for _, _, fs in os.walk(tmp_path / "checkpoint") for f in fs if f.endswith(".pth")]
    assert len(saved) == 1, saved
    after = ns["agent"].state_dict()
    assert set(after) == set(before)
    moved = sum(float((after[k].float().cpu() - before[k].float()).abs().sum()) for k in before)
    assert moved > 0 and all(torch.isfinite(v.float()).all() for v in after.values())
