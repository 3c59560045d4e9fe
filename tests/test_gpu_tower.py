"""GPU parity of the agent's 3-D tower: cmr_tower_forward (tcgen05 + TMEM,
cmr_agent_b200/csrc/tower_kernels.cuh) through its Python binding against

  * tests/golden/tower.npz - produced by the reference's own ConvBNReLURes1D modules composed as CMRAgent.forward
    composes them (tests/golden/make_golden.py tower),
  * oracle/tower_oracle.py on seeded weights at KITTI size, ragged sizes and many short episodes (CTAs that cross
    episode boundaries),
  * the reference's CMRAgent itself on the GPU (eval), with `accelerate_agent` routing its 3-D half through the kernel.

Tolerance: 1e-5 of the output's scale (max |embed_3d| of the episode) - tests/test_tower_oracle.py explains why an
elementwise bound cannot be met even by the reference's own float32 evaluation."""
import os
import sys

import numpy as np
import pytest
import torch

from cmr_agent_b200 import synth
from oracle import reference_loader as rl, tower_oracle as to

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _scaled_err(got, want):
    return float(((got - want).abs().max(dim=1)[0] / want.abs().max(dim=1)[0]).max())


def _obs3d(B, N, seed):
    g = torch.Generator().manual_seed(seed)
    xyz = (torch.rand(B, 3, N, generator=g) - 0.5) * 160.0           # metre-scale coordinates, like the clouds
    flags = (torch.rand(B, 2, N, generator=g) < 0.3).float()
    return torch.cat([xyz, flags], dim=1).contiguous()


def _states(seed):
    return [to.make_state(seed + i, cin, cout) for i, (cin, cout) in enumerate(to.TOWER)]


def test_tower_matches_reference_golden(cuda):
    from cmr_agent_b200 import agent_tower
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tower.npz"))
    states, obs3d = make_golden.tower_inputs()
    got = agent_tower.Tower3D(states, cuda)(obs3d.to(cuda)).cpu()
    assert _scaled_err(got, torch.from_numpy(g["embed_3d"])) <= TOL


@pytest.mark.parametrize("B,N,seed", [(2, 40960, 11), (3, 1531, 12), (1, 100, 13), (5, 129, 14), (1, 128, 15),
                                      (300, 200, 16), (37, 4096, 17)])
def test_tower_matches_oracle(cuda, B, N, seed):
    from cmr_agent_b200 import agent_tower
    states, obs3d = _states(100 + seed), _obs3d(B, N, seed)
    want = to.tower(states, obs3d)
    tower = agent_tower.Tower3D(states, cuda)
    got = tower(obs3d.to(cuda)).cpu()
    assert torch.isfinite(got).all()
    assert _scaled_err(got, want) <= TOL
    # a second call on the cached workspace (keys are re-zeroed per call) gives the same bits
    assert torch.equal(tower(obs3d.to(cuda)).cpu(), got)


def test_tower_on_environment_observations(cuda):
    """obs3d exactly as cmr_observe produces it (xyz, predicted-overlap flag, in-frustum flag)."""
    from cmr_agent_b200 import agent_tower, environment as env
    from tests import helpers as hp
    data = hp.to_device(synth.make_batch(2, first_episode=5, seed=hp.SEED), cuda)
    pose, _ = env.init(data)
    _, obs3d = env.observation_from_a_pose(data, pose)
    states = _states(777)
    got = agent_tower.Tower3D(states, cuda)(obs3d).cpu()
    assert _scaled_err(got, to.tower(states, obs3d.cpu())) <= TOL


def test_accelerated_agent_matches_the_reference_agent(cuda):
    """The reference's CMRAgent (random init, eval) with its 3-D half on the kernel against its own forward on the
    GPU.  The reference's cuDNN 1x1 convolutions run in TF32 by default (torch.backends.cudnn.allow_tf32), i.e. they
    are the LESS accurate side; with TF32 off the two agree to 1e-5 of the logits' scale."""
    if not rl.available():
        pytest.skip("no reference tree (oracle/_ref is staged by oracle/make_ref.py in the build container)")
    rl.put_on_path()
    from config import KittiConfiguration
    from models import CMRAgent
    from cmr_agent_b200 import agent_tower
    config = KittiConfiguration()
    torch.manual_seed(4)
    agent = CMRAgent(config).to(cuda).eval()
    # give the BatchNorms non-trivial running statistics
    with torch.no_grad():
        for m in list(agent.state_3d_embed.modules()) + list(agent.state_2d_embed.modules()):
            if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                m.running_mean.normal_(0, 0.2)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.normal_(1.0, 0.2)
                m.bias.normal_(0, 0.1)
    B, N = 2, 40960
    s2 = torch.randn(B, 128, 40, 128, device=cuda)
    s3 = _obs3d(B, N, 21).to(cuda)
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            want = agent(s2, s3)
            agent_tower.accelerate_agent(agent)
            got = agent(s2, s3)
        # training mode / autograd: the reference's own forward, untouched
        agent.train()
        out_train = agent(s2[:1], s3[:1, :, :2048])
        assert out_train[0].requires_grad
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    for a, b in zip(got, want):
        assert float((a - b).abs().max() / b.abs().max()) <= 1e-5


def test_tower_reports_activations_beyond_the_fp16_range(cuda):
    """The split pieces are fp16: an activation beyond 65504 must come out as NaN/inf AND raise the sticky fault word
    (cmr_take_fault() == 3) - never a silently wrong finite number."""
    from cmr_agent_b200 import _lib, agent_tower
    states = _states(321)
    obs3d = _obs3d(2, 2048, 5)
    _lib.take_fault()
    ok = agent_tower.Tower3D(states, cuda)(obs3d.to(cuda))
    torch.cuda.synchronize()
    assert torch.isfinite(ok).all() and _lib.take_fault() == 0
    obs3d[1, :3] *= 1e5                                  # "coordinates" of 10^7 metres in the second episode
    bad = agent_tower.Tower3D(states, cuda)(obs3d.to(cuda))
    torch.cuda.synchronize()
    assert torch.isfinite(bad[0]).all() and torch.equal(bad[0], ok[0])     # episodes are independent
    assert not torch.isfinite(bad[1]).all()
    assert _lib.take_fault() == 3


@pytest.mark.parametrize("dof6", [False, True])
def test_heads_match_the_reference_s_modules(cuda, dof6):
    """agent_tower.Heads (cmr_grouped_linear: layer l of policy_r, policy_t and value in ONE launch) against the
    reference's own nn.Sequential heads (models/CMRAgent.py:70-86) in fp32: 1e-5 of the output's scale; a row's result
    does not depend on the batch it sits in (bit-identical at B = 1 and inside B = 33)."""
    if not rl.available():
        pytest.skip("no reference tree (oracle/_ref)")
    rl.put_on_path()
    from config import KittiConfiguration
    from models import CMRAgent
    from cmr_agent_b200 import agent_tower
    config = KittiConfiguration()
    config.is_6_DoF = dof6
    torch.manual_seed(11)
    agent = CMRAgent(config).to(cuda).eval()
    heads = agent_tower.Heads([agent.policy_r, agent.policy_t, agent.value])
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            for B in (1, 5, 33):
                emb = torch.randn(B, 256, device=cuda) * 3.0
                got = heads(emb)
                want = (agent.policy_r(emb), agent.policy_t(emb), agent.value(emb))
                for g, w_ in zip(got, want):
                    assert g.shape == w_.shape
                    assert float((g - w_).abs().max()) <= 1e-5 * float(w_.abs().max().clamp_min(1e-3)), (B, g.shape)
                one = heads(emb[B - 1:].contiguous())
                for g, o in zip(got, one):
                    assert torch.equal(g[B - 1:], o)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    with pytest.raises(Exception):
        agent_tower.Heads([torch.nn.Sequential(torch.nn.Linear(300, 4).to(cuda))])


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("shape,pool", [((2, 8, 6, 16), 0), ((3, 12, 4, 8), 1), ((2, 8, 5, 16), 2), ((1, 128, 40, 128), 1),
                                        ((2, 256, 5, 16), 2), ((1, 132, 5, 7), 2), ((3, 128, 9, 15), 2)])
def test_conv_epilogue_matches_torch(cuda, shape, pool, channels_last):
    """cmr_conv_epilogue = pool(LeakyReLU(x * scale[c] + shift[c])) against torch's own ops on the folded form, on NCHW
    and on channels_last data."""
    from cmr_agent_b200 import _lib
    B, C, H, W = shape
    fmt = torch.channels_last if channels_last else torch.contiguous_format
    g = torch.Generator().manual_seed(B * 100 + C)
    x = torch.randn(shape, generator=g).to(cuda).contiguous(memory_format=fmt)
    scale = (torch.rand(C, generator=g) + 0.5).to(cuda)
    shift = torch.randn(C, generator=g).to(cuda)
    want = torch.nn.functional.leaky_relu(x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1), 0.01)
    if pool == 1:
        want = torch.nn.functional.avg_pool2d(want, 2, 2)
    elif pool == 2:
        want = torch.nn.functional.avg_pool2d(want, (H, W), 1)
    got = torch.empty(want.shape, device=cuda, memory_format=fmt)
    _lib.call("cmr_conv_epilogue", _lib.ptr(x), _lib.ptr(scale), _lib.ptr(shift), 0.01, pool, 1 if channels_last else 0,
              B, C, H, W, _lib.ptr(got), _lib.stream())
    torch.cuda.synchronize()
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 2e-6 * float(want.abs().max())


@pytest.mark.parametrize("shape", [(1, 128, 40, 128), (3, 128, 40, 80), (2, 36, 6, 10), (2, 8, 5, 7), (1, 3, 4, 4)])
def test_to_channels_last_is_torchs_copy(cuda, shape):
    """cmr_to_channels_last = x.contiguous(memory_format=torch.channels_last) bit for bit (the observation handed to the
    2-D head, models/CMRAgent.py:89), on the reference's shapes, on partial tiles and on shapes that take the scalar kernel."""
    from cmr_agent_b200 import _lib
    B, C, H, W = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g).to(cuda)
    want = x.contiguous(memory_format=torch.channels_last)
    got = torch.full(shape, float("nan"), device=cuda).contiguous(memory_format=torch.channels_last)
    _lib.call("cmr_to_channels_last", _lib.ptr(x), B, C, H, W, _lib.ptr(got), _lib.stream())
    torch.cuda.synchronize()
    assert got.is_contiguous(memory_format=torch.channels_last) or C == 1
    assert torch.equal(got, want)
    assert torch.equal(got.permute(0, 2, 3, 1).contiguous().view(-1), want.permute(0, 2, 3, 1).contiguous().view(-1))


def _action_rows(S, seed):
    """Rows of logits that stress the tie-breaking: plain random rows at several scales, exact duplicates of the
    maximum, and maxima one or two ulps apart (their probabilities usually round to the same float)."""
    g = torch.Generator().manual_seed(seed)
    rows = [torch.randn(6000, S, generator=g) * s for s in (1.0, 0.01, 10.0, 1e-4, 50.0)]
    base = torch.randn(6000, S, generator=g)
    top = base.max(dim=1, keepdim=True).values
    j = torch.randint(0, S, (6000, 1), generator=g)
    dup = base.clone().scatter_(1, j, top)                                    # a second copy of the maximum
    up = base.clone().scatter_(1, j, torch.nextafter(top, top + 1))           # one ulp above the old maximum
    up2 = torch.nextafter(up, up + 1).where(torch.zeros_like(up, dtype=torch.bool).scatter_(1, j, True), up)
    down = base.clone().scatter_(1, j, torch.nextafter(top, top - 1))         # one ulp below
    return torch.cat(rows + [dup, up, up2, down, base * 0.0, base * 0.0 + 3.5])


@pytest.mark.parametrize("S", [11, 9, 16, 13])
def test_deterministic_action_is_torchs_argmax_of_probs(cuda, S):
    """cmr_deterministic_action against models/CMRAgent.py:118-124 on the same GPU: the probabilities of
    Categorical(logits=x) bit for bit (every operation and the order of both sums are torch's), hence the same argmax -
    including rows whose two largest probabilities round to the same float."""
    from cmr_agent_b200 import _lib
    x = _action_rows(S, 100 + S)
    B = x.shape[0] // 6 * 2
    r = x[:B * 3].reshape(B, 3, S).to(cuda)
    wide = torch.zeros(B, 3 * S + 5, device=cuda)                             # the translation logits: a column block
    wide[:, 2:2 + 3 * S] = x[-B * 3:].reshape(B, 3 * S).to(cuda)
    t = wide[:, 2:2 + 3 * S].reshape(B, 3, S)
    assert t.stride() == (3 * S + 5, S, 1)
    a_r = torch.empty(B, 3, device=cuda, dtype=torch.int64)
    a_t = torch.empty(B, 3, device=cuda, dtype=torch.int64)
    p_r = torch.empty(B, 3, S, device=cuda)
    p_t = torch.empty(B, 3, S, device=cuda)
    _lib.call("cmr_deterministic_action", _lib.ptr(r), 3, r.stride(0), _lib.ptr(t), 3, t.stride(0), B, S, _lib.ptr(a_r),
              _lib.ptr(a_t), _lib.ptr(p_r), _lib.ptr(p_t), _lib.stream())
    torch.cuda.synchronize()
    for logits, probs, act in ((r, p_r, a_r), (t, p_t, a_t)):
        want = torch.distributions.Categorical(logits=logits).probs
        assert torch.equal(probs, want), f"{int((probs != want).sum())} of {want.numel()} probabilities differ"
        assert torch.equal(act, torch.argmax(want, dim=-1))
    ties = (p_r == p_r.max(dim=-1, keepdim=True).values).sum(-1) > 1
    assert int(ties.sum()) > 100                                               # the tie-breaking was exercised


def test_accelerated_agent_takes_the_reference_s_deterministic_actions(cuda):
    """accelerate_agent's action_from_logits: the reference's static method (models/CMRAgent.py:117-128) on the same
    logits gives the same actions; sampling (deterministic=False) and autograd stay the reference's."""
    if not rl.available():
        pytest.skip("no reference tree (oracle/_ref is staged by oracle/make_ref.py in the build container)")
    from cmr_agent_b200 import agent_tower
    rl.put_on_path()
    from config import KittiConfiguration
    from models import CMRAgent
    config = KittiConfiguration()
    torch.manual_seed(5)
    agent = agent_tower.accelerate_agent(CMRAgent(config).to(cuda).eval())
    r = torch.randn(8, agent.degree_r, config.num_steps, device=cuda)
    t = torch.randn(8, agent.degree_t, config.num_steps, device=cuda)
    with torch.no_grad():
        n0 = _lib_launches()
        got = agent.action_from_logits(r, t, deterministic=True)
        assert _lib_launches() == n0 + 1
        want = CMRAgent.action_from_logits(r, t, deterministic=True)
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
        assert got[0].dtype == want[0].dtype and got[0].shape == want[0].shape
        n0 = _lib_launches()
        agent.action_from_logits(r, t, deterministic=False)
        assert _lib_launches() == n0
    n0 = _lib_launches()
    agent.action_from_logits(r, t, deterministic=True)                        # autograd on: the reference's function
    assert _lib_launches() == n0


def _lib_launches():
    from cmr_agent_b200 import _lib
    return _lib.launch_count()
