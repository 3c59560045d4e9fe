"""Host-side logic of the product that needs no GPU: step tables, sys.modules aliasing, episode
sharding, the scalar all-reduce payload (world_size-2 gloo)."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

from cmr_agent_b200 import dist as cdist
from cmr_agent_b200 import environment as drop_in
from cmr_agent_b200 import synth
from oracle import env_oracle as eo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_step_tables_equal_the_reference_expression():
    cfg = synth.StepConfig()
    rot, tt = drop_in.build_step_tables(cfg.r_steps, cfg.t_steps)
    assert rot.dtype == torch.float32 and tt.dtype == torch.float32
    for i in range(11):
        move = torch.zeros(1, 3)
        move[0, 1] = cfg.r_steps[i]          # f64 -> f32 on assignment, like environment.py:197
        want = eo.euler_angles_to_matrix(move, "XYZ")[0]
        # (Rx(0) @ Ry) @ Rz(0) equals Ry up to the sign of zeros
        assert torch.equal(rot[1, i] + 0.0, want + 0.0)
        assert float(tt[i]) == float(cfg.t_steps[i].float())
    with pytest.raises(ValueError):
        drop_in.euler_angles_to_matrix(torch.zeros(2, 3), "XX")
    with pytest.raises(ValueError):
        drop_in.euler_angles_to_matrix(torch.zeros(2, 2), "XYZ")
    ang = torch.rand(5, 3)
    assert torch.equal(drop_in.euler_angles_to_matrix(ang, "XYZ"), eo.euler_angles_to_matrix(ang, "XYZ"))


def _run(code):
    out = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, (out.stdout[-1000:], out.stderr[-3000:])


def test_install_without_a_reference_tree_registers_the_drop_ins():
    _run("""
        import sys
        sys.path.insert(0, %r)
        import cmr_agent_b200
        cmr_agent_b200.install()
        from environment import environment as env
        import importlib
        pn = importlib.import_module("models.pointnet_util")
        assert env.__name__ == "cmr_agent_b200.environment", env.__name__
        assert pn.__name__ == "cmr_agent_b200.pointnet_util"
        for f in ("init", "to_disentangled", "observation_from_a_pose", "expert", "step", "reward",
                  "euler_angles_to_matrix", "_axis_angle_rotation", "DEVICE"):
            assert hasattr(env, f), f
        for f in ("square_distance", "index_points", "farthest_point_sample", "query_ball_point",
                  "sample_and_group", "sample_and_group_all"):
            assert hasattr(pn, f), f
        cmr_agent_b200.uninstall()
        assert "environment.environment" not in sys.modules
        assert "models.pointnet_util" not in sys.modules
        print("ok")
    """ % ROOT)


def _needs_reference():
    from oracle import reference_loader as rl
    if not rl.available():
        pytest.skip("no reference tree (/root/reference or oracle/_ref)")
    return rl.REFERENCE_ROOT


def test_install_before_the_reference_imports_keeps_its_packages_whole():
    """INTEGRATION.md section 2 order: install() first, then the reference's own import lines
    (Train_Agent.py:13-16, models/PointNN.py:7).  environment.buffer must still resolve and the
    nn.Module classes of models/pointnet_util.py must still be there."""
    ref = _needs_reference()
    _run("""
        import sys
        sys.path.insert(0, %r)
        from oracle import reference_loader as rl
        rl.put_on_path()
        import cmr_agent_b200
        from cmr_agent_b200 import environment as drop_env, pointnet_util as drop_pn
        cmr_agent_b200.install()
        from config import KittiConfiguration
        from models import CMRAgent, MultiHeadModel
        from environment import environment as env
        from environment.buffer import Buffer
        import environment as env_pkg, models.PointNN as pnn, models.pointnet_util as pu
        assert env is drop_env and env_pkg.environment is drop_env
        assert env_pkg.__file__.startswith(%r)
        assert Buffer.__module__ == "environment.buffer"
        assert pu.__file__.startswith(%r)                      # the real module, patched
        assert pnn.index_points is drop_pn.index_points and pnn.square_distance is drop_pn.square_distance
        for f in ("farthest_point_sample", "query_ball_point", "sample_and_group", "sample_and_group_all"):
            assert getattr(pu, f) is getattr(drop_pn, f), f
        for c in ("PointNetSetAbstraction", "PointNetSetAbstractionMsg", "PointNetFeaturePropagation", "pc_normalize",
                  "timeit"):
            assert hasattr(pu, c), c
        # the classes reach the drop-ins through their module globals
        assert pu.PointNetSetAbstraction.forward.__globals__["sample_and_group"] is drop_pn.sample_and_group
        cmr_agent_b200.uninstall()
        assert "environment.environment" not in sys.modules or sys.modules["environment.environment"] is not drop_env
        print("ok")
    """ % (ROOT, ref, ref))


def test_accelerated_agent_keeps_the_reference_s_paths_where_the_kernels_do_not_apply():
    """accelerate_agent on a CPU agent: forward and action_from_logits are wrapped, and on CPU tensors, in training mode
    or with sampling they are the reference's own functions (models/CMRAgent.py:88-128) - same results, no library call."""
    _needs_reference()
    _run("""
        import sys
        sys.path.insert(0, %r)
        import torch
        from oracle import reference_loader as rl
        rl.put_on_path()
        from cmr_agent_b200 import agent_tower
        from config import KittiConfiguration
        from models import CMRAgent
        config = KittiConfiguration()
        torch.manual_seed(3)
        agent = CMRAgent(config).eval()
        reference_forward = agent.forward
        agent = agent_tower.accelerate_agent(agent)
        assert agent._cmr_b200_reference_action is CMRAgent.action_from_logits
        r = torch.randn(4, agent.degree_r, config.num_steps)
        t = torch.randn(4, agent.degree_t, config.num_steps)
        with torch.no_grad():
            got = agent.action_from_logits(r, t, deterministic=True)          # CPU logits: the reference's function
            want = CMRAgent.action_from_logits(r, t, deterministic=True)
            assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
            torch.manual_seed(9); a = agent.action_from_logits(r, t, deterministic=False)
            torch.manual_seed(9); b = CMRAgent.action_from_logits(r, t, deterministic=False)
            assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
            s2, s3 = torch.randn(1, 128, 40, 128), torch.randn(1, 5, 256)
            out, ref = agent(s2, s3), reference_forward(s2, s3)               # CPU state: the reference's forward
            assert all(torch.equal(x, y) for x, y in zip(out, ref))
        print("ok")
    """ % ROOT)


def test_install_after_the_reference_imports_repoints_bound_names():
    ref = _needs_reference()
    _run("""
        import sys
        sys.path.insert(0, %r)
        from oracle import reference_loader as rl
        rl.put_on_path()
        from models import CMRAgent
        from environment import environment as env
        from environment.buffer import Buffer
        import models.PointNN as pnn, models.pointnet_util as pu
        orig_sq, orig_env = pnn.square_distance, env
        import types
        driver = types.ModuleType("fake_driver"); driver.env = env; sys.modules["fake_driver"] = driver
        import cmr_agent_b200
        from cmr_agent_b200 import environment as drop_env, pointnet_util as drop_pn
        cmr_agent_b200.install()
        import environment as env_pkg
        assert env_pkg.environment is drop_env and sys.modules["environment.environment"] is drop_env
        assert driver.env is drop_env                                   # `from environment import environment as env`
        assert pnn.square_distance is drop_pn.square_distance and pnn.index_points is drop_pn.index_points
        assert pu.sample_and_group is drop_pn.sample_and_group
        cmr_agent_b200.uninstall()
        assert pnn.square_distance is orig_sq and driver.env is orig_env and pu.square_distance is orig_sq
        assert sys.modules["environment.environment"] is orig_env
        print("ok")
    """ % ROOT)


def test_signatures_match_the_reference_functions():
    import inspect
    from cmr_agent_b200 import pointnet_util as pn
    from oracle import reference_loader as rl
    if not rl.available():
        pytest.skip("/root/reference not present")
    ref_env, ref_pn = rl.environment(), rl.pointnet_util()

    def names(f):
        return list(inspect.signature(f).parameters)

    for f in ("init", "to_disentangled", "expert", "step", "reward", "euler_angles_to_matrix"):
        assert names(getattr(drop_in, f))[: len(names(getattr(ref_env, f)))] == names(getattr(ref_env, f)), f
    assert names(drop_in.observation_from_a_pose)[:2] == names(ref_env.observation_from_a_pose)
    for f in ("square_distance", "index_points", "farthest_point_sample", "query_ball_point", "sample_and_group",
              "sample_and_group_all"):
        want = names(getattr(ref_pn, f))
        got = inspect.signature(getattr(pn, f)).parameters
        assert list(got)[: len(want)] == want, f
        # anything beyond the reference's parameters is an optional extension (e.g. method="auto")
        assert all(p.default is not inspect.Parameter.empty for p in list(got.values())[len(want):]), f


@pytest.mark.parametrize("total,world", [(32, 1), (32, 8), (64, 8), (10, 4), (3, 8), (0, 2)])
def test_shard_range_partitions_episodes(total, world):
    seen = []
    for r in range(world):
        lo, hi = cdist.shard_range(total, r, world)
        assert 0 <= lo <= hi <= total
        seen += list(range(lo, hi))
        for e in range(lo, hi):
            assert cdist.owner_of(e, total, world) == r
    assert seen == list(range(total))
    sizes = [cdist.shard_range(total, r, world)[1] - cdist.shard_range(total, r, world)[0] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1


def test_metric_sums_single_process():
    m = cdist.MetricSums()
    m.add([1.0, 20.0, 3.0], [1.0, 1.0, 6.0], reward=[0.5, -0.5, 0.0])
    s = m.all_reduce().summary()
    assert s["episodes"] == 3 and abs(s["recall"] - 1 / 3) < 1e-12 and abs(s["rre_mean"] - 8.0) < 1e-12


_WORKER = """
import os, sys, json
sys.path.insert(0, %r)
import torch
from cmr_agent_b200 import dist as cdist, synth
rank, world, _ = cdist.init_from_env(backend="gloo")
total = 11
lo, hi = cdist.shard_range(total, rank, world)
# every rank scores only its own episodes; errors are a deterministic function of the episode id
eps = torch.arange(lo, hi, dtype=torch.float64)
m = cdist.MetricSums().add(eps * 2.0, eps * 0.5, reward=eps * 0.1).all_reduce()
slow = cdist.max_over_ranks(float(rank + 1), "cpu")
cdist.barrier()
if rank == 0:
    print(json.dumps({"summary": m.summary(), "slow": slow, "world": world}))
"""


def test_world_size_2_gloo_sharding_and_allreduce(tmp_path):
    import json
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    eps = torch.arange(11, dtype=torch.float64)
    ref = cdist.MetricSums().add(eps * 2.0, eps * 0.5, reward=eps * 0.1).summary()
    assert res["world"] == 2 and res["slow"] == 2.0
    for k, v in ref.items():
        assert abs(res["summary"][k] - v) < 1e-9, k


def test_reference_arm_of_the_bench_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU path the driver times beside ours) runs without a GPU and prints one JSON
    line with the contract's keys; a small sample keeps it to a few seconds."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-episodes", "2", "--iters", "2"], capture_output=True, text=True,
                         timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "steps/s" and line["value"] > 0
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    from oracle import reference_loader as rl
    # the real reference (staged as oracle/_ref or present as /root/reference) when available, else the oracle's port
    assert line["cpu_baseline"]["kind"] == ("reference" if rl.available() else "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["config"]["workload"] == "kitti_b32x10"


def test_reference_arm_runs_on_rank_0_only():
    """Under torchrun the driver starts one process per GPU: every rank but 0 leaves the reference arm at once,
    without output and with exit code 0."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29571")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=root, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip() == ""
