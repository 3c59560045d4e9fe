"""Generate the golden fixtures in this directory FROM THE REAL REFERENCE.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
It imports /root/reference/environment/environment.py and models/pointnet_util.py by file path
(with the torch_scatter/open3d shims of oracle/shims.py), feeds them seeded synthetic inputs from
cmr_agent_b200.synth, and stores the OUTPUTS (plus the tiny inputs that cannot be regenerated:
poses, cloud means, FPS start indices).  Bulk inputs are regenerated from the seed at test time and
verified against the sha256 stored here.  The fixtures travel to the GPU box; the reference does not.

    python tests/golden/make_golden.py dataset | cost_volume | tower     # one fixture only
dataset.npz: the reference's FarthestSampler class + scipy's cKDTree.  cost_volume.npz: the reference's own statements of
IterModel.py:96-172 and :272-351, compiled from its syntax tree (the class cannot be constructed without a GPU - __init__ calls
.cuda(), :29 - and forward() runs the whole model around them).  tower.npz: the reference's ConvBNReLURes1D modules composed as CMRAgent.forward composes them.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from cmr_agent_b200 import synth  # noqa: E402
from oracle import reference_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

ENV_CASES = {
    # name: (batch, shape kwargs, iterations, store_full_obs2d)
    "env_small": (2, dict(num_pt=4096, img_h=64, img_w=256), 4, True),
    "env_ragged": (2, dict(num_pt=1531, img_h=36, img_w=100), 3, True),     # N % 4 != 0, P % 128 != 0
    "env_kitti": (2, dict(num_pt=40960, img_h=160, img_w=512), 4, False),
    "env_nuscenes": (1, dict(num_pt=40960, img_h=160, img_w=320, unique=(26000, 34000)), 3, False),
}
SEED = 2023


def sha(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(t.numpy()).tobytes())
    return h.hexdigest()


def env_case(name, batch, shape, iters, full):
    env = reference_loader.environment()
    data = synth.make_batch(batch, seed=SEED, **shape)
    cfg = synth.StepConfig()
    a_r, a_t = synth.make_actions(batch, iters, seed=SEED)
    H, W = shape["img_h"] // 4, shape["img_w"] // 4
    pose, target = env.init(data)
    target = env.to_disentangled(target.clone(), data["pc"])
    out = {
        "inputs_sha": sha(data["pc"], data["pc_geo_feat"], data["img_geo_feat"], data["pc_overlap_pred"],
                          data["pc_mask"], data["pc_in_cam_space"], data["K"]),
        "mean": data["pc"].mean(dim=2).numpy(),
        "pose_target_disentangled": target.numpy(),
        "a_r": a_r.numpy(), "a_t": a_t.numpy(),
    }
    # start from a non-trivial pose so that points are visible: apply two expert-free random steps first
    prev = None
    for it in range(iters):
        o2, o3 = env.observation_from_a_pose(data, pose)
        out[f"pose_{it}"] = pose.clone().numpy()
        # integer by-products recomputed with the reference's own lines (54-72) for ALL points
        mean = data["pc"].mean(dim=2, keepdim=True)
        X = pose[:, :3, :3] @ (data["pc"] - mean) + mean + pose[:, :3, 3:4]
        U = data["K"] @ X
        U[:, 0:2] = U[:, 0:2] / U[:, 2:3]
        inc = (U[:, 0] >= 0) & (U[:, 0] <= W - 1) & (U[:, 1] >= 0) & (U[:, 1] <= H - 1) & (U[:, 2] > 0)
        uv = U[:, 0:2].round().int()
        idx = uv[:, 1] * W + uv[:, 0]
        idx[~inc] = H * W
        assert torch.equal(inc.float(), o3[:, 4]), "in-frustum by-product disagrees with observation_3d"
        out[f"idx_{it}"] = idx.numpy().astype(np.int32)
        out[f"incam_{it}"] = np.packbits(inc.numpy(), axis=1)
        if full:
            out[f"obs2d_proj_{it}"] = o2[:, 64:].numpy()
        else:
            out[f"obs2d_proj_chansum_{it}"] = o2[:, 64:].double().sum(dim=1).numpy()
            out[f"obs2d_proj_absmax_{it}"] = o2[:, 64:].abs().amax(dim=1).numpy()
        assert torch.equal(o2[:, :64], data["img_geo_feat"])
        assert torch.equal(o3[:, :3], data["pc"]) and torch.equal(o3[:, 3], data["pc_overlap_pred"].float())
        env.step(a_r[it], a_t[it], pose, cfg)
        rew, dist = env.reward(pose, data, prev)
        out[f"reward_{it}"] = rew.numpy()
        out[f"dist_{it}"] = dist.numpy()
        prev = dist
    out["pose_final"] = pose.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: getattr(v, "shape", v) for k, v in out.items() if k.startswith(("idx_0", "obs2d", "mean"))})


def step_case():
    """Every (bin) of the 3-DoF tables and a batch of random 6-DoF actions through the reference's step()."""
    env = reference_loader.environment()
    out = {}
    for dof6 in (False, True):
        cfg = synth.StepConfig(is_6_DoF=dof6)
        B = 121 if not dof6 else 256
        g = torch.Generator().manual_seed(SEED + 5)
        pose = torch.eye(4).repeat(B, 1, 1)
        # random but valid starting rotations/translations
        ang = torch.rand(B, 3, generator=g) * 6.0 - 3.0
        pose[:, :3, :3] = env.euler_angles_to_matrix(ang, "XYZ")
        pose[:, :3, 3] = torch.randn(B, 3, generator=g) * 5
        if dof6:
            a_r = torch.randint(0, 11, (B, 3), generator=g)
            a_t = torch.randint(0, 11, (B, 3), generator=g)
        else:
            grid = torch.arange(121)
            a_r = (grid % 11).view(B, 1)
            a_t = torch.stack([grid // 11, (grid * 7) % 11], 1)
        tag = "6" if dof6 else "3"
        out[f"pose_in_{tag}"] = pose.clone().numpy()
        out[f"a_r_{tag}"] = a_r.numpy()
        out[f"a_t_{tag}"] = a_t.numpy()
        out[f"pose_out_{tag}"] = env.step(a_r, a_t, pose, cfg).numpy()
    np.savez_compressed(os.path.join(HERE, "step.npz"), **out)
    print("step", {k: v.shape for k, v in out.items()})


def pointnet_case():
    pn = reference_loader.pointnet_util()
    out = {}
    # duplicate-padded cloud (exact distance ties) and a plain one
    for tag, unique in (("dup", (2600, 2600)), ("plain", None)):
        xyz = synth.make_cloud_batch(2, num_pt=4096, seed=SEED, unique=unique)
        out[f"xyz_sha_{tag}"] = sha(xyz)
        torch.manual_seed(SEED)
        fps = pn.farthest_point_sample(xyz, 128)
        out[f"fps_{tag}"] = fps.numpy()
        new_xyz = pn.index_points(xyz, fps)
        out[f"new_xyz_{tag}"] = new_xyz.numpy()
        d = pn.square_distance(new_xyz, xyz)
        out[f"sqdist_rowsum_{tag}"] = d.double().sum(-1).numpy()
        out[f"knn16_raw_{tag}"] = d.argsort()[:, :, :16].numpy()                 # unstable order (A.7)
        out[f"knn16_stable_{tag}"] = d.argsort(stable=True)[:, :, :16].numpy()
        for r in (0.5, 2.0):
            out[f"ball_{r}_{tag}"] = pn.query_ball_point(r, 32, xyz, new_xyz).numpy()
        torch.manual_seed(SEED + 1)
        nx, npts, gxyz, fidx = pn.sample_and_group(64, 1.5, 16, xyz, xyz * 0.5 + 1.0, returnfps=True)
        out[f"sag_new_xyz_{tag}"] = nx.numpy()
        out[f"sag_new_points_{tag}"] = npts.numpy()
        out[f"sag_fps_{tag}"] = fidx.numpy()
    # full-size FPS (config 4 shape, one cloud): 40960 -> 1280
    xyz = synth.make_cloud_batch(1, num_pt=40960, seed=SEED + 100)
    out["xyz_sha_full"] = sha(xyz)
    torch.manual_seed(SEED + 2)
    out["fps_full"] = pn.farthest_point_sample(xyz, 1280).numpy()
    np.savez_compressed(os.path.join(HERE, "pointnet.npz"), **out)
    print("pointnet", {k: getattr(v, "shape", v) for k, v in out.items()})


def dataset_inputs(case):
    """Seeded float64 inputs of the dataset-side cases (regenerated identically by tests/test_dataset_ops.py)."""
    rng = np.random.RandomState(SEED + {"kitti": 0, "small": 1, "ties": 2}[case])
    if case == "kitti":      # KittiDataset.py:359-360: 10240 of the cloud's points -> 1280 nodes; 40960 points
        pc = rng.uniform(-40.0, 40.0, size=(3, 40960))
        return pc, pc[:, rng.choice(40960, 1280 * 8, replace=False)], 1280
    if case == "small":
        pc = rng.randn(3, 3001) * 10.0
        return pc, pc[:, rng.choice(3001, 777, replace=False)], 50
    pc = np.round(rng.randn(3, 2048) * 2.0)      # integer coordinates: many exact distance ties and duplicates
    return pc, pc[:, :1500], 64


def dataset_case():
    """FarthestSampler.sample of the REAL reference class + scipy's cKDTree nearest node (KittiDataset.py:107-126,
    :365-366) on seeded inputs."""
    from scipy.spatial import cKDTree
    kd = reference_loader.kitti_dataset()
    out = {}
    for case in ("kitti", "small", "ties"):
        pc, sub, k = dataset_inputs(case)
        state = np.random.get_state()
        np.random.seed(SEED)
        init_idx = np.random.randint(len(sub))          # what :118 is about to draw
        np.random.seed(SEED)
        node, idx = kd.FarthestSampler().sample(sub, k)
        np.random.set_state(state)
        assert idx[0] == init_idx
        out[case + "_init"] = np.int64(init_idx)
        out[case + "_fps_idx"] = idx
        out[case + "_fps_pts"] = node
        if case != "ties":                              # scipy leaves exact ties unspecified
            out[case + "_nearest"] = cKDTree(node.T).query(pc.T, k=1)[1].astype(np.int64)
        h = hashlib.sha256()
        h.update(np.ascontiguousarray(pc).tobytes())
        h.update(np.ascontiguousarray(sub).tobytes())
        out[case + "_sha"] = np.frombuffer(h.hexdigest().encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "dataset.npz"), **out)
    print("dataset.npz", {k: getattr(v, "shape", ()) for k, v in out.items()})


COST_VOLUME_CASES = {
    # name: (clouds, points, nlabel, R_amplitude, T_amplitude, every point masked in)
    # the reference hard-codes the 40 x 128 grid (dump bin 5120, IterModel.py:311) and 64 channels (:341)
    "sparse": (1, 4096, 3, 0.15, 2.0, False),
    "shared": (2, 3001, 3, 0.3, 3.0, False),      # two clouds, the first one's mask selects for both (:272); nlabel is odd (:29)
    "dense": (1, 8192, 3, 0.1, 1.0, True),
}


def cost_volume_inputs(case):
    """Seeded inputs of a cost-volume case: the batch, the cloud AT THE CURRENT POSE (data_batch['pc_i']; here
    the ground-truth registration, so that the sampled poses keep it in view), the mask and the scores."""
    B, N, nlabel, r_amp, t_amp, dense = COST_VOLUME_CASES[case]
    data = synth.make_batch(B, seed=SEED + 11, num_pt=N, img_h=160, img_w=512)
    g = torch.Generator().manual_seed(SEED + 12)
    scores = torch.rand(B, N, generator=g)
    mask = data["pc_overlap_pred"].clone()
    if dense:
        mask[:] = True
    P = data["P"][:, 0:3, :]
    pc_i = P[:, :, 0:3] @ data["pc"] + P[:, :, 3:4]
    amp = (torch.full((B, 1), r_amp), torch.full((B, 1), t_amp))
    return data, pc_i, mask, scores, nlabel, amp


def _reference_iter_model_pieces():
    """The statements of models/IterModel.py that make up the cost-volume warp, taken from the reference's OWN
    source through its syntax tree and compiled as they stand: the methods ``angle2matrix`` and ``sample_poses``
    (:96-172) and the part of ``forward`` from the mask selection to the cropped outputs (:272-351).  The class
    cannot be constructed without a GPU (``__init__`` calls ``.cuda()``, :29) and ``forward`` runs the whole model
    around these lines."""
    import ast
    path = os.path.join(reference_loader.REFERENCE_ROOT, "models", "IterModel.py")
    tree = ast.parse(open(path).read(), path)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "IterModel")
    fn = {n.name: n for n in cls.body if isinstance(n, ast.FunctionDef)}

    def assigns(node, name):
        return isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id == name for t in node.targets)

    body = fn["forward"].body
    first = next(i for i, n in enumerate(body) if assigns(n, "pc_mask"))
    # the warp ends where the image features enter (:353); the statement before crops the occupancy (:351)
    last = next(i for i, n in enumerate(body) if i > first and assigns(n, "img_geo_feat")) - 1
    assert assigns(body[last], "pc_warped_occupancy")
    block = ast.Module(body=body[first:last + 1], type_ignores=[])
    methods = ast.Module(body=[fn["angle2matrix"], fn["sample_poses"]], type_ignores=[])
    return compile(methods, path, "exec"), compile(block, path, "exec"), (body[first].lineno, body[last].end_lineno)


def cost_volume_case():
    """Run the reference's own lines on the CPU: ``Tensor.cuda`` is the identity while they execute and
    ``torch_scatter`` is the stand-in of oracle/shims.py (pinned with the environment fixtures)."""
    import types
    from oracle import shims
    methods_code, block_code, lines = _reference_iter_model_pieces()
    print("cost volume: IterModel.py:%d-%d" % lines)
    shims.install()
    import torch_scatter
    out = {"reference_lines": np.array(lines, dtype=np.int64)}
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for case in COST_VOLUME_CASES:
            data, pc_i, mask, scores, nlabel, (r_amp, t_amp) = cost_volume_inputs(case)
            ns = {"torch": torch, "np": np, "math": __import__("math")}
            exec(methods_code, ns)
            model = types.SimpleNamespace(nlabel=nlabel)
            half = (nlabel - 1) / 2                                  # IterModel.py:29 for this nlabel
            model.base = torch.from_numpy(np.array(range(int(-half), int(half) + 1))).unsqueeze(0)
            model.angle2matrix = types.MethodType(ns["angle2matrix"], model)
            model.sample_poses = types.MethodType(ns["sample_poses"], model)
            batch = {"pc_overlap_pred": mask, "pc_overlap_pred_standby": mask, "pc_i": pc_i, "K": data["K"],
                     "img": data["img"], "pc_geo_feat": data["pc_geo_feat"], "pc_is_in_cam_scores": scores,
                     "R_amplitude": r_amp, "T_amplitude": t_amp}
            env = {"torch": torch, "torch_scatter": torch_scatter, "self": model, "data_batch": batch}
            exec(block_code, env)
            wf, occ = env["pc_warped_geo_feat"], env["pc_warped_occupancy"]
            assert tuple(wf.shape) == (pc_i.shape[0], nlabel ** 3, 64, 5120) and float(occ.sum()) > 0
            out[case + "_poses"] = env["delta_RT"].numpy()           # sample_poses' output, [B, nlabel^3, 3, 4]
            out[case + "_occupancy"] = occ.numpy()
            out[case + "_features_sha"] = np.frombuffer(sha(wf).encode(), dtype=np.uint8)
            # the channels of a few hundred pixels in full (the rest is covered by the hash)
            flat = wf.permute(0, 1, 3, 2).reshape(-1, 64)
            rows = torch.nonzero(occ.reshape(-1) > 0)[:, 0][::37][:512]
            out[case + "_sample_rows"] = rows.numpy()
            out[case + "_sample_features"] = flat[rows].numpy()
            out[case + "_inputs_sha"] = np.frombuffer(
                sha(data["pc"], data["pc_geo_feat"], mask, data["K"], pc_i, scores).encode(), dtype=np.uint8)
    finally:
        torch.Tensor.cuda = real_cuda
    np.savez_compressed(os.path.join(HERE, "cost_volume.npz"), **out)
    print("cost_volume.npz", {k: getattr(v, "shape", ()) for k, v in out.items()},
          os.path.getsize(os.path.join(HERE, "cost_volume.npz")) // 1024, "KiB")


def tower_inputs():
    """Seeded weights of the four blocks (oracle/tower_oracle.make_state) and an obs3d [2, 5, 4096] made the way the
    environment makes it: xyz, predicted-overlap flag, in-frustum flag (environment.py:121-124)."""
    from oracle import tower_oracle
    states = [tower_oracle.make_state(SEED + 40 + i, cin, cout) for i, (cin, cout) in enumerate(tower_oracle.TOWER)]
    data = synth.make_batch(2, seed=SEED + 21, num_pt=4096, img_h=160, img_w=512)
    g = torch.Generator().manual_seed(SEED + 22)
    in_cam = (torch.rand(2, 4096, generator=g) < 0.3).float()
    obs3d = torch.cat([data["pc"], data["pc_overlap_pred"].float().unsqueeze(1), in_cam.unsqueeze(1)], dim=1).contiguous()
    return states, obs3d


def reference_tower(states, obs3d):
    """The reference's own ConvBNReLURes1D modules (models/PointNN.py:260-282) in eval mode, loaded with `states` and
    composed exactly as CMRAgent.forward does (models/CMRAgent.py:92-101; the module list of :25-29)."""
    import importlib
    import types
    from oracle import shims, tower_oracle
    shims.install()
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.image", "tensorboardX"):
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
    if reference_loader.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, reference_loader.REFERENCE_ROOT)
    pn = importlib.import_module("models.PointNN")
    layers = []
    for sd, (cin, cout) in zip(states, tower_oracle.TOWER):
        m = pn.ConvBNReLURes1D(cin, cout)
        m.load_state_dict(sd)
        layers.append(m.eval())
    with torch.no_grad():
        embed_3d = obs3d
        for step, layer in enumerate(layers):                     # CMRAgent.py:94-100
            feat_3d = layer(embed_3d)
            embed_3d = torch.max(feat_3d, dim=2, keepdim=True)[0]
            if step < len(layers) - 1:
                embed_3d = embed_3d.repeat(1, 1, feat_3d.shape[2])
                embed_3d = torch.cat([feat_3d, embed_3d], dim=1)
        return embed_3d.view(embed_3d.shape[0], -1)               # :101


def tower_case():
    states, obs3d = tower_inputs()
    want = reference_tower(states, obs3d)
    np.savez_compressed(os.path.join(HERE, "tower.npz"), embed_3d=want.numpy(),
                        obs3d_sha=np.frombuffer(sha(obs3d).encode(), dtype=np.uint8))
    print("tower.npz", tuple(want.shape))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "tower":   # only the tower fixture
        assert reference_loader.available(), "needs /root/reference"
        tower_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "cost_volume":   # only the cost-volume fixture
        assert reference_loader.available(), "needs /root/reference"
        cost_volume_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "dataset":   # only the dataset-side fixture
        assert reference_loader.available(), "needs /root/reference"
        dataset_case()
        sys.exit(0)
    assert reference_loader.available(), "needs /root/reference"
    torch.set_num_threads(1)
    for name, (b, shape, iters, full) in ENV_CASES.items():
        env_case(name, b, shape, iters, full)
    step_case()
    pointnet_case()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
    dataset_case()
    cost_volume_case()
    tower_case()
