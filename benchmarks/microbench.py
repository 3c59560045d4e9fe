#!/usr/bin/env python
"""Micro-benchmarks for BASELINE.json configs 4 and 5 (not the driver's bench contract - that is
bench.py).  Prints one JSON line per case; all timings are CUDA events on the launch stream after
warm-up, inputs resident in HBM.

  python benchmarks/microbench.py frontend [--batch 128]   FPS 40960->1280 + kNN k=64 + grouping
  python benchmarks/microbench.py sweep                    observe (project + tile scatter), 16K-128K points
  python benchmarks/microbench.py env [--batch 32]         per-kernel times of one rollout iteration
  python benchmarks/microbench.py cost_volume              IterModel's 729-pose warp of one KITTI cloud
  python benchmarks/microbench.py tower [--batch 32]       the agent's 3-D tower (tcgen05) vs the reference's modules on cuDNN
  python benchmarks/microbench.py sample [--batch 32]      image features sampled bilinearly at the points' projections
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from cmr_agent_b200 import _lib, synth  # noqa: E402
from cmr_agent_b200 import environment as env  # noqa: E402
from cmr_agent_b200 import pointnet_util as pn  # noqa: E402

PEAK = 6553.0
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def time_cuda(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2] / 1e3


def time_pair(fa, fb, warm=3, reps=10):
    """fa(); fb() back to back, an event between them: median seconds of each half."""
    for _ in range(warm):
        fa(); fb()
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(reps)]
    for e in ev:
        e[0].record(); fa(); e[1].record(); fb(); e[2].record()
    torch.cuda.synchronize()
    ta = sorted(e[0].elapsed_time(e[1]) for e in ev)
    tb = sorted(e[1].elapsed_time(e[2]) for e in ev)
    return ta[len(ta) // 2] / 1e3, tb[len(tb) // 2] / 1e3


def frontend(args):
    dev = torch.device("cuda:0")
    B, N, S, K = args.batch, 40960, 1280, 64
    g = torch.Generator().manual_seed(1)
    # clouds with the KITTI-shaped distribution; generated once on the host
    xyz = synth.make_cloud_batch(min(B, 8), num_pt=N, seed=2023)
    xyz = xyz.repeat((B + xyz.shape[0] - 1) // xyz.shape[0], 1, 1)[:B].contiguous()
    xyz += torch.randn(B, 1, 3, generator=g) * 0.01          # distinct clouds
    xyz = xyz.to(dev)
    start = torch.randint(0, N, (B,), generator=g).to(dev)
    feats = torch.randn(B, N, 3, generator=g).to(dev)
    t_fps = time_cuda(lambda: pn.farthest_point_sample_from(xyz, S, start), reps=5)
    t_fps_cluster = time_cuda(lambda: pn.farthest_point_sample_from(xyz, S, start, method="cluster"), reps=3)
    t_knn_brute = time_cuda(lambda: pn.knn_point(K, xyz, pn.index_points(xyz, pn.farthest_point_sample_from(xyz, S, start)), method="brute"), reps=3)
    fps = pn.farthest_point_sample_from(xyz, S, start)
    new_xyz = pn.index_points(xyz, fps)
    t_knn = time_cuda(lambda: pn.knn_point(K, xyz, new_xyz), reps=5)
    idx = pn.knn_point(K, xyz, new_xyz)
    t_ball = time_cuda(lambda: pn.query_ball_point(1.0, K, xyz, new_xyz), reps=5)
    t_grp = time_cuda(lambda: pn.group_points(xyz, feats, new_xyz, idx), reps=5)
    t_idx = time_cuda(lambda: pn.index_points(xyz, idx), reps=5)
    out = {
        "bench": "frontend", "batch": B, "N": N, "npoint": S, "k": K,
        "fps_ms": t_fps * 1e3, "fps_cluster_kernel_ms": t_fps_cluster * 1e3, "knn_brute_plus_fps_ms": t_knn_brute * 1e3, "fps_rounds_per_s": B * S / t_fps, "fps_clouds_per_s": B / t_fps,
        "fps_byte_floor_frac": (12.0 * N + 8 * S) * B / t_fps / 1e9 / PEAK,
        "knn_ms": t_knn * 1e3, "knn_queries_per_s": B * S / t_knn, "knn_pair_evals_per_s": B * S * N / t_knn,
        "knn_byte_floor_frac": (12.0 * (N + S) + 8 * S * K) * B / t_knn / 1e9 / PEAK,
        "ball_ms": t_ball * 1e3, "group_ms": t_grp * 1e3, "index_points_ms": t_idx * 1e3,
        "group_gbs": (8.0 * S * K + 4.0 * S * K * 6 * 2) * B / t_grp / 1e9,
        "total_ms": (t_fps + t_knn + t_grp) * 1e3,
    }
    print(json.dumps(out), flush=True)


def sweep(args):
    dev = torch.device("cuda:0")
    for N in (16384, 32768, 65536, 131072):
        for frac in (0.25, 1.0):
            B = max(1, min(256, int(1.2e9 // (33 * N + 12 * 64 * 5120 + 256 * N * frac * 0.3))))
            cpu = synth.make_batch(min(B, 4), seed=7, num_pt=N, img_h=160, img_w=512)
            rep = (B + 3) // 4
            data = {}
            for k in ("pc", "pc_overlap_pred", "pc_geo_feat", "img_geo_feat"):
                data[k] = cpu[k].repeat(rep, *([1] * (cpu[k].dim() - 1)))[:B].contiguous().to(dev)
            data["K"] = cpu["K"].repeat(rep, 1, 1)[:B].contiguous()
            data["img"] = torch.zeros(1, 1, 1, 1).expand(B, 3, 160, 512)
            if frac == 1.0:
                data["pc_overlap_pred"][:] = True
            pose = torch.eye(4, device=dev).repeat(B, 1, 1)
            pose[:, 2, 3] = 3.0
            o = env.observation_from_a_pose(data, pose, return_pixels=True)
            mvis = int(o[3].sum())
            ep = env.episode_state(data)
            st = _lib.stream
            obs2d = torch.empty(B, 128, 40, 128, device=dev)
            obs3d = torch.empty(B, 5, N, device=dev)
            p = _lib.ptr

            copied = ctypes.c_int(0)   # 1: k_project carried the image half of obs2d, 0: k_tile_gather does

            def proj():
                _lib.call("cmr_project", p(ep.pc), p(ep.overlap), p(ep.K), p(pose), p(ep.mean), p(ep.ws), B, N, 64, 40,
                          128, p(obs3d), None, None, p(ep.img_feat), p(obs2d), ctypes.byref(copied), 1, st())

            def scat():
                _lib.call("cmr_tile_scatter", p(ep.img_feat), p(ep.K), p(ep.ws), B, N, 64, 40, 128,
                          0 if copied.value else 1, p(obs2d), st())

            t_p, t_s = time_pair(proj, scat)   # a scatter consumes what ONE project left
            img_bytes = 8.0 * 64 * 5120 * B                  # image half of obs2d, in and out (TMA), by whoever carries it
            bytes_p = 33.0 * N * B + (img_bytes if copied.value else 0.0)                     # obs3d stream (+ image)
            bytes_s = 4.0 * 64 * mvis + 4.0 * 64 * 5120 * B + (0.0 if copied.value else img_bytes)   # rows + projected half
            print(json.dumps({
                "bench": "sweep", "N": N, "overlap_frac": frac, "batch": B, "m_vis_per_episode": mvis / B,
                "image_half_in": "k_project" if copied.value else "k_tile_gather",
                "project_us": t_p * 1e6, "project_gbs": bytes_p / t_p / 1e9, "project_frac": bytes_p / t_p / 1e9 / PEAK,
                "scatter_us": t_s * 1e6, "scatter_gbs": bytes_s / t_s / 1e9, "scatter_frac": bytes_s / t_s / 1e9 / PEAK,
                "observe_steps_per_s": B / (t_p + t_s),
                "observe_frac": (bytes_p + bytes_s) / (t_p + t_s) / 1e9 / PEAK}), flush=True)
            del data, obs2d, obs3d, ep
            torch.cuda.empty_cache()


def env_kernels(args):
    dev = torch.device("cuda:0")
    B, N = args.batch, 40960
    cpu = synth.make_batch(B, seed=2023, num_pt=N, img_h=160, img_w=512)
    data = dict(cpu)
    for k in ("pc", "pc_overlap_pred", "pc_geo_feat", "img_geo_feat"):
        data[k] = cpu[k].to(dev)
    cfg = synth.StepConfig(device=dev)
    a_r, a_t = synth.make_actions(B, 1)
    a_r, a_t = a_r[0].to(dev), a_t[0].to(dev)
    pose, _ = env.init(data)
    o = env.observation_from_a_pose(data, pose, return_pixels=True)
    mvis = int(o[3].sum())
    ep = env.episode_state(data)
    p, st = _lib.ptr, _lib.stream
    obs2d = torch.empty(B, 128, 40, 128, device=dev)
    obs3d = torch.empty(B, 5, N, device=dev)
    feat = data["pc_geo_feat"]
    res = {"bench": "env", "batch": B, "m_vis_per_episode": mvis / B, "m_per_episode": float(cpu["pc_overlap_pred"].sum()) / B}

    def rec(name, fn, nbytes):
        t = time_cuda(fn, reps=20)
        res[name + "_us"] = t * 1e6
        res[name + "_gbs"] = nbytes / t / 1e9
        res[name + "_frac"] = nbytes / t / 1e9 / PEAK

    M = float(cpu["pc_overlap_pred"].sum())
    rec("prepare", lambda: _lib.call("cmr_episode_prepare", p(ep.overlap), p(feat), B, N, 64, p(ep.ws), st()),
        B * N * (1 + 256.0) + M * 256.0)
    copied = ctypes.c_int(0)
    t_p, t_s = time_pair(
        lambda: _lib.call("cmr_project", p(ep.pc), p(ep.overlap), p(ep.K), p(pose), p(ep.mean), p(ep.ws), B,
                          N, 64, 40, 128, p(obs3d), None, None, p(ep.img_feat), p(obs2d), ctypes.byref(copied), 1, st()),
        lambda: _lib.call("cmr_tile_scatter", p(ep.img_feat), p(ep.K), p(ep.ws), B, N, 64, 40, 128,
                          0 if copied.value else 1, p(obs2d), st()),
        reps=20)
    img_bytes = 8.0 * 64 * 5120 * B
    res["image_half_in"] = "k_project" if copied.value else "k_tile_gather"
    for name, t, nbytes in (("project", t_p, 33.0 * N * B + (img_bytes if copied.value else 0.0)),
                            ("scatter", t_s, 256.0 * mvis + 4.0 * 64 * 5120 * B + (0.0 if copied.value else img_bytes))):
        res[name + "_us"], res[name + "_gbs"], res[name + "_frac"] = t * 1e6, nbytes / t / 1e9, nbytes / t / 1e9 / PEAK
    rec("step", lambda: env.step(a_r, a_t, pose, cfg), 64.0 * B)
    env.reward(pose, data, None)
    rec("reward", lambda: env.reward(pose, data, None), 25.0 * N * B)
    rec("observe_api", lambda: env.observation_from_a_pose(data, pose), 33.0 * N * B + 256.0 * mvis + 12.0 * 64 * 5120 * B)
    print(json.dumps(res), flush=True)


def cost_volume_bench(args):
    """models/IterModel.py:272-351 at the reference's sizes: one KITTI cloud, 9^3 = 729 candidate poses."""
    import time
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_cost_volume import _case
    from cmr_agent_b200 import cost_volume
    from oracle import cost_volume_oracle as cvo
    dev = torch.device("cuda:0")
    nlabel = 9
    data, mask, poses, scores = _case(1, 40960, 160, 512, nlabel, 3)
    H, W, Kp = 40, 128, nlabel ** 3
    dargs = (data["pc"].to(dev), mask.to(dev), poses.to(dev), data["K"], data["pc_geo_feat"].to(dev), scores.to(dev), H, W)
    t = time_cuda(lambda: cost_volume.warp(*dargs), warm=2, reps=5)
    wf, occ = cost_volume.warp(*dargs)
    vis = float((occ > 0).sum()) / Kp
    M = int(mask.sum())
    out_bytes = Kp * 68.0 * H * W * 4
    # CPU: the reference's own chunking (200 poses at a time), here 27 poses timed and scaled
    t0 = time.time()
    cvo.warp(data["pc"], mask, poses[:, :27], data["K"], data["pc_geo_feat"], scores, H, W)
    t_cpu = (time.time() - t0) * Kp / 27
    print(json.dumps({"bench": "cost_volume", "poses": Kp, "N": 40960, "masked_points": M, "occupied_pixels_per_pose": vis,
                      "warp_ms": t * 1e3, "poses_per_s": Kp / t, "output_gbs": out_bytes / t / 1e9,
                      "output_frac_of_hbm": out_bytes / t / 1e9 / PEAK, "cpu_port_ms_scaled_from_27_poses": t_cpu * 1e3,
                      "cpu_threads": torch.get_num_threads()}), flush=True)


def sample_bench(args):
    """environment.sample_image_features (north_star's point-side bilinear gather; an extension) at KITTI
    size, ground-truth pose (about 30 % of the points inside the frustum).  Bytes: 12N in, 4CN + N out per episode."""
    from oracle import env_oracle as eo
    dev = torch.device("cuda:0")
    B, N, C = args.batch, 40960, 64
    cpu = synth.make_batch(min(B, 8), seed=2023, num_pt=N, img_h=160, img_w=512)
    rep = (B + cpu["pc"].shape[0] - 1) // cpu["pc"].shape[0]
    data = {}
    for k in ("pc", "pc_overlap_pred", "pc_geo_feat", "img_geo_feat", "K", "P"):
        data[k] = cpu[k].repeat(rep, *([1] * (cpu[k].dim() - 1)))[:B].contiguous()
    pose = eo.to_disentangled(data["P"].clone(), data["pc"]).contiguous().to(dev)
    for k in ("pc", "pc_overlap_pred", "pc_geo_feat", "img_geo_feat"):
        data[k] = data[k].to(dev)
    data["img"] = torch.zeros(1, 1, 1, 1).expand(B, 3, 160, 512)
    feats, cam = env.sample_image_features(data, pose)            # builds the pixel-major image once
    t = time_cuda(lambda: env.sample_image_features(data, pose), warm=3, reps=20)
    byts = (12.0 * N + 4.0 * C * N + N) * B
    print(json.dumps({"bench": "sample", "batch": B, "N": N, "C": C, "in_frustum_frac": float(cam.float().mean()),
                      "sample_us": t * 1e6, "points_per_s": B * N / t, "gbs": byts / t / 1e9, "frac_of_hbm": byts / t / 1e9 / PEAK}),
          flush=True)


def tower_bench(args):
    """models/CMRAgent.py:92-101 at KITTI size: cmr_tower_forward against the reference's own ConvBNReLURes1D modules
    (oracle/_ref, eager cuDNN/cuBLAS on the same GPU, TF32 on - torch's default - and off).
    FLOPs counted: the dense products the reference evaluates per point (2 x MACs, full 128-wide first layers)."""
    from cmr_agent_b200 import agent_tower
    from oracle import reference_loader as rl, tower_oracle as to
    dev = torch.device("cuda:0")
    B, N = args.batch, 40960
    g = torch.Generator().manual_seed(3)
    obs3d = torch.cat([(torch.rand(B, 3, N, generator=g) - 0.5) * 160, (torch.rand(B, 2, N, generator=g) < 0.3).float()], 1).to(dev)
    states = [to.make_state(50 + i, ci, co) for i, (ci, co) in enumerate(to.TOWER)]
    tower = agent_tower.Tower3D(states, dev)
    t = time_cuda(lambda: tower(obs3d), warm=3, reps=20)
    macs_ref = sum(ci * ci + ci * co + (ci * co if ci != co else 0) for ci, co in to.TOWER)       # per point, as the reference computes
    macs_here = 5 * 5 + 2 * 5 * 64 + 2 * (64 * 128 + 128 * 64 + 64 * 64) + (64 * 128 + 128 * 128)  # the repeated max is a bias
    res = {"bench": "tower", "B": B, "N": N, "tower_ms": t * 1e3, "points_per_s": B * N / t,
           "tflops_reference_count": 2.0 * macs_ref * B * N / t / 1e12, "tflops_executed_logical": 2.0 * macs_here * B * N / t / 1e12,
           "tensor_passes_per_product": 3, "tflops_tensor_pipe": 3 * 2.0 * (macs_here - 665) * B * N / t / 1e12,
           "hbm_bytes_per_point": 20 + 256 * 6, }   # obs3d + three feature maps (two fp16 planes) written and read once
    res["hbm_gbs"] = res["hbm_bytes_per_point"] * B * N / t / 1e9
    res["hbm_frac"] = res["hbm_gbs"] / PEAK
    try:
        bf16 = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
        res["tensor_frac_of_bf16_sustained"] = res["tflops_tensor_pipe"] / bf16
    except Exception:
        pass
    if rl.available():
        rl.put_on_path()
        import importlib
        pnn = importlib.import_module("models.PointNN")
        layers = []
        for sd, (ci, co) in zip(states, to.TOWER):
            m = pnn.ConvBNReLURes1D(ci, co)
            m.load_state_dict(sd)
            layers.append(m.to(dev).eval())

        def ref():                                            # CMRAgent.py:92-101 verbatim control flow
            with torch.no_grad():
                embed_3d = obs3d
                for step, layer in enumerate(layers):
                    feat_3d = layer(embed_3d)
                    embed_3d = torch.max(feat_3d, dim=2, keepdim=True)[0]
                    if step < len(layers) - 1:
                        embed_3d = embed_3d.repeat(1, 1, feat_3d.shape[2])
                        embed_3d = torch.cat([feat_3d, embed_3d], dim=1)
                return embed_3d.view(embed_3d.shape[0], -1)
        want64 = None
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            tr = time_cuda(ref, warm=2, reps=5)
            out = ref()
            res[f"reference_eager_ms_tf32_{'on' if tf32 else 'off'}"] = tr * 1e3
            res[f"speedup_vs_reference_tf32_{'on' if tf32 else 'off'}"] = tr / t
            got = tower(obs3d)
            res[f"scaled_err_vs_reference_tf32_{'on' if tf32 else 'off'}"] = float(
                ((got - out).abs().max(dim=1)[0] / out.abs().max(dim=1)[0]).max())
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["frontend", "sweep", "env", "cost_volume", "tower", "sample"])
    ap.add_argument("--batch", type=int, default=None)
    a = ap.parse_args()
    if a.batch is None:
        a.batch = 128 if a.which == "frontend" else 32
    {"frontend": frontend, "sweep": sweep, "env": env_kernels, "cost_volume": cost_volume_bench, "tower": tower_bench, "sample": sample_bench}[a.which](a)
