#!/usr/bin/env python
"""bench.py - registration steps/sec of the CMR-Agent geometric hot path on B200.

Workload (BASELINE.json configs[1]): KITTI-shaped batch of 32 episodes x 10 agent iterations per GPU
(40960-pt clouds, 160x512 image -> 40x128 feature grid, 64 channels), synthetic data.
One bench "step" = one rollout of that batch: per-episode prepare (cloud mean + predicted-overlap
compaction, done once per episode like the drop-in does) + 10 x (observation_from_a_pose + step +
reward) = batch*10 registration steps.  Episodes shard across GPUs by rank with no
data-path collective ("weak" scaling: 32 episodes per GPU).

  value  : registration steps/s with all inputs resident in HBM, kernels called through the C ABI; the rollout is
           captured once as a CUDA graph and replayed.  --streams 2 (default): the reward of an iteration runs on
           a second stream beside the next observation (observe and step keep their order); the replayed rollout's
           outputs are compared bit for bit with a one-stream eager rollout before anything is timed.
  e2e    : the same rollout through the reference-facing drop-in API (cmr_agent_b200.environment)
           from pinned HOST tensors: H2D of every input of the rollout and D2H of the per-iteration
           reward/distance and the final poses are inside the timed region.
  roofline: the slower of the two observe stages (k_project | k_tile_gather), timed live with CUDA
           events on the launch stream in an instrumented pass (one stream, every rollout queued behind a spin
           kernel so that the host's enqueue rate does not show); the other one is roofline_secondary.
  cpu_baseline: the oracle's torch-CPU port of the reference path (oracle/env_oracle.py) on this
           box's host cores, bounded sample.
`--impl reference` times that CPU port alone (the reference is pure Python and cannot travel to the
GPU box; the port restates it operator for operator and is pinned bit-exact against it).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

from cmr_agent_b200 import dist as cdist  # noqa: E402
from cmr_agent_b200 import synth  # noqa: E402

METRIC = "registration steps/sec (40960 pts, 160x512 img)"
UNIT = "steps/s"
SHAPE = dict(num_pt=40960, img_h=160, img_w=512, channels=64)
SEED = 2023


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="episodes per GPU")
    ap.add_argument("--iters", type=int, default=10, help="agent iterations per episode (config.action_num)")
    ap.add_argument("--cpu-episodes", type=int, default=32, help="episodes in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--streams", type=int, default=2, choices=[1, 2],
                    help="2: the reward runs on a second stream beside the next observation (default); 1: one stream")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < 10.0:      # nvidia-smi takes a moment to come up
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln))

    def mark(self):
        return time.time()

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ts, ln in self.lines:
            if ts < t_begin or ts > t_end + 0.12:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for name, val in zip(self.NAMES, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def load_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """dram bytes per launch, {kernel: bytes}, from the committed ncu --set full capture, if any."""
    path = os.path.join(ROOT, "profiles", "dram_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


# ------------------------------------------------------------------------------- CPU baseline (port)
def cpu_rollout(data, a_r, a_t, iters, cfg):
    """One rollout of the reference path on the CPU through the oracle port: init, to_disentangled,
    iters x (observation_from_a_pose + step + reward).  Returns registration steps done."""
    from oracle import env_oracle as eo
    pose, target = eo.init(data)
    eo.to_disentangled(target, data["pc"])
    prev = None
    for it in range(iters):
        eo.observation_from_a_pose(data, pose)
        eo.step(a_r[it], a_t[it], pose, cfg)
        _, prev = eo.reward(pose, data, prev)
    return data["pc"].shape[0] * iters


def time_cpu_port(episodes, iters, repeats, threads_options):
    data = synth.make_batch(episodes, seed=SEED, **SHAPE)
    a_r, a_t = synth.make_actions(episodes, iters, seed=SEED)
    cfg = synth.StepConfig()
    best = None
    for nt in threads_options:
        torch.set_num_threads(nt)
        cpu_rollout(data, a_r, a_t, iters, cfg)  # warm-up
        times = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            n = cpu_rollout(data, a_r, a_t, iters, cfg)
            times.append(time.perf_counter() - t0)
        rate = n / statistics.median(times)
        if best is None or rate > best[0]:
            best = (rate, nt, statistics.median(times))
    return best


def run_reference_arm(args, rank):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is pure
    Python (cannot travel to this box), so the oracle's operator-for-operator port is timed."""
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    episodes, iters = args.cpu_episodes, args.iters
    data = synth.make_batch(episodes, seed=SEED, **SHAPE)
    a_r, a_t = synth.make_actions(episodes, iters, seed=SEED)
    cfg = synth.StepConfig()
    # pick the better of {all cores, 1 thread} once, during warm-up (multi-threading hurts the small ops)
    cand = {}
    for nt in sorted({ncores, 1}, reverse=True):
        torch.set_num_threads(nt)
        cpu_rollout(data, a_r, a_t, iters, cfg)
        t0 = time.perf_counter()
        cpu_rollout(data, a_r, a_t, iters, cfg)
        cand[nt] = time.perf_counter() - t0
    nt = min(cand, key=cand.get)
    torch.set_num_threads(nt)
    for _ in range(max(args.warmup - 2, 0)):
        cpu_rollout(data, a_r, a_t, iters, cfg)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        done += cpu_rollout(data, a_r, a_t, iters, cfg)
    dt = time.perf_counter() - t0
    value = done / dt
    sample = f"{episodes} episodes x {iters} iterations per step (bounded sample of the 32x10 workload)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "kitti_b32x10", "episodes_per_gpu": 32, "iterations": iters,
                   "num_pt": SHAPE["num_pt"], "image": "160x512", "grid": "40x128", "channels": 64,
                   "reference_sample_episodes_per_step": episodes},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nt, "kind": "port", "sample": sample,
                         "host_cores": ncores},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------- B200 arm
class DeviceRollout:
    """The rollout with every input resident in HBM, calling libcmr_b200.so through its C ABI."""

    def __init__(self, cpu, a_r, a_t, dev, iters):
        from cmr_agent_b200 import _lib, environment as env
        self.lib = _lib
        self.iters = iters
        self.dev = dev
        B, _, N = cpu["pc"].shape
        C = cpu["pc_geo_feat"].shape[1]
        H, W = cpu["img"].shape[2] // 4, cpu["img"].shape[3] // 4
        self.dims = (B, N, C, H, W)
        g = lambda k: cpu[k].to(dev)  # noqa: E731
        self.pc, self.feat, self.img_feat = g("pc"), g("pc_geo_feat"), g("img_geo_feat")
        self.overlap = g("pc_overlap_pred").contiguous().view(torch.uint8)
        self.K = g("K").contiguous()
        self.target = g("pc_in_cam_space").contiguous()
        self.mask = (g("pc_mask") != 0).contiguous().view(torch.uint8)
        self.a_r, self.a_t = a_r.to(dev).contiguous(), a_t.to(dev).contiguous()
        rot, tt = env.build_step_tables(synth.StepConfig().r_steps, synth.StepConfig().t_steps)
        self.rot, self.tt, self.nbins = rot.to(dev), tt.to(dev), int(tt.shape[0])
        lib = _lib.load()
        self.ws = torch.empty(lib.cmr_workspace_bytes(B, N, C, H * W), dtype=torch.uint8, device=dev)
        self.scratch = torch.zeros(lib.cmr_reward_scratch_bytes(B), dtype=torch.uint8, device=dev)
        self.obs2d = torch.empty(B, 2 * C, H, W, device=dev)
        self.obs3d = torch.empty(B, 5, N, device=dev)
        self.pose = torch.empty(B, 4, 4, device=dev)
        self.eye = torch.eye(4, device=dev).repeat(B, 1, 1)
        self.mvis = torch.zeros(iters, B, dtype=torch.int32, device=dev)
        self.rew = torch.empty(iters, B, device=dev)
        self.dist = torch.empty(iters, B, device=dev)
        self.mean = torch.empty(B, 3, device=dev)
        self.copied = ctypes.c_int(0)
        self.two_streams = False
        self.side = torch.cuda.Stream(dev)

    def run(self, events=None, count_visible=True):
        """events: per iteration (e0, e1, e2) recorded before cmr_project, between the two observe kernels
        and after cmr_tile_scatter, on the launch stream.  count_visible: also count the visible
        predicted-overlap points per episode (M_vis of the roofline; one memset + a few atomics per step)."""
        L, p, st = self.lib, self.lib.ptr, self.lib.stream()
        B, N, C, H, W = self.dims
        L.call("cmr_cloud_mean", p(self.pc), B, N, p(self.mean), st)    # environment.py:46 - once per episode
        L.call("cmr_episode_prepare", p(self.overlap), p(self.feat), B, N, C, p(self.ws), st)
        self.pose.copy_(self.eye)                                    # env.init
        for it in range(self.iters):
            if events is not None:
                events[it][0].record()
            mv = p(self.mvis[it]) if count_visible else None
            prev = p(self.dist[it - 1]) if it else None
            if events is None and self.two_streams:
                # the reward is a training signal nothing in the loop waits for: it runs on a second stream beside
                # the next projection and scatter.  Everything the agent's next action would depend on (observe,
                # then step) keeps its order on the first stream.  (A C-ABI host owns the streams it passes in.)
                main, side = torch.cuda.current_stream(), self.side
                sst = ctypes.c_void_p(side.cuda_stream)
                L.call("cmr_observe", p(self.pc), p(self.overlap), p(self.img_feat), p(self.K), p(self.pose), p(self.mean),
                       p(self.ws), B, N, C, H, W, p(self.obs2d), p(self.obs3d), None, mv, st)
                if it:
                    main.wait_event(rewarded)   # the previous reward may still be reading the pose this step rewrites
                L.call("cmr_step", p(self.pose), p(self.a_r[it]), p(self.a_t[it]), p(self.rot), p(self.tt), self.nbins,
                       0, B, st)
                stepped = torch.cuda.Event()
                stepped.record(main)
                side.wait_event(stepped)
                L.call("cmr_reward", p(self.target), p(self.pc), p(self.mask), p(self.mean), p(self.pose), prev, 0, B, N,
                       p(self.scratch), p(self.rew[it]), p(self.dist[it]), sst)
                rewarded = torch.cuda.Event()
                rewarded.record(side)
                if it == self.iters - 1:
                    main.wait_stream(side)
                continue
            if events is None:   # the product call
                L.call("cmr_observe", p(self.pc), p(self.overlap), p(self.img_feat), p(self.K), p(self.pose), p(self.mean),
                       p(self.ws), B, N, C, H, W, p(self.obs2d), p(self.obs3d), None, mv, st)
            else:                # the same two launches with an event between them
                L.call("cmr_project", p(self.pc), p(self.overlap), p(self.K), p(self.pose), p(self.mean), p(self.ws),
                       B, N, C, H, W, p(self.obs3d), None, mv, p(self.img_feat), p(self.obs2d),
                       ctypes.byref(self.copied), 1, st)   # 1 = CMR_PROJECT_PAIRED
                events[it][1].record()
                L.call("cmr_tile_scatter", p(self.img_feat), p(self.K), p(self.ws), B, N, C, H, W,
                       0 if self.copied.value else 1, p(self.obs2d), st)
                events[it][2].record()
            L.call("cmr_step", p(self.pose), p(self.a_r[it]), p(self.a_t[it]), p(self.rot), p(self.tt), self.nbins,
                   0, B, st)
            L.call("cmr_reward", p(self.target), p(self.pc), p(self.mask), p(self.mean), p(self.pose), prev, 0, B, N,
                   p(self.scratch), p(self.rew[it]), p(self.dist[it]), st)


class HostRollout:
    """The same rollout through the drop-in API from pinned host memory (the e2e arm)."""

    DEVICE_KEYS = ("pc", "pc_overlap_pred", "pc_geo_feat", "img_geo_feat")

    def __init__(self, cpu, a_r, a_t, dev, iters, features_resident=False):
        """features_resident: the two feature tensors and the predicted-overlap mask - which CMR-Agent's feature
        network produces ON the device (models/MultiHeadModel.py:234-241) - stay in HBM between steps; only what the
        DataLoader delivers on the host (cloud, intrinsics, ground truth) is uploaded every step."""
        self.features_resident = features_resident
        self.cpu = {k: (v.pin_memory() if (isinstance(v, torch.Tensor) and k != "img") else v) for k, v in cpu.items()}
        self.a_r, self.a_t = a_r.pin_memory(), a_t.pin_memory()
        self.dev, self.iters = dev, iters
        B = cpu["pc"].shape[0]
        up = (("pc",) if features_resident else self.DEVICE_KEYS) + ("K", "P", "pc_in_cam_space", "pc_mask")
        self.h2d = sum(self.cpu[k].numel() * self.cpu[k].element_size() for k in up)
        self.resident = {k: self.cpu[k].to(dev) for k in self.DEVICE_KEYS if k != "pc"} if features_resident else {}
        self.h2d += self.a_r.numel() * 8 + self.a_t.numel() * 8
        self.d2h = iters * B * 4 * 2 + B * 16 * 4
        self.out_rew = torch.empty(iters, B, 1, 1).pin_memory()
        self.out_dist = torch.empty(iters, B, 1, 1).pin_memory()
        self.out_pose = torch.empty(B, 4, 4).pin_memory()
        self.cfg = synth.StepConfig(device=dev)

    def run(self):
        from cmr_agent_b200 import environment as env
        data = dict(self.cpu)                                        # a fresh dict per batch, like the DataLoader's
        for k in self.DEVICE_KEYS:                                   # what the feature network leaves on the device
            data[k] = self.resident[k] if k in self.resident else self.cpu[k].to(self.dev, non_blocking=True)
        a_r = self.a_r.to(self.dev, non_blocking=True)
        a_t = self.a_t.to(self.dev, non_blocking=True)
        pose, target = env.init(data)                                # H2D of P
        env.to_disentangled(target, data["pc"])
        prev = None
        for it in range(self.iters):
            env.observation_from_a_pose(data, pose)                  # first call: H2D of K + per-episode prepare
            env.step(a_r[it], a_t[it], pose, self.cfg)
            rew, prev = env.reward(pose, data, prev)                 # first call: H2D of pc_in_cam_space, pc_mask
            self.out_rew[it].copy_(rew, non_blocking=True)           # D2H of the step's result
            self.out_dist[it].copy_(prev, non_blocking=True)
        self.out_pose.copy_(pose, non_blocking=True)
        torch.cuda.current_stream().synchronize()                    # the host reads the results


def timed(fn, steps, dev):
    cdist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    cdist.barrier()
    return e0.elapsed_time(e1) / 1e3


def run_b200_arm(args, rank, world, local):
    from cmr_agent_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    _lib.load()
    B, iters = args.batch, args.iters
    first = rank * B                                                 # disjoint episodes per rank
    cpu = synth.make_batch(B, first_episode=first, seed=SEED, **SHAPE)
    a_r, a_t = synth.make_actions(B, iters, seed=SEED, first_episode=first)
    roll = DeviceRollout(cpu, a_r, a_t, dev, iters)
    for _ in range(max(args.warmup, 3)):
        roll.run()
    torch.cuda.synchronize(dev)
    # what one stream, launched eagerly, leaves behind: the timed (graph, possibly two-stream) rollout must leave the same bits
    want = [t.clone() for t in (roll.pose, roll.rew, roll.dist, roll.obs2d, roll.obs3d)]
    roll.two_streams = args.streams == 2

    sampler = ClockSampler(local)
    # ---- timed region: the rollout captured ONCE into a CUDA graph (kernels, the memset and the two tiny torch
    # ops of the prepare step; PDL edges and the cluster launch included) and replayed K times
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        roll.run(count_visible=False)
    torch.cuda.current_stream().wait_stream(side)
    launches0 = _lib.launch_count()
    with torch.cuda.graph(graph):
        roll.run(count_visible=False)     # M_vis is counted in the instrumented eager pass below
    launches_per_step = _lib.launch_count() - launches0
    for _ in range(max(args.warmup, 3)):
        graph.replay()
    torch.cuda.synchronize(dev)
    for name, a, b_ in zip(("pose", "reward", "distance", "obs2d", "obs3d"), want,
                           (roll.pose, roll.rew, roll.dist, roll.obs2d, roll.obs3d)):
        if not torch.equal(a, b_):
            raise SystemExit(f"bench: the timed rollout's {name} differs from the one-stream eager rollout")
    del want
    sampler.start()
    t_begin = sampler.mark()
    dt = timed(graph.replay, args.steps, dev)
    launches = launches_per_step * args.steps
    # ---- the same rollout launched eagerly with CUDA events around the two observe stages (events cannot
    # be read back from inside a graph): per-kernel durations for the roofline.  Every rollout is queued behind
    # a 3 ms spin kernel, so that its launches are all enqueued before the first one runs: what the events
    # bracket is then the GPU's own time (kernel + launch gap), not the host's enqueue rate.
    roll.run(count_visible=True)          # visible predicted-overlap points per step (M_vis of the roofline)
    torch.cuda.synchronize(dev)
    esteps = max(3, min(args.steps, 50))
    events = [[tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(iters)]
              for _ in range(esteps)]
    for s_no in range(esteps):
        torch.cuda._sleep(6_000_000)
        roll.run(events[s_no], count_visible=False)
    torch.cuda.synchronize(dev)
    # one iteration of the instrumented pass = from one k_project's start event to the next one's
    if iters > 1:
        iter_s = statistics.mean(ev[i][0].elapsed_time(ev[i + 1][0]) for ev in events for i in range(iters - 1)) / 1e3
    else:
        iter_s = statistics.mean(ev[0][0].elapsed_time(ev[0][2]) for ev in events) / 1e3
    dt_eager = iter_s * iters * esteps
    t_end = sampler.mark()
    note = "sampled during the timed regions (graph replay + eager instrumented pass)"
    if t_end - t_begin < 0.5:
        # shorter than a few nvidia-smi periods: keep the identical load running (untimed) until ~0.6 s of
        # samples exist, so the clocks are still read UNDER THIS LOAD
        while time.time() - t_begin < 0.6:
            graph.replay()
            torch.cuda.synchronize(dev)
        t_end = sampler.mark()
        note = "timed regions < 0.5 s: sampled over them plus an identical untimed load that follows"
    clocks = sampler.stop(t_begin, t_end)
    clocks["note"] = note
    dt = cdist.max_over_ranks(dt, dev)
    dt_eager = cdist.max_over_ranks(dt_eager, dev)
    steps_done = B * iters * args.steps * world
    value = steps_done / dt

    # ---- roofline of the two observe kernels, from the events of the timed region; the one that takes
    # longer is reported as "roofline" (the dominant kernel), the other as "roofline_secondary".
    # Algorithmic bytes: observe = 33N + 4*C*M_vis + 12*C*P per episode, split as
    #   k_project      33N (pc, overlap -> obs3d) + 8*C*P (image half of obs2d, carried as TMA traffic)
    #   k_tile_gather (cmr_tile_scatter)           4*C*M_vis (feature rows of the visible points) + 4*C*P
    #                  (projected half of obs2d)
    _, N, C, H, W = roll.dims
    P = H * W
    proj_s = statistics.mean(e[0].elapsed_time(e[1]) for ev in events for e in ev) / 1e3
    scat_s = statistics.mean(e[1].elapsed_time(e[2]) for ev in events for e in ev) / 1e3
    mvis = roll.mvis.sum(dim=1).float().mean().item()                # visible overlap points per launch (whole batch)
    copied = bool(roll.copied.value)
    bytes_proj = 33.0 * N * B + (8.0 * C * P * B if copied else 0.0)
    bytes_scat = 4.0 * C * mvis + 4.0 * C * P * B + (0.0 if copied else 8.0 * C * P * B)
    peak, peak_src = load_peak()
    traffic = load_traffic() or {}

    def roof(name, nbytes, sec):
        return {"kernel": name, "bound": "hbm", "achieved": nbytes / sec / 1e9, "peak": peak, "unit": "GB/s",
                "frac": nbytes / sec / 1e9 / peak, "traffic": traffic.get(name),
                "algorithmic_bytes_per_launch": nbytes, "avg_launch_us": sec * 1e6,
                "share_of_step": sec * iters * esteps / dt_eager, "peak_source": peak_src}

    r_proj, r_scat = roof("k_project", bytes_proj, proj_s), roof("k_tile_gather", bytes_scat, scat_s)
    roofline, roofline2 = (r_proj, r_scat) if proj_s >= scat_s else (r_scat, r_proj)
    roofline["m_vis_per_episode"] = mvis / B
    roofline["observe_frac"] = (bytes_proj + bytes_scat) / (proj_s + scat_s) / 1e9 / peak

    # ---- e2e through the drop-in API from host memory
    e2e = None
    if not args.no_e2e:
        host = HostRollout(cpu, a_r, a_t, dev, iters)
        for _ in range(2):
            host.run()
        k = max(3, min(args.steps, 10))
        dte = cdist.max_over_ranks(timed(host.run, k, dev), dev)
        e2e = {"value": B * iters * k * world / dte, "unit": UNIT, "h2d_bytes_per_step": host.h2d,
               "d2h_bytes_per_step": host.d2h, "ms_per_step": dte / k * 1e3, "steps": k}
        # the same with the feature tensors left where CMR-Agent's feature network puts them (in HBM): reported
        # beside e2e, which uploads them too and is bound by PCIe (377 of its 421 MB per step are features)
        host2 = HostRollout(cpu, a_r, a_t, dev, iters, features_resident=True)
        for _ in range(2):
            host2.run()
        dte2 = cdist.max_over_ranks(timed(host2.run, k, dev), dev)
        e2e["features_resident"] = {"value": B * iters * k * world / dte2, "unit": UNIT, "h2d_bytes_per_step": host2.h2d,
                                    "d2h_bytes_per_step": host2.d2h, "ms_per_step": dte2 / k * 1e3}
        del host2

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only), bounded sample
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ncores = os.cpu_count() or 1
        rate, nt, sec = time_cpu_port(args.cpu_episodes, iters, repeats=8, threads_options=sorted({ncores, 1}))
        cpu_base = {"value": rate, "unit": UNIT, "cores": nt, "kind": "port", "host_cores": ncores,
                    "sample": f"{args.cpu_episodes} episodes x {iters} iterations, median of 8 "
                              f"({sec:.2f} s each), better of 1 and {ncores} threads"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "timing": {"value": "K replays of the rollout captured as one CUDA graph"
                                + (" (two streams: the reward of an iteration runs beside the next iteration's observation; "
                                   "observe and step keep their order)" if args.streams == 2 else " (one stream)")
                                + ", CUDA events, max over ranks; its outputs equal a one-stream eager rollout's bit for bit",
                       "roofline": f"eager pass of {esteps} rollouts, each queued behind a 3 ms spin kernel, CUDA events around "
                                   "the observe stages; share_of_step is relative to an iteration of that pass",
                       "eager_us_per_iteration": iter_s * 1e6},
            "config": {"workload": "kitti_b32x10", "episodes_per_gpu": B, "iterations": iters,
                       "registration_steps_per_bench_step": B * iters * world, "num_pt": N, "image": "160x512",
                       "grid": f"{H}x{W}", "channels": C, "sharding": f"episodes/{world}gpu, no data-path collective", "streams": args.streams,
                       "l2": "per-rollout inputs (%.0f MB) exceed the 126 MB L2" %
                             ((roll.feat.numel() + roll.pc.numel() * 2 + roll.img_feat.numel()) * 4 / 1e6)},
            "roofline": roofline, "roofline_secondary": roofline2, "cpu_baseline": cpu_base, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        run_reference_arm(args, rank)
        return
    rank, world, local = cdist.init_from_env()
    try:
        run_b200_arm(args, rank, world, local)
    finally:
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
