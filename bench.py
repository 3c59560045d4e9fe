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
  e2e    : the same rollout from pinned HOST buffers through the library's rollout session (cmr_session_*, one
           native call per rollout, two rollouts in flight: the upload of rollout k+1 runs under the kernels of
           rollout k).  H2D of every input and D2H of the per-iteration rewards/distances and the final poses are
           inside the timed region.  e2e.python_api: the same through the drop-in module functions.
  roofline: the slower of the two observe stages (k_project | k_tile_gather), timed live with CUDA events on the
           launch stream (eager pass, every rollout queued behind a spin kernel so that the host's enqueue rate does
           not show); observe_frac is the PAIR timed around the product call cmr_observe (programmatic dependent launch
           between the two kernels intact), observe_frac_split the sum of the two separately timed stages.
  cpu_baseline / --impl reference: the reference's own environment.py (oracle/_ref, staged by oracle/make_ref.py) on
           this box's host cores; the oracle's port if the staged reference is absent.
  secondary: the other BASELINE configs (NuScenes-shaped 64 episodes over the GPUs, PointNN front-end, B=1 / B=8
           latency through the Python API, KITTI with every predicted-overlap point in view, the agent's 3-D tower,
           the reference's environment.py on CUDA tensors, the scalar all-reduce on NCCL).
"""
import argparse
import contextlib
import ctypes
import json
import math
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
NUSCENES = dict(num_pt=40960, img_h=160, img_w=320, channels=64, unique=(26000, 34000))
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
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--streams", type=int, default=2, choices=[1, 2],
                    help="2: the reward runs on a second stream beside the next observation (default); 1: one stream")
    return ap.parse_args()


def workload_config(args):
    """The SAME dict in both arms (the driver compares them)."""
    return {"workload": "kitti_b32x10", "episodes_per_gpu": args.batch, "iterations": args.iters,
            "num_pt": SHAPE["num_pt"], "image": "160x512", "grid": "40x128", "channels": SHAPE["channels"],
            "data": "synthetic KITTI-shaped (cmr_agent_b200/synth.py), seed 2023 + episode"}


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


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", 1400.0)), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """dram bytes per launch, {kernel: bytes}, from the committed ncu --set full capture, if any."""
    path = os.path.join(ROOT, "profiles", "dram_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


def pin_to_gpu_numa_node(dev_index):
    """Run this process (and first-touch its pinned buffers) on the NUMA node the GPU hangs off."""
    info = {"numa_node": None, "cpus": None}
    try:
        prop = torch.cuda.get_device_properties(dev_index)
        bus = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node >= 0:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                spec = f.read().strip()
            cpus = set()
            for part in spec.split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            allowed = os.sched_getaffinity(0) & cpus
            if allowed:
                os.sched_setaffinity(0, allowed)
                info["cpus"] = len(allowed)
    except Exception as e:  # pragma: no cover - topology files differ between hosts
        info["error"] = str(e)[:80]
    return info


# ------------------------------------------------------------------------------- CPU baseline
def cpu_env_module():
    """(module, kind): the reference's own environment.py staged as oracle/_ref (or /root/reference in the build
    container) if present, else the oracle's operator-for-operator port."""
    from oracle import reference_loader as rl
    if rl.available():
        try:
            mod = rl.environment()
            # the module's global DEVICE is "cuda if available" (environment.py:11): this arm is the HOST path
            mod.DEVICE = torch.device("cpu")
            return mod, "reference"
        except Exception:
            pass
    from oracle import env_oracle
    return env_oracle, "port"


def cpu_rollout(mod, data, a_r, a_t, iters, cfg):
    """One rollout of the reference path on the CPU: init, to_disentangled, iters x (observation_from_a_pose + step +
    reward) - Test_Agent.py:150-170 with Train_Agent.py's reward call.  Returns registration steps done."""
    pose, target = mod.init(data)
    mod.to_disentangled(target, data["pc"])
    prev = None
    for it in range(iters):
        mod.observation_from_a_pose(data, pose)
        pose = mod.step(a_r[it], a_t[it], pose, cfg)
        _, prev = mod.reward(pose, data, prev)
    return data["pc"].shape[0] * iters


def pick_threads(mod, data, a_r, a_t, iters, cfg, ncores):
    cand = {}
    for nt in sorted({ncores, 1}, reverse=True):   # multi-threading hurts the small ops on some hosts: take the better
        torch.set_num_threads(nt)
        cpu_rollout(mod, data, a_r, a_t, iters, cfg)
        t0 = time.perf_counter()
        cpu_rollout(mod, data, a_r, a_t, iters, cfg)
        cand[nt] = time.perf_counter() - t0
    nt = min(cand, key=cand.get)
    torch.set_num_threads(nt)
    return nt


def time_cpu_baseline(episodes, iters, repeats):
    mod, kind = cpu_env_module()
    data = synth.make_batch(episodes, seed=SEED, **SHAPE)
    a_r, a_t = synth.make_actions(episodes, iters, seed=SEED)
    cfg = synth.StepConfig()
    ncores = os.cpu_count() or 1
    nt = pick_threads(mod, data, a_r, a_t, iters, cfg, ncores)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        n = cpu_rollout(mod, data, a_r, a_t, iters, cfg)
        times.append(time.perf_counter() - t0)
    sec = statistics.median(times)
    return {"value": n / sec, "unit": UNIT, "cores": nt, "kind": kind, "host_cores": ncores,
            "sample": f"{episodes} episodes x {iters} iterations, median of {repeats} ({sec:.2f} s each), "
                      f"better of 1 and {ncores} threads"}


def run_reference_arm(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores."""
    if rank != 0:
        return
    mod, kind = cpu_env_module()
    ncores = os.cpu_count() or 1
    episodes, iters = args.cpu_episodes, args.iters
    data = synth.make_batch(episodes, seed=SEED, **SHAPE)
    a_r, a_t = synth.make_actions(episodes, iters, seed=SEED)
    cfg = synth.StepConfig()
    nt = pick_threads(mod, data, a_r, a_t, iters, cfg, ncores)
    for _ in range(max(args.warmup - 2, 0)):
        cpu_rollout(mod, data, a_r, a_t, iters, cfg)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        done += cpu_rollout(mod, data, a_r, a_t, iters, cfg)
    dt = time.perf_counter() - t0
    value = done / dt
    sample = f"{episodes} episodes x {iters} iterations per step (bounded sample of the 32x10 workload)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nt, "kind": kind, "sample": sample,
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
        self.pose0 = None                      # start pose other than identity (secondary runs)
        self.mvis = torch.zeros(iters, B, dtype=torch.int32, device=dev)
        self.rew = torch.empty(iters, B, device=dev)
        self.dist = torch.empty(iters, B, device=dev)
        self.mean = torch.empty(B, 3, device=dev)
        self.copied = ctypes.c_int(0)
        self.two_streams = False
        self.side = torch.cuda.Stream(dev)

    def _reward(self, it, stream):
        """environment.reward as shipped (:263-302): the distance ignores the pose, so the product computes it once per
        batch (cmr_reward, iteration 0) and compares the memoised value afterwards (cmr_reward_compare)."""
        L, p = self.lib, self.lib.ptr
        B, N = self.dims[0], self.dims[1]
        prev = p(self.dist[it - 1]) if it else None
        if it == 0:
            L.call("cmr_reward", p(self.target), p(self.pc), p(self.mask), p(self.mean), p(self.pose), prev, 0, B, N,
                   p(self.scratch), p(self.rew[it]), p(self.dist[it]), stream)
        else:
            L.call("cmr_reward_compare", p(self.dist[0]), prev, B, p(self.rew[it]), p(self.dist[it]), stream)

    def run(self, events=None, count_visible=True, pair_events=None):
        """events: per iteration (e0, e1, e2) recorded before cmr_project, between the two observe kernels
        and after cmr_tile_scatter, on the launch stream.  pair_events: per iteration (e0, e1) around the product
        call cmr_observe.  count_visible: also count the visible predicted-overlap points per episode (M_vis of the
        roofline; one memset + a few atomics per step)."""
        L, p, st = self.lib, self.lib.ptr, self.lib.stream()
        B, N, C, H, W = self.dims
        L.call("cmr_cloud_mean", p(self.pc), B, N, p(self.mean), st)    # environment.py:46 - once per episode
        L.call("cmr_episode_prepare", p(self.overlap), p(self.feat), B, N, C, p(self.ws), st)
        self.pose.copy_(self.eye if self.pose0 is None else self.pose0)   # env.init
        for it in range(self.iters):
            if events is not None:
                events[it][0].record()
            mv = p(self.mvis[it]) if count_visible else None
            if events is None and pair_events is None and self.two_streams:
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
                self._reward(it, sst)
                rewarded = torch.cuda.Event()
                rewarded.record(side)
                if it == self.iters - 1:
                    main.wait_stream(side)
                continue
            if events is None:   # the product call
                if pair_events is not None:
                    pair_events[it][0].record()
                L.call("cmr_observe", p(self.pc), p(self.overlap), p(self.img_feat), p(self.K), p(self.pose), p(self.mean),
                       p(self.ws), B, N, C, H, W, p(self.obs2d), p(self.obs3d), None, mv, st)
                if pair_events is not None:
                    pair_events[it][1].record()
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
            self._reward(it, st)


def capture(roll, dev):
    """Warm up on a side stream, then capture one rollout as a CUDA graph."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        roll.run(count_visible=False)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        roll.run(count_visible=False)
    return graph


class HostRollout:
    """The same rollout through the drop-in API from pinned host memory (e2e.python_api)."""

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


class SessionRollouts:
    """e2e: K rollouts from pinned host buffers through the native rollout session, `depth` in flight."""

    def __init__(self, cpu, a_r, a_t, dev, iters, depth=2, features_resident=False):
        from cmr_agent_b200 import session
        B, _, N = cpu["pc"].shape
        C = cpu["pc_geo_feat"].shape[1]
        H, W = cpu["img"].shape[2] // 4, cpu["img"].shape[3] // 4
        self.depth = depth
        self.ses = session.RolloutSession(B, N, C, H, W, iters, synth.StepConfig(), depth=depth,
                                          features_resident=features_resident, device=dev)
        self.data = {k: (v.pin_memory() if (isinstance(v, torch.Tensor) and k != "img") else v) for k, v in cpu.items()}
        if features_resident:
            self.data["pc_geo_feat"], self.data["img_geo_feat"] = cpu["pc_geo_feat"].to(dev), cpu["img_geo_feat"].to(dev)
        self.a_r, self.a_t = a_r.pin_memory(), a_t.pin_memory()
        self.d2h = iters * B * 4 * 2 + 2 * B * 16 * 4
        self.last = None

    def run(self, k):
        """k rollouts, pipelined; every rollout's results are read on the host."""
        tickets = []
        for i in range(k):
            tickets.append(self.ses.submit(self.data, self.a_r, self.a_t))
            if i >= self.depth - 1:
                self.last = self.ses.wait(tickets[i - (self.depth - 1)])
        for t in tickets[max(0, k - (self.depth - 1)):]:
            self.last = self.ses.wait(t)


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


def timed_wall(fn, dev):
    """Wall clock between two device synchronisations (host-driven pipelines: the host's waits are part of it)."""
    cdist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    cdist.barrier()
    return dt


def gather_floats(value, dev, world):
    """[value of rank 0, ..., value of rank world-1]"""
    if world == 1:
        return [float(value)]
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    out = [torch.zeros_like(t) for _ in range(world)]
    torch.distributed.all_gather(out, t)
    return [float(x.item()) for x in out]


def observe_pair_bytes(roll, mvis_per_launch):
    B, N, C, H, W = roll.dims
    return 33.0 * N * B + 4.0 * C * mvis_per_launch + 12.0 * C * H * W * B


def measure_observe(roll, dev, esteps, iters):
    """Eager instrumented passes (every rollout queued behind a spin kernel): per-stage times with an event between the
    two kernels, and the pair timed around the product call."""
    roll.run(count_visible=True)
    torch.cuda.synchronize(dev)
    mvis = roll.mvis.sum(dim=1).float().mean().item()
    events = [[tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(iters)] for _ in range(esteps)]
    for s_no in range(esteps):
        torch.cuda._sleep(6_000_000)
        roll.run(events[s_no], count_visible=False)
    torch.cuda.synchronize(dev)
    pair = [[tuple(torch.cuda.Event(enable_timing=True) for _ in range(2)) for _ in range(iters)] for _ in range(esteps)]
    for s_no in range(esteps):
        torch.cuda._sleep(6_000_000)
        roll.run(None, count_visible=False, pair_events=pair[s_no])
    torch.cuda.synchronize(dev)
    proj_s = statistics.mean(e[0].elapsed_time(e[1]) for ev in events for e in ev) / 1e3
    scat_s = statistics.mean(e[1].elapsed_time(e[2]) for ev in events for e in ev) / 1e3
    pair_s = statistics.mean(e[0].elapsed_time(e[1]) for ev in pair for e in ev) / 1e3
    if iters > 1:
        iter_s = statistics.mean(ev[i][0].elapsed_time(ev[i + 1][0]) for ev in events for i in range(iters - 1)) / 1e3
    else:
        iter_s = statistics.mean(ev[0][0].elapsed_time(ev[0][2]) for ev in events) / 1e3
    return dict(proj_s=proj_s, scat_s=scat_s, pair_s=pair_s, iter_s=iter_s, mvis=mvis)


# ------------------------------------------------------------------------------------- secondary
def api_latency(dev, B, iters=10, reps=20):
    """Wall-clock microseconds per call of the Python drop-in functions (eager, one stream, features on the device):
    what the reference's interactive loop pays per iteration at this batch size."""
    from cmr_agent_b200 import environment as env
    cpu = synth.make_batch(B, seed=SEED + 5, **SHAPE)
    data = dict(cpu)
    for k in ("pc", "pc_overlap_pred", "pc_geo_feat", "img_geo_feat"):
        data[k] = cpu[k].to(dev)
    cfg = synth.StepConfig(device=dev)
    a_r, a_t = synth.make_actions(B, iters, seed=3)
    a_r, a_t = a_r.to(dev), a_t.to(dev)
    pose, _ = env.init(data)
    env.observation_from_a_pose(data, pose)
    _, prev = env.reward(pose, data, None)
    torch.cuda.synchronize(dev)

    eye = pose.clone()

    def loop(fn):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(reps):
            pose.copy_(eye)                      # every repetition is a rollout from the start pose, like the timed one
            for it in range(iters):
                fn(it)
        torch.cuda.synchronize(dev)
        return (time.perf_counter() - t0) / (reps * iters) * 1e6

    out = {"batch": B}
    out["observe_api_us"] = loop(lambda it: env.observation_from_a_pose(data, pose))
    out["step_api_us"] = loop(lambda it: env.step(a_r[it], a_t[it], pose, cfg))
    out["reward_api_us"] = loop(lambda it: env.reward(pose, data, prev))

    def full(it):
        env.observation_from_a_pose(data, pose)
        env.step(a_r[it], a_t[it], pose, cfg)
        env.reward(pose, data, prev)
    out["api_us_per_iteration"] = loop(full)
    out["iterate_api_us"] = loop(lambda it: env.iterate(data, pose, a_r[it], a_t[it], cfg, prev_distance=prev))
    out["steps_per_s"] = B / out["api_us_per_iteration"] * 1e6
    env.clear_cache()
    return out


def frontend_bench(dev, batch=128, with_cpu=True):
    """BASELINE config 4: FPS 40960 -> 1280 + kNN k = 64 + grouping, batch 128; beside it the reference's own
    pointnet_util on the host for ONE cloud, scaled by the batch (square_distance would materialise 80 GB batched)."""
    from cmr_agent_b200 import pointnet_util as pn
    N, S, K = 40960, 1280, 64
    g = torch.Generator().manual_seed(1)
    base = synth.make_cloud_batch(min(batch, 8), num_pt=N, seed=SEED)
    xyz = base.repeat((batch + base.shape[0] - 1) // base.shape[0], 1, 1)[:batch].contiguous()
    xyz += torch.randn(batch, 1, 3, generator=g) * 0.01
    xyz_d = xyz.to(dev)
    start = torch.randint(0, N, (batch,), generator=g)
    start_d = start.to(dev)
    feats = torch.randn(batch, N, 3, generator=g).to(dev)

    def t_ms(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(dev)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in ev:
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize(dev)
        return statistics.median(a.elapsed_time(b) for a, b in ev)

    out = {"batch": batch, "N": N, "npoint": S, "k": K}
    out["fps_ms"] = t_ms(lambda: pn.farthest_point_sample_from(xyz_d, S, start_d))
    fps = pn.farthest_point_sample_from(xyz_d, S, start_d)
    new_xyz = pn.index_points(xyz_d, fps)
    out["knn_ms"] = t_ms(lambda: pn.knn_point(K, xyz_d, new_xyz))
    idx = pn.knn_point(K, xyz_d, new_xyz)
    out["ball_ms"] = t_ms(lambda: pn.query_ball_point(1.0, K, xyz_d, new_xyz))
    out["group_ms"] = t_ms(lambda: pn.group_points(xyz_d, feats, new_xyz, idx))
    out["total_ms"] = out["fps_ms"] + out["knn_ms"] + out["group_ms"]
    if with_cpu:
        from oracle import reference_loader as rl
        if rl.available():
            ref = rl.pointnet_util()
            kind = "reference"
        else:
            from oracle import pointnet_oracle as ref
            kind = "port"
        one = xyz[:1]
        best = None
        for nt in sorted({os.cpu_count() or 1, 1}):
            torch.set_num_threads(nt)
            torch.manual_seed(0)
            t0 = time.perf_counter()
            c = ref.farthest_point_sample(one, S)
            t_f = time.perf_counter() - t0
            nx = ref.index_points(one, c)
            t0 = time.perf_counter()
            d = ref.square_distance(nx, one)
            kn = d.argsort()[:, :, :K]
            t_k = time.perf_counter() - t0
            t0 = time.perf_counter()
            ref.index_points(one, kn)
            t_g = time.perf_counter() - t0
            if best is None or t_f + t_k + t_g < sum(best[:3]):
                best = (t_f, t_k, t_g, nt)
        out["cpu_reference"] = {"kind": kind, "threads": best[3], "fps_ms_per_cloud": best[0] * 1e3,
                                "knn_ms_per_cloud": best[1] * 1e3, "group_ms_per_cloud": best[2] * 1e3,
                                "total_ms_scaled_to_batch": sum(best[:3]) * 1e3 * batch,
                                "note": "one cloud timed on the host, x batch (the batched call would materialise 80 GB)"}
        out["speedup_vs_cpu_reference"] = out["cpu_reference"]["total_ms_scaled_to_batch"] / out["total_ms"]
    return out


def gpu_torch_baseline(dev, cpu, a_r, a_t, iters):
    """The reference's environment.py itself on CUDA tensors (what a CMR-Agent user runs today), same rollout."""
    from oracle import reference_loader as rl
    if not rl.available():
        return {"unavailable": "oracle/_ref not staged (run oracle/make_ref.py where /root/reference exists)"}
    ref = rl.environment()
    old_device = ref.DEVICE
    ref.DEVICE = dev                                            # environment.py:11, as on the authors' GPU box
    data = dict(cpu)
    for k in ("pc", "pc_overlap_pred", "pc_geo_feat", "img_geo_feat"):
        data[k] = cpu[k].to(dev)
    cfg = synth.StepConfig(device=dev)
    ar, at = a_r.to(dev), a_t.to(dev)

    def rollout():
        return cpu_rollout(ref, data, ar, at, iters, cfg)
    try:
        with torch.no_grad():
            rollout()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            n = 0
            for _ in range(2):
                n += rollout()
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
    finally:
        ref.DEVICE = old_device
    return {"value": n / dt, "unit": UNIT, "ms_per_rollout": dt / 2 * 1e3, "kind": "reference environment.py on CUDA tensors "
            "(torch " + torch.__version__ + ", torch_scatter shim), features resident, 2 rollouts"}


def agent_loop_bench(dev, batches=(1, 8, 32), iters=10, reps=3):
    """BASELINE configs[0] / SURVEY 8d: the reference's own Test_Agent.py:150-170 loop - observation, the UNCHANGED
    random-init CMRAgent (eval), deterministic actions, step - timed per iteration (wall clock, synchronised at both
    ends; the loop's host work is part of what a user waits for) three ways: the reference's environment.py on CUDA
    tensors, the drop-in environment, the drop-in environment with the agent's 3-D tower on tcgen05
    (agent_tower.accelerate_agent).  Inputs resident on the device, as after the reference's feature networks."""
    from oracle import reference_loader as rl
    if not rl.available():
        return {"unavailable": "oracle/_ref not staged (run oracle/make_ref.py where /root/reference exists)"}
    import cmr_agent_b200
    from cmr_agent_b200 import agent_tower, environment as drop_env
    rl.put_on_path()
    cmr_agent_b200.install()
    out = {"loop": "Test_Agent.py:150-170, %d iterations, median of %d" % (iters, reps), "batches": []}
    try:
        from config import KittiConfiguration
        from models import CMRAgent
        ref_env = rl.environment()
        old_device = ref_env.DEVICE
        ref_env.DEVICE = dev
        config = KittiConfiguration()
        config.action_num = iters

        def loop(env, agent, data):
            pose_source, pose_target = env.init(data)
            pose_target = env.to_disentangled(pose_target, data["pc"])
            for _ in range(config.action_num):
                s2, s3 = env.observation_from_a_pose(data, pose_source)
                lr, lt, _ = agent(s2, s3)
                a_r, a_t = agent.action_from_logits(lr, lt, deterministic=True)
                pose_source = env.step(a_r, a_t, pose_source, config)
            return pose_source

        def timed_loop(env, agent, data):
            with torch.no_grad():
                loop(env, agent, data)
                ts = []
                for _ in range(reps):
                    torch.cuda.synchronize(dev)
                    t0 = time.perf_counter()
                    loop(env, agent, data)
                    torch.cuda.synchronize(dev)
                    ts.append(time.perf_counter() - t0)
            return statistics.median(ts) / iters * 1e3

        try:
            for B in batches:
                cpu = synth.make_batch(B, first_episode=3, seed=SEED, **SHAPE)
                data = dict(cpu)
                for k in ("pc", "pc_overlap_pred", "pc_geo_feat", "img_geo_feat"):
                    data[k] = cpu[k].to(dev)
                torch.manual_seed(SEED)
                agent = CMRAgent(config).to(dev).eval()
                rec = {"batch": B, "reference_ms_per_iteration": timed_loop(ref_env, agent, data),
                       "drop_in_env_ms_per_iteration": timed_loop(drop_env, agent, data)}
                agent_tower.accelerate_agent(agent)
                rec["drop_in_env_and_tower_ms_per_iteration"] = timed_loop(drop_env, agent, data)
                # the same loop captured once as a CUDA graph (environment.capture_rollout with the agent as policy)
                torch.distributions.Distribution.set_default_validate_args(False)   # its check reads the host
                try:
                    with torch.no_grad():
                        roll = drop_env.capture_rollout(
                            data, config, with_reward=False, iters=iters, reusable=True,   # the per-batch preparation is in the graph
                            policy=lambda s2, s3: agent.action_from_logits(*agent(s2, s3)[:2], deterministic=True))
                        roll.replay()
                        ts = []
                        for _ in range(reps):
                            torch.cuda.synchronize(dev)
                            t0 = time.perf_counter()
                            roll.replay()
                            torch.cuda.synchronize(dev)
                            ts.append(time.perf_counter() - t0)
                    rec["captured_ms_per_iteration"] = statistics.median(ts) / iters * 1e3
                    del roll
                finally:
                    torch.distributions.Distribution.set_default_validate_args(True)
                best = min(rec["drop_in_env_and_tower_ms_per_iteration"], rec["captured_ms_per_iteration"])
                rec["speedup"] = rec["reference_ms_per_iteration"] / best
                rec["steps_per_s"] = B / best * 1e3
                rec["reference_steps_per_s"] = B / rec["reference_ms_per_iteration"] * 1e3
                out["batches"].append(rec)
                del agent, data
                torch.cuda.empty_cache()
        finally:
            ref_env.DEVICE = old_device
    finally:
        cmr_agent_b200.uninstall()
    return out


def tower_bench(dev, B, N, bf16_peak):
    from cmr_agent_b200 import agent_tower
    from oracle import tower_oracle as to
    g = torch.Generator().manual_seed(3)
    obs3d = torch.cat([(torch.rand(B, 3, N, generator=g) - 0.5) * 160, (torch.rand(B, 2, N, generator=g) < 0.3).float()], 1).to(dev)
    states = [to.make_state(50 + i, ci, co) for i, (ci, co) in enumerate(to.TOWER)]
    tower = agent_tower.Tower3D(states, dev)
    for _ in range(3):
        tower(obs3d)
    torch.cuda.synchronize(dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in ev:
        a.record()
        tower(obs3d)
        b.record()
    torch.cuda.synchronize(dev)
    t = statistics.median(a.elapsed_time(b) for a, b in ev) / 1e3
    macs = 5 * 5 + 2 * 5 * 64 + 2 * (64 * 128 + 128 * 64 + 64 * 64) + (64 * 128 + 128 * 128)
    tensor_tf = 3 * 2.0 * (macs - 665) * B * N / t / 1e12
    return {"B": B, "N": N, "ms": t * 1e3, "tflops_logical": 2.0 * macs * B * N / t / 1e12, "tflops_tensor_pipe": tensor_tf,
            "frac_of_bf16_sustained": tensor_tf / bf16_peak, "passes": "3 x fp16 (split operands), fp32 accumulate in TMEM"}


def _median_ms(fn, dev, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize(dev)
    return statistics.median(a.elapsed_time(b) for a, b in ev)


def cost_volume_bench(dev, peak, nlabel=9):
    """IterModel's cost-volume warp (models/IterModel.py:272-351) at the reference's sizes - one KITTI cloud,
    nlabel^3 = 729 candidate poses around the ground truth, 68 x 5120 floats out per pose.  The whole call (prepare + warp)."""
    from cmr_agent_b200 import cost_volume
    data = synth.make_batch(1, seed=SEED + 3, **SHAPE)
    g = torch.Generator().manual_seed(SEED + 4)
    scores = torch.rand(1, SHAPE["num_pt"], generator=g)
    base = torch.linspace(-(nlabel - 1) / 2, (nlabel - 1) / 2, nlabel)
    gt = data["P"][0, 0:3, :]
    poses = []
    for ry in base * (0.3 / (nlabel - 1)):
        c, s_ = math.cos(float(ry)), math.sin(float(ry))
        R = torch.tensor([[c, 0.0, s_], [0.0, 1.0, 0.0], [-s_, 0.0, c]])
        for tx in base * (4.0 / (nlabel - 1)):
            for tz in base * (4.0 / (nlabel - 1)):
                poses.append(torch.cat([R @ gt[:, 0:3], R @ gt[:, 3:4] + torch.tensor([[float(tx)], [0.0], [float(tz)]])], dim=1))
    poses = torch.stack(poses).unsqueeze(0).contiguous()
    H, W = SHAPE["img_h"] // 4, SHAPE["img_w"] // 4
    dargs = (data["pc"].to(dev), data["pc_overlap_pred"][0].to(dev), poses.to(dev), data["K"], data["pc_geo_feat"].to(dev),
             scores.to(dev), H, W)
    ms = _median_ms(lambda: cost_volume.warp(*dargs), dev)
    wf, occ = cost_volume.warp(*dargs)
    out_bytes = poses.shape[1] * 68.0 * H * W * 4
    return {"poses": int(poses.shape[1]), "masked_points": int(data["pc_overlap_pred"][0].sum()),
            "occupied_pixels_per_pose": float((occ > 0).sum()) / poses.shape[1], "ms": ms, "output_gbs": out_bytes / ms / 1e6,
            "output_frac_of_hbm": out_bytes / ms / 1e6 / peak}


def sample_bench(dev, peak, cpu, B):
    """environment.sample_image_features (north_star's point-side bilinear gather, an extension - SURVEY D1) on this rank's
    episodes at the ground-truth pose.  Bytes per episode: 12N in, 4CN + N out (the feature map stays in L2)."""
    from cmr_agent_b200 import environment as env
    data = dict(cpu)
    for k in ("pc", "pc_overlap_pred", "pc_geo_feat", "img_geo_feat"):
        data[k] = cpu[k].to(dev)
    pose = cpu["P"].to(dev).clone()
    env.to_disentangled(pose, data["pc"])
    feats, cam = env.sample_image_features(data, pose)
    ms = _median_ms(lambda: env.sample_image_features(data, pose), dev, warm=3, reps=10)
    N, C = SHAPE["num_pt"], 64
    byts = (12.0 * N + 4.0 * C * N + N) * B
    return {"batch": B, "in_frustum_frac": float(cam.float().mean()), "us": ms * 1e3, "gbs": byts / ms / 1e6,
            "frac_of_hbm": byts / ms / 1e6 / peak}


def secondary_block(args, rank, world, local, dev, peak, bf16_peak, cpu, a_r, a_t):
    from cmr_agent_b200 import _lib
    sec = {}
    iters = args.iters
    # ---- config 3: NuScenes-shaped, 64 episodes in total, split over the GPUs (strong scaling)
    lo, hi = cdist.shard_range(64, rank, world)
    if hi > lo:
        ncpu = synth.make_batch(hi - lo, first_episode=1000 + lo, seed=SEED, **NUSCENES)
        nar, nat = synth.make_actions(hi - lo, iters, seed=SEED + 1, first_episode=lo)
        nroll = DeviceRollout(ncpu, nar, nat, dev, iters)
        for _ in range(3):
            nroll.run(count_visible=False)
        nroll.two_streams = args.streams == 2
        graph = capture(nroll, dev)
        for _ in range(3):
            graph.replay()
        k = max(args.steps, 20)
        dt = cdist.max_over_ranks(timed(graph.replay, k, dev), dev)
        m = measure_observe(nroll, dev, 5, iters)
        frac = observe_pair_bytes(nroll, m["mvis"]) / m["pair_s"] / 1e9 / peak
        fr = gather_floats(frac, dev, world)
        if rank == 0:
            sec["nuscenes_64_episodes"] = {"value": 64 * iters * k / dt, "unit": UNIT, "scaling": "strong",
                                           "episodes_per_gpu": hi - lo, "grid": "40x80", "ms_per_rollout": dt / k * 1e3,
                                           "observe_frac_per_rank": fr, "m_vis_per_episode": m["mvis"] / (hi - lo)}
        del nroll, graph
    # ---- KITTI with the poses at the ground truth: every predicted-overlap point that the camera sees is in view
    groll = DeviceRollout(cpu, a_r * 0 + 5, a_t * 0 + 5, dev, iters)       # bin 5 = the zero step: the pose stays there
    from cmr_agent_b200 import environment as env
    target = cpu["P"].to(dev).clone()
    env.to_disentangled(target, groll.pc)
    groll.pose0 = target
    for _ in range(3):
        groll.run(count_visible=False)
    m = measure_observe(groll, dev, 5, iters)
    B = args.batch
    b_pair = observe_pair_bytes(groll, m["mvis"])
    gt = {"m_vis_per_episode": m["mvis"] / B, "observe_pair_us": m["pair_s"] * 1e6,
          "observe_frac": b_pair / m["pair_s"] / 1e9 / peak,
          "project_us": m["proj_s"] * 1e6, "gather_us": m["scat_s"] * 1e6,
          "gather_frac": (4.0 * 64 * m["mvis"] + 12.0 * 64 * 5120 * B) / m["scat_s"] / 1e9 / peak}
    fr = gather_floats(gt["observe_frac"], dev, world)
    if rank == 0:
        gt["observe_frac_per_rank"] = fr
        sec["kitti_at_ground_truth_pose"] = gt
    del groll
    torch.cuda.empty_cache()
    # ---- the one collective of the design on NCCL: recall / RRE / RTE sums of this rank's episodes (dist.MetricSums)
    ms = cdist.MetricSums(device=dev)
    g = torch.Generator().manual_seed(SEED + rank)
    ms.add(torch.rand(args.batch, generator=g) * 20, torch.rand(args.batch, generator=g) * 10, reward=torch.zeros(args.batch))
    ms.all_reduce()
    if rank == 0:
        s = ms.summary()
        sec["metric_all_reduce"] = {"backend": "nccl" if world > 1 else "none (1 rank)", "episodes": s["episodes"],
                                    "recall": s["recall"]}
    if rank == 0 and world == 1:
        def comparator(name, fn):
            """The two entries that run the REFERENCE's code (its environment / its agent from oracle/_ref) beside ours:
            informational, and not allowed to take the bench line down with them."""
            try:
                with contextlib.redirect_stdout(sys.stderr):     # the reference's config class prints a banner
                    sec[name] = fn()
            except Exception as e:                               # noqa: BLE001
                sec[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        # ---- config 1 / training batch: latency through the Python API at the reference's own batch sizes
        sec["api_latency"] = [api_latency(dev, b) for b in (1, 8, 32)]
        # ---- config 4
        sec["frontend_b128"] = frontend_bench(dev, 128, with_cpu=not args.no_cpu_baseline)
        # ---- the agent's 3-D tower
        sec["tower3d"] = tower_bench(dev, args.batch, SHAPE["num_pt"], bf16_peak)
        # ---- the cost-volume warp and the point-side bilinear gather (north_star)
        sec["cost_volume_729_poses"] = cost_volume_bench(dev, peak)
        sec["sample_image_features"] = sample_bench(dev, peak, cpu, args.batch)
        torch.cuda.empty_cache()
        # ---- the reference's own environment.py on CUDA tensors
        comparator("gpu_torch_baseline", lambda: gpu_torch_baseline(dev, cpu, a_r, a_t, iters))
        # ---- configs[0]: the reference's Test_Agent loop with its unchanged CMRAgent in it
        comparator("test_agent_loop", lambda: agent_loop_bench(dev, iters=iters))
    return sec


def run_b200_arm(args, rank, world, local):
    from cmr_agent_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    _lib.load()
    numa = pin_to_gpu_numa_node(local)
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
    launches0 = _lib.launch_count()
    graph = capture(roll, dev)
    launches_per_step = (_lib.launch_count() - launches0) // 2
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
    # ---- the same load for >= 0.6 s, so that nvidia-smi (100 ms period) samples clocks INSIDE a timed region
    k_long = max(args.steps, int(0.6 / max(dt / args.steps, 1e-6)) + 1)
    dt_long = timed(graph.replay, k_long, dev)
    t_end = sampler.mark()
    clocks = sampler.stop(t_begin, t_end)
    clocks["note"] = f"sampled during the K timed replays and the {k_long} replays of `sustained` that follow"
    # ---- per-stage times for the roofline
    esteps = max(3, min(args.steps, 50))
    m = measure_observe(roll, dev, esteps, iters)
    dt_eager = m["iter_s"] * iters * esteps
    dt = cdist.max_over_ranks(dt, dev)
    dt_long = cdist.max_over_ranks(dt_long, dev)
    steps_done = B * iters * args.steps * world
    value = steps_done / dt

    # ---- roofline of the two observe kernels; the one that takes longer is "roofline" (the dominant kernel).
    # Algorithmic bytes: observe = 33N + 4*C*M_vis + 12*C*P per episode, split as
    #   k_project      33N (pc, overlap -> obs3d)
    #   k_tile_gather  4*C*M_vis (feature rows of the visible points) + 12*C*P (image half in and out, projected half)
    _, N, C, H, W = roll.dims
    P = H * W
    proj_s, scat_s, mvis = m["proj_s"], m["scat_s"], m["mvis"]
    copied = bool(roll.copied.value)
    bytes_proj = 33.0 * N * B + (8.0 * C * P * B if copied else 0.0)
    bytes_scat = 4.0 * C * mvis + 4.0 * C * P * B + (0.0 if copied else 8.0 * C * P * B)
    peak, bf16_peak, peak_src = load_peaks()
    traffic = load_traffic() or {}

    def roof(name, nbytes, sec):
        return {"kernel": name, "bound": "hbm", "achieved": nbytes / sec / 1e9, "peak": peak, "unit": "GB/s",
                "frac": nbytes / sec / 1e9 / peak, "traffic": traffic.get(name),
                "algorithmic_bytes_per_launch": nbytes, "avg_launch_us": sec * 1e6,
                "share_of_step": sec * iters * esteps / dt_eager, "peak_source": peak_src}

    r_proj, r_scat = roof("k_project", bytes_proj, proj_s), roof("k_tile_gather", bytes_scat, scat_s)
    roofline, roofline2 = (r_proj, r_scat) if proj_s >= scat_s else (r_scat, r_proj)
    roofline["m_vis_per_episode"] = mvis / B
    roofline["observe_pair_us"] = m["pair_s"] * 1e6
    roofline["observe_frac"] = (bytes_proj + bytes_scat) / m["pair_s"] / 1e9 / peak
    roofline["observe_frac_split"] = (bytes_proj + bytes_scat) / (proj_s + scat_s) / 1e9 / peak

    # ---- e2e from host memory
    e2e = None
    if not args.no_e2e:
        k = max(3, min(args.steps, 10))
        ses = SessionRollouts(cpu, a_r, a_t, dev, iters, depth=2)
        ses.run(3)
        dte = cdist.max_over_ranks(timed_wall(lambda: ses.run(k), dev), dev)
        st = ses.ses.stats()
        rates = gather_floats(st["h2d_gbs"], dev, world)
        e2e = {"value": B * iters * k * world / dte, "unit": UNIT, "h2d_bytes_per_step": int(st["h2d_bytes_per_rollout"]),
               "d2h_bytes_per_step": ses.d2h, "ms_per_step": dte / k * 1e3, "steps": k,
               "path": "cmr_session_submit/wait (native rollout session), 2 rollouts in flight, pinned host buffers",
               "h2d_gbs_per_rank": rates, "numa": numa}
        ses.ses.close()
        del ses
        ses2 = SessionRollouts(cpu, a_r, a_t, dev, iters, depth=2, features_resident=True)
        ses2.run(3)
        k2 = max(k, 20)
        dte2 = cdist.max_over_ranks(timed_wall(lambda: ses2.run(k2), dev), dev)
        st2 = ses2.ses.stats()
        e2e["features_resident"] = {"value": B * iters * k2 * world / dte2, "unit": UNIT,
                                    "h2d_bytes_per_step": int(st2["h2d_bytes_per_rollout"]), "d2h_bytes_per_step": ses2.d2h,
                                    "ms_per_step": dte2 / k2 * 1e3, "h2d_gbs_per_rank": gather_floats(st2["h2d_gbs"], dev, world)}
        ses2.ses.close()
        del ses2
        # the same through the Python drop-in functions (one stream, uploads and kernels in series)
        host = HostRollout(cpu, a_r, a_t, dev, iters)
        for _ in range(2):
            host.run()
        dtp = cdist.max_over_ranks(timed(host.run, k, dev), dev)
        host2 = HostRollout(cpu, a_r, a_t, dev, iters, features_resident=True)
        for _ in range(2):
            host2.run()
        dtp2 = cdist.max_over_ranks(timed(host2.run, k, dev), dev)
        e2e["python_api"] = {"value": B * iters * k * world / dtp, "h2d_bytes_per_step": host.h2d, "ms_per_step": dtp / k * 1e3,
                             "features_resident": {"value": B * iters * k * world / dtp2, "h2d_bytes_per_step": host2.h2d,
                                                   "ms_per_step": dtp2 / k * 1e3}}
        del host, host2
        torch.cuda.empty_cache()

    secondary = None
    if not args.no_secondary:
        secondary = secondary_block(args, rank, world, local, dev, peak, bf16_peak, cpu, a_r, a_t)

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only), bounded sample
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = time_cpu_baseline(args.cpu_episodes, iters, repeats=5)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "run": {"registration_steps_per_bench_step": B * iters * world, "streams": args.streams,
                    "sharding": f"episodes/{world}gpu, no data-path collective",
                    "l2": "per-rollout inputs (%.0f MB) exceed the 126 MB L2" %
                          ((roll.feat.numel() + roll.pc.numel() * 2 + roll.img_feat.numel()) * 4 / 1e6)},
            "timing": {"value": "K replays of the rollout captured as one CUDA graph"
                                + (" (two streams: the reward of an iteration runs beside the next iteration's observation; "
                                   "observe and step keep their order)" if args.streams == 2 else " (one stream)")
                                + ", CUDA events, max over ranks; its outputs equal a one-stream eager rollout's bit for bit",
                       "roofline": f"eager passes of {esteps} rollouts, each queued behind a 3 ms spin kernel: CUDA events between the "
                                   "two observe stages (per kernel) and around cmr_observe (the pair); share_of_step is relative to "
                                   "an iteration of the per-kernel pass",
                       "eager_us_per_iteration": m["iter_s"] * 1e6},
            "sustained": {"value": B * iters * k_long * world / dt_long, "unit": UNIT, "steps": k_long, "seconds": dt_long},
            "roofline": roofline, "roofline_secondary": roofline2, "cpu_baseline": cpu_base, "e2e": e2e,
            "secondary": secondary, "gpu_launches": int(launches), "clocks": clocks,
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
