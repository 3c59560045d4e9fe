"""Episode sharding across the GPUs of one box.

Registration episodes are independent (the reference even loops over them serially,
/root/reference/environment/environment.py:39,279), so the multi-GPU path is: one process per
GPU, contiguous blocks of episodes per rank, no data-path collective, and ONE all-reduce of a
handful of scalars per evaluation - the cross-episode reductions the reference's drivers do on
the host (Test_Agent.py:198-206, Train_Agent.py:200-201,309).  NCCL on GPUs, gloo in CPU tests.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank).
    A plain single-process run returns (0, 1, 0) without creating a process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_range(total, rank, world):
    """Contiguous block [lo, hi) of `total` episodes owned by `rank` (sizes differ by at most 1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError((rank, world))
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def owner_of(episode, total, world):
    base, extra = divmod(total, world)
    cut = extra * (base + 1)
    return episode // (base + 1) if episode < cut else extra + (episode - cut) // max(base, 1)


class MetricSums:
    """The scalar payload of one evaluation: sums that make recall / mean / std of the rotation and
    translation errors and the mean reward (Test_Agent.py:198-206) after ONE all-reduce(SUM)."""

    FIELDS = ("count", "success", "err_r", "err_t", "err_r_sq", "err_t_sq", "reward")

    def __init__(self, device="cpu"):
        self.buf = torch.zeros(len(self.FIELDS), dtype=torch.float64, device=device)

    def add(self, err_r, err_t, reward=None, rte_thresh=5.0, rre_thresh=10.0):
        err_r = torch.as_tensor(err_r, dtype=torch.float64, device=self.buf.device).reshape(-1)
        err_t = torch.as_tensor(err_t, dtype=torch.float64, device=self.buf.device).reshape(-1)
        ok = (err_t < rte_thresh) & (err_r < rre_thresh)                      # Test_Agent.py:198
        vals = [float(err_r.numel()), ok.sum(), err_r.sum(), err_t.sum(), (err_r ** 2).sum(), (err_t ** 2).sum(),
                torch.as_tensor(0.0 if reward is None else reward, dtype=torch.float64).sum()]
        self.buf += torch.stack([torch.as_tensor(v, dtype=torch.float64, device=self.buf.device) for v in vals])
        return self

    def all_reduce(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM)
        return self

    def summary(self):
        v = dict(zip(self.FIELDS, self.buf.tolist()))
        n = max(v["count"], 1.0)
        mean_r, mean_t = v["err_r"] / n, v["err_t"] / n
        return {
            "episodes": int(v["count"]),
            "recall": v["success"] / n,
            "rre_mean": mean_r, "rte_mean": mean_t,
            "rre_std": max(v["err_r_sq"] / n - mean_r ** 2, 0.0) ** 0.5,
            "rte_std": max(v["err_t_sq"] / n - mean_t ** 2, 0.0) ** 0.5,
            "reward_mean": v["reward"] / n,
        }


def max_over_ranks(value, device):
    """Max of a python float over all ranks (timings are reported as the slowest rank)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
