"""Stand-ins for python modules the reference imports but this image lacks.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

torch_scatter (rusty1s/pytorch_scatter, version unpinned by the reference -
it ships no requirements file) is used on the hot path at
/root/reference/environment/environment.py:79.  Its published behaviour for
the three functions the reference calls, restated:

  scatter_sum(src, index, dim):  out = zeros(size with size[dim] =
        index.max()+1); out.scatter_add_(dim, index, src)
  scatter_mean(src, index, dim): sum as above; count = scatter_sum(ones);
        count clamped to >= 1; out / count (true divide for float)
  scatter_max(src, index, dim):  (max values, argmax) - only needed so that
        /root/reference/models/PointNN.py imports.

open3d is imported (environment.py:9) but never called at run time.
"""
import sys
import types

import torch


def _broadcast(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    for _ in range(index.dim(), src.dim()):
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    index = _broadcast(index, src, dim)
    if out is None:
        size = list(src.size())
        if dim_size is not None:
            size[dim] = dim_size
        elif index.numel() == 0:
            size[dim] = 0
        else:
            size[dim] = int(index.max()) + 1
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    out = scatter_sum(src, index, dim, out, dim_size)
    dim_size = out.size(dim)
    index_dim = dim
    if index_dim < 0:
        index_dim = index_dim + src.dim()
    if index.dim() <= index_dim:
        index_dim = index.dim() - 1
    ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
    count = scatter_sum(ones, index, index_dim, None, dim_size)
    count[count < 1] = 1
    count = _broadcast(count, out, dim)
    if out.is_floating_point():
        out.true_divide_(count)
    else:
        out.div_(count, rounding_mode="floor")
    return out


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    index_b = _broadcast(index, src, dim)
    size = list(src.size())
    size[dim] = dim_size if dim_size is not None else int(index_b.max()) + 1
    vals = torch.full(size, float("-inf"), dtype=src.dtype, device=src.device)
    vals = vals.scatter_reduce(dim, index_b, src, reduce="amax", include_self=True)
    hit = src == vals.gather(dim, index_b)
    pos = torch.arange(src.size(dim), device=src.device)
    shape = [1] * src.dim()
    shape[dim] = -1
    pos = pos.view(shape).expand_as(src)
    sentinel = src.size(dim)
    cand = torch.where(hit, pos, torch.full_like(pos, sentinel))
    arg = torch.full(size, sentinel, dtype=torch.long, device=src.device)
    arg = arg.scatter_reduce(dim, index_b, cand, reduce="amin", include_self=True)
    return vals, arg


def install():
    """Register the shims in sys.modules (idempotent, never overrides a real module)."""
    if "torch_scatter" not in sys.modules:
        try:
            import torch_scatter  # noqa: F401
        except Exception:
            m = types.ModuleType("torch_scatter")
            m.scatter_sum = scatter_sum
            m.scatter_add = scatter_sum
            m.scatter_mean = scatter_mean
            m.scatter_max = scatter_max
            m.__cmr_shim__ = True
            sys.modules["torch_scatter"] = m
    if "open3d" not in sys.modules:
        try:
            import open3d  # noqa: F401
        except Exception:
            m = types.ModuleType("open3d")
            m.__cmr_shim__ = True
            sys.modules["open3d"] = m
