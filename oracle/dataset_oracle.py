"""CPU restatement of the dataset-side geometric operations (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

  farthest_sample   dataset/KittiDataset.py:107-126 (FarthestSampler.sample; NuScenesDataset.py:25-44 is the same
                    code): float64 numpy, squared distances ((p0 - pts) ** 2).sum(axis=0), np.argmax = first
                    maximum, np.minimum update.  The start index is an explicit argument here (the reference draws
                    ``np.random.randint(len(pts))``, :118).
  nearest_index     what ``cKDTree(nodes.T).query(points.T, k=1)[1]`` returns (KittiDataset.py:365-366): the exact
                    nearest neighbour in float64.  scipy (1.18.1 here, unpinned by the reference) is a third-party
                    dependency; its published behaviour - exact Euclidean nearest neighbour - is restated by brute
                    force; ties (unspecified in scipy) go to the lowest index.

Pinned by tests/test_dataset_ops.py against the real reference class and scipy (build container) and against
tests/golden/dataset.npz (everywhere).
"""
import numpy as np


def farthest_sample(pts, k, init_idx):
    pts = np.asarray(pts, dtype=np.float64)
    out = np.zeros((3, k))
    idx = np.zeros(k, dtype=np.int64)
    out[:, 0] = pts[:, init_idx]                                         # :119
    idx[0] = init_idx
    d = ((out[:, 0:1] - pts) ** 2).sum(axis=0)                           # :121
    for i in range(1, k):
        j = int(np.argmax(d))                                            # :123
        out[:, i] = pts[:, j]
        idx[i] = j
        d = np.minimum(d, ((out[:, i:i + 1] - pts) ** 2).sum(axis=0))    # :126
    return out, idx


def nearest_index(points, nodes, chunk=4096):
    points = np.asarray(points, dtype=np.float64)
    nodes = np.asarray(nodes, dtype=np.float64)
    out = np.empty(points.shape[1], dtype=np.int64)
    for s in range(0, points.shape[1], chunk):
        p = points[:, s:s + chunk]
        d = ((p[:, :, None] - nodes[:, None, :]) ** 2).sum(axis=0)       # (dx^2 + dy^2) + dz^2
        out[s:s + chunk] = np.argmin(d, axis=1)
    return out
