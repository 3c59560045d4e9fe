"""Where does environment.sample_image_features differ from its CPU statement?   python benchmarks/debug/sample_check.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cmr_agent_b200 import environment as env  # noqa: E402
from oracle import sample_oracle  # noqa: E402
from test_sample import _case  # noqa: E402
from tests.helpers import to_device  # noqa: E402

data, pose = _case(2, 40960, 160, 512, 5, True, 64)
want, cam = sample_oracle.sample_image_features(data, pose)
dev = torch.device("cuda:0")
from oracle import env_oracle  # noqa: E402
dd = to_device(data, dev)
dd["_cmr_b200_mean_override"] = env_oracle.cloud_mean(data["pc"]).reshape(2, 3)      # the same fp32 mean on both sides
got, gcam = env.sample_image_features(dd, pose.to(dev))
got = got.cpu()
d = (got - want).abs()
bad_pt = (d > 0).any(dim=1)
print("points in frustum", int(cam.sum()), "points with a differing channel", int(bad_pt.sum()), "max abs diff", float(d.max()),
      "scale", float(want.abs().max()))
u, v, _ = sample_oracle.project_all(data, pose)
idx = bad_pt.nonzero()[:5]
for b, n in idx.tolist():
    print("point", b, n, "u", float(u[b, n]), "v", float(v[b, n]), "channels differing", int((d[b, :, n] > 0).sum()),
          "max", float(d[b, :, n].max()))
