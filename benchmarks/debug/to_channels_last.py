"""cmr_to_channels_last against torch's x.contiguous(memory_format=channels_last) on the observation's shape
[B, 128, 40, 128] (CUDA events, L2 flushed by the 168 MB the B = 32 case moves; smaller batches stay in L2 as they do in
the agent's loop)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cmr_agent_b200 import _lib

dev = torch.device("cuda:0")
for B in (1, 8, 32):
    x = torch.randn(B, 128, 40, 128, device=dev)
    y = torch.empty_like(x, memory_format=torch.channels_last)

    def ours():
        _lib.call("cmr_to_channels_last", _lib.ptr(x), B, 128, 40, 128, _lib.ptr(y), _lib.stream())

    def torchs():
        y.copy_(x)

    out = {"batch": B, "bytes": 2 * x.numel() * 4}
    for name, fn in (("torch_us", torchs), ("ours_us", ours)):
        for _ in range(5):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) * 1e3 / 50
    out["ours_gbs"] = out["bytes"] / out["ours_us"] / 1e3
    assert torch.equal(y, x.contiguous(memory_format=torch.channels_last))
    print(json.dumps(out))
