"""The agent's 2-D head (CMRAgent.state_2d_embed, cuDNN) in NCHW and in channels_last, captured as a CUDA graph."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from oracle import reference_loader as rl
rl.put_on_path()
from config import KittiConfiguration
from models import CMRAgent
dev = torch.device("cuda:0")
config = KittiConfiguration()
for B in (1, 32):
    for fmt in ("nchw", "nhwc", "nhwc_in_only"):
        torch.manual_seed(1)
        agent = CMRAgent(config).to(dev).eval()
        head = agent.state_2d_embed
        x = torch.randn(B, 128, 40, 128, device=dev)
        if fmt == "nhwc":
            head = head.to(memory_format=torch.channels_last)
        if fmt != "nchw":
            x = x.contiguous(memory_format=torch.channels_last)
        with torch.no_grad():
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(3):
                    y = head(x)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                y = head(x)
            g.replay(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(50):
                g.replay()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 50
        print(f"B={B} {fmt}: {dt * 1e6:.1f} us per forward, out {tuple(y.shape)} sum {float(y.sum()):.6f}")
