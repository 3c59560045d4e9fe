"""The eight 3x3 convolutions of the agent's 2-D head alone (cuDNN, no bias), NCHW against channels_last, as a CUDA graph."""
import os, sys, time
import torch
import torch.nn.functional as F
dev = torch.device("cuda:0")
shapes = [(40, 128), (40, 128), (20, 64), (20, 64), (10, 32), (10, 32), (5, 16), (5, 16)]
for B in (1, 8, 32):
    for fmt in ("nchw", "nhwc"):
        torch.manual_seed(0)
        mf = torch.channels_last if fmt == "nhwc" else torch.contiguous_format
        ws = [torch.randn(128, 128, 3, 3, device=dev).contiguous(memory_format=mf) * 0.03 for _ in shapes]
        xs = [torch.randn(B, 128, h, w, device=dev).contiguous(memory_format=mf) for h, w in shapes]
        def run():
            return [F.conv2d(x, w, None, 1, 1) for x, w in zip(xs, ws)]
        with torch.no_grad():
            s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(3): ys = run()
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g): ys = run()
            g.replay(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(50): g.replay()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 50
        print(f"B={B} {fmt}: {dt*1e6:.1f} us for the 8 convolutions; out contiguous-in-format {ys[0].is_contiguous(memory_format=mf)}")
