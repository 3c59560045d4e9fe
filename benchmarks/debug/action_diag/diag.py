"""Which step of cmr_deterministic_action differs from torch's Categorical(logits=x).probs?  Emulates the assumed
operation order with torch's own elementwise kernels and compares every stage bitwise."""
import ctypes, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(HERE))))
import torch
from cmr_agent_b200 import _lib

dev = torch.device("cuda:0")
S = 11
g = torch.Generator().manual_seed(1)
x = torch.cat([torch.randn(20000, S, generator=g) * s for s in (1.0, 0.01, 10.0)]).to(dev).reshape(-1, 3, S)
R = x.shape[0]


def frac(a, b):
    return float((a != b).float().mean())


def pad16(e):
    return torch.cat([e, torch.zeros(e.shape[:-1] + (16 - S,), device=dev)], -1)


def tree_down(v):          # shuffle-down 1, 2, 4 over 8 lanes, lane 0
    a = v[..., 0::2] + v[..., 1::2]
    b = a[..., 0::2] + a[..., 1::2]
    return b[..., 0] + b[..., 1]


def butterfly16(v):        # xor 8, 4, 2, 1
    a = v[..., :8] + v[..., 8:]
    b = a[..., :4] + a[..., 4:]
    c = b[..., :2] + b[..., 2:]
    return c[..., 0] + c[..., 1]


def seq(v):
    s = v[..., 0].clone()
    for i in range(1, v.shape[-1]):
        s = s + v[..., i]
    return s


m = x.amax(-1, keepdim=True)
e = (x - m).exp()
s_torch = e.sum(-1)
p16 = pad16(e)
cands = {
    "stride8 + down(1,2,4)": tree_down(p16[..., :8] + p16[..., 8:]),
    "sequential": seq(e),
    "16 lanes down(1,2,4,8)": (lambda a: (lambda b: (lambda c: c[..., 0] + c[..., 1])(b[..., 0::2] + b[..., 1::2]))(a[..., 0::2] + a[..., 1::2]))(p16[..., 0::2] + p16[..., 1::2]),
    "butterfly16": butterfly16(p16),
    "4 accumulators stride 1": ((e[..., 0] + e[..., 4] + e[..., 8]) + (e[..., 1] + e[..., 5] + e[..., 9])) + (e[..., 2] + e[..., 6] + e[..., 10]) + (e[..., 3] + e[..., 7]),
    "stride4 (4 lanes) + down(1,2)": (lambda t: (t[..., 0] + t[..., 1]) + (t[..., 2] + t[..., 3]))(
        torch.stack([(e[..., j] + e[..., j + 4]) + (e[..., j + 8] if j + 8 < S else 0) for j in range(4)], -1)),
}
print("logsumexp's sum, fraction of rows differing from torch's e.sum(-1):")
for k, v in cands.items():
    print(f"   {k:34s} {frac(v, s_torch):.4f}")
lse_t = torch.logsumexp(x, -1, keepdim=True)
mm = torch.where(m.abs() == float("inf"), torch.zeros_like(m), m)
print("lse from torch's sum + log + add vs torch.logsumexp:", frac(s_torch.unsqueeze(-1).log() + mm, lse_t))
cat = torch.distributions.Categorical(logits=x)
n = x - lse_t
print("n = x - lse vs Categorical.logits:", frac(n, cat.logits))
M = n.amax(-1, keepdim=True)
E = (n - M).exp()
sm = torch.softmax(n, -1)
print("softmax(n) vs Categorical.probs:", frac(sm, cat.probs))
P16 = pad16(E)
c2 = {"butterfly16": butterfly16(P16), "sequential": seq(E), "stride8 + down": tree_down(P16[..., :8] + P16[..., 8:]),
      "torch sum": E.sum(-1)}
print("softmax's sum -> p = E / sum, fraction of ELEMENTS differing from torch.softmax:")
for k, v in c2.items():
    print(f"   {k:34s} {frac(E / v.unsqueeze(-1), sm):.4f}")

# the kernel against torch
a_r = torch.empty(R, 3, device=dev, dtype=torch.int64); a_t = torch.empty_like(a_r)
p_r = torch.empty(R, 3, S, device=dev); p_t = torch.empty_like(p_r)
_lib.call("cmr_deterministic_action", _lib.ptr(x), 3, x.stride(0), _lib.ptr(x), 3, x.stride(0), R, S, _lib.ptr(a_r), _lib.ptr(a_t),
          _lib.ptr(p_r), _lib.ptr(p_t), _lib.stream())
torch.cuda.synchronize()
print("kernel probs vs Categorical.probs (elements):", frac(p_r, cat.probs), " actions:", frac(a_r, cat.probs.argmax(-1)))
emu = E / butterfly16(P16).unsqueeze(-1)
print("kernel probs vs the emulation with butterfly16 (elements):", frac(p_r, emu))

# elementwise functions of this build against torch's kernels
if not os.path.exists(os.path.join(HERE, "libelem.so")):          # this library's flags (cmr_agent_b200/build.py)
    import subprocess
    subprocess.run(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "--fmad=false", "--prec-div=true",
                    "--prec-sqrt=true", "--ftz=false", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-o",
                    os.path.join(HERE, "libelem.so"), os.path.join(HERE, "elem.cu")], check=True)
lib = ctypes.CDLL(os.path.join(HERE, "libelem.so"))
a = (torch.rand(1 << 20, device=dev) * -20.0)
b = torch.rand(1 << 20, device=dev) * 10 + 1e-3
eo, lo, do = torch.empty_like(a), torch.empty_like(a), torch.empty_like(a)
vp = lambda t: ctypes.c_void_p(t.data_ptr())
rc = lib.elem(vp(a), vp(b), a.numel(), vp(eo), vp(lo), vp(do), _lib.stream())
torch.cuda.synchronize()
print("rc", rc, "expf vs torch.exp:", frac(eo, a.exp()), " logf vs torch.log:", frac(lo, b.log()), " div:", frac(do, a / b))
sub = x - m
print("x - m via torch vs x + (-1)*m:", frac(sub, torch.add(x, m, alpha=-1)))
