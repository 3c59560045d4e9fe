"""The agent's 3-D state tower on tcgen05 tensor cores.

Reference: ``CMRAgent.state_3d_embed`` (models/CMRAgent.py:25-29: four ``ConvBNReLURes1D`` blocks,
models/PointNN.py:260-282) and the loop of ``CMRAgent.forward`` over them (:92-101: block, max over the points,
``repeat`` + ``cat``).  EVAL MODE ONLY: every BatchNorm1d is folded into the convolution before it here (host
logic, a few 128 x 128 matrices), the folded fp32 matrices are packed once per set of weights by
``cmr_tower_pack`` and ``cmr_tower_forward`` (include/cmr_b200.h) runs the four blocks as hand-written sm_100a
kernels (cmr_agent_b200/csrc/tower_kernels.cuh).  Training (batch statistics, autograd) stays on the reference's
own modules - ``accelerate_agent`` only takes over ``forward`` while the module is in eval mode and autograd is off.

    tower = Tower3D.from_agent(agent)          # agent: the reference's CMRAgent, weights loaded
    embed_3d = tower(observation_3d)           # [B, 5, N] -> [B, 128]  (CMRAgent.py:101)
    accelerate_agent(agent)                    # agent(state_2d, state_3d) now uses it in eval/no_grad
                                               # (and ``Heads``: the actor-critic heads, CMRAgent.py:70-86,106-113)
"""
import torch

from . import _lib

EPS_DEFAULT = 1e-5
TOWER_DIMS = ((5, 64), (128, 64), (128, 64), (128, 128))      # CMRAgent.py:25-29 with embed_dim 64


def fold_conv_bn(sd, conv, bn, eps=EPS_DEFAULT):
    """(W', b') of a 1x1 Conv1d followed by an eval-mode BatchNorm1d (keys of ConvBNReLURes1D.state_dict())."""
    scale = sd[bn + ".weight"] / torch.sqrt(sd[bn + ".running_var"] + eps)
    W = sd[conv + ".weight"][:, :, 0] * scale[:, None]
    b = (sd[conv + ".bias"] - sd[bn + ".running_mean"]) * scale + sd[bn + ".bias"]
    return W.float().contiguous(), b.float().contiguous()


class Tower3D:
    """Packed weights of the four blocks + a cached workspace."""

    def __init__(self, state_dicts, device, eps=EPS_DEFAULT):
        if len(state_dicts) != 4:
            raise _lib.CmrError("the 3-D tower has four blocks (CMRAgent.py:25-29)")
        lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CmrError("Tower3D needs a CUDA device: cmr_agent_b200 has no CPU fallback")
        self.blobs = []
        self._keep = []
        for i, (sd, (cin, cout)) in enumerate(zip(state_dicts, TOWER_DIMS)):
            sd = {k: v.detach().to(self.device) for k, v in sd.items() if torch.is_tensor(v)}
            if tuple(sd["net.0.weight"].shape[:2]) != (cin, cin) or tuple(sd["net.3.weight"].shape[:2]) != (cout, cin):
                raise _lib.CmrError(f"block {i}: expected ConvBNReLURes1D({cin}, {cout}) (embed_dim 64)")
            W1, b1 = fold_conv_bn(sd, "net.0", "net.1", eps)
            W2, b2 = fold_conv_bn(sd, "net.3", "net.4", eps)
            Ws = bs = None
            if "shortcut.0.weight" in sd:
                Ws, bs = fold_conv_bn(sd, "shortcut.0", "shortcut.1", eps)
            kind = 0 if i == 0 else (2 if i == 3 else 1)
            if (kind == 2) != (Ws is None):
                raise _lib.CmrError(f"block {i}: unexpected shortcut structure")
            blob = torch.zeros(lib.cmr_tower_blob_bytes(kind) + 128, dtype=torch.uint8, device=self.device)
            off = (-blob.data_ptr()) % 128
            blob = blob[off: off + lib.cmr_tower_blob_bytes(kind)]
            _lib.call("cmr_tower_pack", kind, _lib.ptr(W1), _lib.ptr(b1), _lib.ptr(W2), _lib.ptr(b2), _lib.ptr(Ws),
                      _lib.ptr(bs), _lib.ptr(blob), _lib.stream())
            self._keep.append((W1, b1, W2, b2, Ws, bs))          # alive until the pack kernel has run
            self.blobs.append(blob)
        self._ws = None
        self._ws_key = None

    @classmethod
    def from_agent(cls, agent):
        """``agent``: the reference's CMRAgent (or anything with ``state_3d_embed``: four ConvBNReLURes1D)."""
        layers = list(agent.state_3d_embed)
        dev = next(agent.parameters()).device
        eps = layers[0].net[1].eps
        return cls([l.state_dict() for l in layers], dev, eps)

    def _workspace(self, B, N):
        if self._ws_key != (B, N):
            nbytes = _lib.load().cmr_tower_workspace_bytes(B, N)
            raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
            off = (-raw.data_ptr()) % 1024
            self._ws = raw[off: off + nbytes]
            self._ws_key = (B, N)
        return self._ws

    @torch.no_grad()
    def __call__(self, obs3d):
        obs3d = _lib.require_cuda(obs3d, "state_3d", torch.float32)
        if obs3d.dim() != 3 or obs3d.shape[1] != 5:
            raise _lib.CmrError("state_3d must be [B,5,N] (environment.py:121-124)")
        obs3d = obs3d if obs3d.is_contiguous() else obs3d.contiguous()
        B, _, N = obs3d.shape
        out = torch.empty(B, 128, device=obs3d.device, dtype=torch.float32)
        ws = self._workspace(B, N)
        _lib.call("cmr_tower_forward", _lib.ptr(obs3d), _lib.ptr(self.blobs[0]), _lib.ptr(self.blobs[1]),
                  _lib.ptr(self.blobs[2]), _lib.ptr(self.blobs[3]), _lib.ptr(ws), B, N, _lib.ptr(out), _lib.stream())
        return out


class Heads:
    """The actor-critic heads (models/CMRAgent.py:70-86: ``policy_r``, ``policy_t``, ``value`` - each
    Linear, LeakyReLU, Linear, LeakyReLU, Linear) as three launches of ``cmr_grouped_linear``: layer l of ALL heads at
    once, every head reading its own slice of the previous layer's output.  fp32, deterministic."""

    MAX_K = 256

    def __init__(self, heads):
        import ctypes
        import numpy as np
        lin = [[m for m in h if isinstance(m, torch.nn.Linear)] for h in heads]
        act = [[m for m in h if isinstance(m, torch.nn.LeakyReLU)] for h in heads]
        depth = len(lin[0])
        ok = depth >= 1 and all(len(l) == depth and len(a) == depth - 1 and len(list(h)) == 2 * depth - 1
                                for l, a, h in zip(lin, act, heads))
        slopes = {a.negative_slope for al in act for a in al}
        if not ok or len(slopes) > 1 or len(heads) > 8:
            raise _lib.CmrError("Heads: every head must be Linear (LeakyReLU Linear)* of one depth and one slope")
        if any(m.in_features > self.MAX_K or m.bias is None for l in lin for m in l):
            raise _lib.CmrError("Heads: layers wider than 256 inputs, or without a bias, stay on torch")
        if len({l[0].in_features for l in lin}) != 1:
            raise _lib.CmrError("Heads: the heads read the same embedding")
        self.slope = slopes.pop() if slopes else 0.0
        self.device = lin[0][0].weight.device
        self.in_features = lin[0][0].in_features
        self.layers = []
        for l in range(depth):
            mods = [h[l] for h in lin]
            W = torch.cat([m.weight.detach().float().reshape(-1) for m in mods]).contiguous()
            b = torch.cat([m.bias.detach().float() for m in mods]).contiguous()
            desc, n0, w_off, in_off = [], 0, 0, 0
            for m in mods:
                desc += [0 if l == 0 else in_off, m.in_features, n0, n0 + m.out_features, w_off]
                n0 += m.out_features
                w_off += m.weight.numel()
                in_off += m.in_features
            d = np.asarray(desc, dtype=np.int64)
            in_stride = self.in_features if l == 0 else in_off
            self.layers.append(dict(W=W, b=b, desc=d, dptr=d.ctypes.data_as(ctypes.c_void_p), groups=len(mods), N=n0,
                                    in_stride=in_stride, act=1 if l < depth - 1 else 0))
        self.splits = [m.out_features for m in (h[-1] for h in lin)]

    @torch.no_grad()
    def __call__(self, embedding):
        x = _lib.require_cuda(embedding, "state_embedding", torch.float32)
        x = x if x.is_contiguous() else x.contiguous()
        if x.dim() != 2 or x.shape[1] != self.in_features:
            raise _lib.CmrError(f"state_embedding must be [B,{self.in_features}]")
        B = x.shape[0]
        for L in self.layers:
            out = torch.empty(B, L["N"], device=x.device, dtype=torch.float32)
            _lib.call("cmr_grouped_linear", _lib.ptr(x), L["in_stride"], _lib.ptr(L["W"]), _lib.ptr(L["b"]), L["dptr"],
                      L["groups"], B, L["N"], float(self.slope), L["act"], _lib.ptr(out), L["N"], _lib.stream())
            x = out
        return torch.split(x, self.splits, dim=1)


class Head2D:
    """The convolutional part of the agent's 2-D head (models/CMRAgent.py:34-58: eight ``Conv2d 3x3 [BatchNorm2d]
    LeakyReLU [AvgPool2d]`` stages), EVAL MODE.  The convolutions stay cuDNN's (``F.conv2d`` without the bias); what
    follows each of them - bias add, BatchNorm2d, LeakyReLU, AvgPool2d: up to four elementwise launches over the whole
    feature map - is ONE launch of ``cmr_conv_epilogue`` with bias and BatchNorm folded into a per-channel scale/shift.
    ``modules``: the leading part of ``state_2d_embed`` up to and including its last AvgPool2d."""

    def __init__(self, modules, channels_last=True):
        nn = torch.nn
        # channels_last: the whole head runs on torch's channels_last layout ([B][H][W][C] in memory) - cuDNN's
        # tensor-core convolutions are NHWC kernels; handed NCHW tensors they convert every layer's input and output
        # (measured: the eight convolutions take 111 / 219 / 485 us at B = 1 / 8 / 32 in NCHW, 56 / 111 / 246 in
        # channels_last).  The observation is converted once; values and shapes are the same.
        self.channels_last = bool(channels_last)
        self.stages = []
        mods = list(modules)
        i = 0
        while i < len(mods):
            conv = mods[i]
            if not (isinstance(conv, nn.Conv2d) and conv.groups == 1 and tuple(conv.stride) == (1, 1) and
                    tuple(conv.dilation) == (1, 1) and conv.padding_mode == "zeros"):
                raise _lib.CmrError("Head2D: expected a plain Conv2d")
            i += 1
            bn = None
            if i < len(mods) and isinstance(mods[i], nn.BatchNorm2d):
                bn = mods[i]
                if not bn.track_running_stats or bn.running_mean is None:
                    raise _lib.CmrError("Head2D: BatchNorm2d without running statistics")
                i += 1
            if not (i < len(mods) and isinstance(mods[i], nn.LeakyReLU)):
                raise _lib.CmrError("Head2D: expected LeakyReLU after the convolution")
            slope = float(mods[i].negative_slope)
            i += 1
            pool = None
            if i < len(mods) and isinstance(mods[i], nn.AvgPool2d):
                pool = mods[i]
                if pool.padding not in (0, (0, 0)) or pool.ceil_mode or pool.divisor_override is not None:
                    raise _lib.CmrError("Head2D: unsupported AvgPool2d")
                i += 1
            self.stages.append(dict(conv=conv, bn=bn, slope=slope, pool=pool, scale=None, shift=None))
        self._sig = None

    def _signature(self):
        ts = []
        for st in self.stages:
            ts += [st["conv"].weight] + ([st["conv"].bias] if st["conv"].bias is not None else [])
            if st["bn"] is not None:
                ts += [t for t in (st["bn"].weight, st["bn"].bias, st["bn"].running_mean, st["bn"].running_var) if t is not None]
        return tuple((t.data_ptr(), t._version) for t in ts)

    def _fold(self):
        for st in self.stages:
            conv, bn = st["conv"], st["bn"]
            dev = conv.weight.device
            bias = conv.bias.detach().float() if conv.bias is not None else torch.zeros(conv.out_channels, device=dev)
            if bn is None:
                scale, shift = torch.ones(conv.out_channels, device=dev), bias
            else:
                g = bn.weight.detach().float() if bn.weight is not None else torch.ones(conv.out_channels, device=dev)
                beta = bn.bias.detach().float() if bn.bias is not None else torch.zeros(conv.out_channels, device=dev)
                scale = g / torch.sqrt(bn.running_var.float() + bn.eps)
                shift = (bias - bn.running_mean.float()) * scale + beta
            st["scale"], st["shift"] = scale.contiguous(), shift.contiguous()
            w = conv.weight.detach()
            st["weight"] = w.contiguous(memory_format=torch.channels_last) if self.channels_last else w

    def _pool_mode(self, pool, H, W):
        if pool is None:
            return 0
        k = pool.kernel_size if isinstance(pool.kernel_size, tuple) else (pool.kernel_size, pool.kernel_size)
        s_ = pool.stride if isinstance(pool.stride, tuple) else (pool.stride, pool.stride)
        if tuple(k) == (2, 2) and tuple(s_) == (2, 2) and H % 2 == 0 and W % (2 if self.channels_last else 4) == 0:
            return 1
        if tuple(k) == (H, W):
            return 2
        return -1

    @torch.no_grad()
    def __call__(self, x):
        sig = self._signature()
        if sig != self._sig:
            self._fold()
            self._sig = sig
        x = _lib.require_cuda(x, "state_2d", torch.float32)
        nhwc = self.channels_last and all(st["conv"].out_channels % 4 == 0 for st in self.stages)
        fmt = torch.channels_last if nhwc else torch.contiguous_format
        if nhwc and x.dim() == 4 and x.is_contiguous() and not x.is_contiguous(memory_format=torch.channels_last):
            # the observation arrives NCHW (environment.py:126): one transposing copy instead of torch's strided one
            y = torch.empty(x.shape, device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
            _lib.call("cmr_to_channels_last", _lib.ptr(x), x.shape[0], x.shape[1], x.shape[2], x.shape[3], _lib.ptr(y), _lib.stream())
            x = y
        for st in self.stages:
            conv = st["conv"]
            x = x.contiguous(memory_format=fmt)            # (a no-op on the converted observation and between the stages)
            x = torch.nn.functional.conv2d(x, st["weight"] if nhwc else conv.weight, None, conv.stride, conv.padding,
                                           conv.dilation, 1)
            x = x.contiguous(memory_format=fmt)
            B, C, H, W = x.shape
            mode = self._pool_mode(st["pool"], H, W)       # 0 none, 1 2x2, 2 whole map, -1 a pooling the kernel lacks
            if nhwc and mode == 1 and W % 2:
                mode = -1
            fused = max(mode, 0)
            if fused == 0 and not nhwc and (H * W) % 4 != 0:   # odd NCHW maps: torch's own ops on the folded form
                x = torch.nn.functional.leaky_relu(x * st["scale"].view(1, -1, 1, 1) + st["shift"].view(1, -1, 1, 1), st["slope"])
            else:
                shape = (B, C, H, W) if fused == 0 else ((B, C, H // 2, W // 2) if fused == 1 else (B, C, 1, 1))
                y = x if fused == 0 else torch.empty(shape, device=x.device, dtype=torch.float32, memory_format=fmt)
                _lib.call("cmr_conv_epilogue", _lib.ptr(x), _lib.ptr(st["scale"]), _lib.ptr(st["shift"]), st["slope"], fused,
                          1 if nhwc else 0, B, C, H, W, _lib.ptr(y), _lib.stream())
                x = y
            if mode == -1:                                 # a pooling shape the kernel does not cover
                x = st["pool"](x)
        return x


def accelerate_agent(agent):
    """Route the 3-D half of ``CMRAgent.forward`` (models/CMRAgent.py:92-101) through ``Tower3D`` whenever the module
    is in eval mode and autograd is off; otherwise the reference's own forward runs untouched.  The packed weights
    are rebuilt when any parameter or buffer of ``state_3d_embed`` has changed (``_version``)."""
    reference_forward = agent.forward
    state = {"tower": None, "sig": None, "heads": None, "hsig": None, "tail": None}
    head_modules = [agent.policy_r, agent.policy_t, agent.value]

    def _sig():
        return tuple((t.data_ptr(), t._version) for t in list(agent.state_3d_embed.parameters()) +
                     list(agent.state_3d_embed.buffers()))

    # the 1x1 tail of the 2-D head (CMRAgent.py:59-61: AvgPool2d over the whole map, then Conv2d 1x1 - LeakyReLU -
    # Conv2d 1x1 on a [B, 2f, 1, 1] tensor) is a two-layer MLP: it goes through the same grouped-linear launches
    # instead of two cuDNN convolutions with their layout conversions.  Anything shaped differently stays on torch.
    mods_2d = list(agent.state_2d_embed)
    pools = [i for i, m in enumerate(mods_2d) if isinstance(m, torch.nn.AvgPool2d)]
    cut = pools[-1] + 1 if pools else len(mods_2d)
    tail_2d = mods_2d[cut:]
    is_1x1 = lambda m: (isinstance(m, torch.nn.Conv2d) and tuple(m.kernel_size) == (1, 1) and tuple(m.stride) == (1, 1) and  # noqa: E731
                        tuple(m.padding) == (0, 0) and m.groups == 1 and m.bias is not None and m.in_channels <= Heads.MAX_K)
    tail_ok = (len(tail_2d) >= 1 and len(tail_2d) % 2 == 1 and
               all(is_1x1(m) if i % 2 == 0 else isinstance(m, torch.nn.LeakyReLU) for i, m in enumerate(tail_2d)))
    body_2d = torch.nn.Sequential(*mods_2d[:cut]) if tail_ok else agent.state_2d_embed
    try:                                  # the convolutional stages with fused epilogues; any other structure stays on torch
        body_2d = Head2D(mods_2d[:cut] if tail_ok else mods_2d)
    except _lib.CmrError:
        pass

    def _as_linear(conv):
        lin = torch.nn.Linear(conv.in_channels, conv.out_channels, device=conv.weight.device)
        lin.weight = torch.nn.Parameter(conv.weight.detach().reshape(conv.out_channels, conv.in_channels), requires_grad=False)
        lin.bias = torch.nn.Parameter(conv.bias.detach(), requires_grad=False)
        return lin

    def _hsig():
        return tuple((t.data_ptr(), t._version) for m in head_modules + (tail_2d if tail_ok else []) for t in m.parameters())

    def forward(state_2d, state_3d):
        if agent.training or torch.is_grad_enabled() or not state_3d.is_cuda:
            return reference_forward(state_2d, state_3d)
        sig = _sig()
        if state["sig"] != sig:
            state["tower"], state["sig"] = Tower3D.from_agent(agent), sig
        hsig = _hsig()
        if state["hsig"] != hsig:
            state["heads"], state["hsig"] = Heads(head_modules), hsig
            state["tail"] = Heads([torch.nn.Sequential(*[_as_linear(m) if i % 2 == 0 else m
                                                         for i, m in enumerate(tail_2d)])]) if tail_ok else None
        embed_2d = body_2d(state_2d)                                                # CMRAgent.py:89-90
        if state["tail"] is not None and embed_2d.shape[2:] == (1, 1):
            embed_2d = state["tail"](embed_2d.reshape(embed_2d.shape[0], -1))[0]
        else:
            for m in (tail_2d if tail_ok else []):
                embed_2d = m(embed_2d)
            embed_2d = embed_2d.reshape(embed_2d.shape[0], -1)
        embed_3d = state["tower"](state_3d)                                         # :92-101
        state_embedding = torch.cat([embed_2d, embed_3d], dim=1)                    # :103
        action_r_logits, action_t_logits, value = state["heads"](state_embedding)    # :106-113, three launches
        action_r_logits = action_r_logits.reshape(action_r_logits.shape[0], agent.degree_r, agent.config.num_steps)
        action_t_logits = action_t_logits.reshape(action_t_logits.shape[0], agent.degree_t, agent.config.num_steps)
        return action_r_logits, action_t_logits, value.unsqueeze(-1)

    reference_action = agent.action_from_logits              # the class's static method (models/CMRAgent.py:117-128)

    def action_from_logits(r_logits, t_logits, deterministic=False):
        """``deterministic=True`` on CUDA logits without autograd: one launch of ``cmr_deterministic_action`` (torch's
        probabilities bit for bit, the first index of the largest); anything else is the reference's own function."""
        ok = (deterministic and not torch.is_grad_enabled() and
              all(t.is_cuda and t.dtype == torch.float32 and t.dim() == 3 and t.shape[0] == r_logits.shape[0] and
                  t.stride(2) == 1 and t.stride(1) == t.shape[2] and t.stride(0) >= t.shape[1] * t.shape[2] and
                  t.data_ptr() % 4 == 0 for t in (r_logits, t_logits)) and
              r_logits.shape[2] == t_logits.shape[2] and 9 <= r_logits.shape[2] <= 16 and r_logits.shape[0] > 0)
        if not ok:
            return reference_action(r_logits, t_logits, deterministic)
        B, S = r_logits.shape[0], r_logits.shape[2]
        action_r = torch.empty(r_logits.shape[:2], device=r_logits.device, dtype=torch.int64)
        action_t = torch.empty(t_logits.shape[:2], device=t_logits.device, dtype=torch.int64)
        _lib.call("cmr_deterministic_action", _lib.ptr(r_logits), r_logits.shape[1], r_logits.stride(0), _lib.ptr(t_logits),
                  t_logits.shape[1], t_logits.stride(0), B, S, _lib.ptr(action_r), _lib.ptr(action_t), None, None, _lib.stream())
        return action_r, action_t

    agent.forward = forward
    agent.action_from_logits = action_from_logits
    agent._cmr_b200_reference_forward = reference_forward
    agent._cmr_b200_reference_action = reference_action
    return agent
