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


def accelerate_agent(agent):
    """Route the 3-D half of ``CMRAgent.forward`` (models/CMRAgent.py:92-101) through ``Tower3D`` whenever the module
    is in eval mode and autograd is off; otherwise the reference's own forward runs untouched.  The packed weights
    are rebuilt when any parameter or buffer of ``state_3d_embed`` has changed (``_version``)."""
    reference_forward = agent.forward
    state = {"tower": None, "sig": None}

    def _sig():
        return tuple((t.data_ptr(), t._version) for t in list(agent.state_3d_embed.parameters()) +
                     list(agent.state_3d_embed.buffers()))

    def forward(state_2d, state_3d):
        if agent.training or torch.is_grad_enabled() or not state_3d.is_cuda:
            return reference_forward(state_2d, state_3d)
        sig = _sig()
        if state["sig"] != sig:
            state["tower"], state["sig"] = Tower3D.from_agent(agent), sig
        embed_2d = agent.state_2d_embed(state_2d)                                   # CMRAgent.py:89-90
        embed_2d = embed_2d.view(embed_2d.shape[0], -1)
        embed_3d = state["tower"](state_3d)                                         # :92-101
        state_embedding = torch.cat([embed_2d, embed_3d], dim=1)                    # :103
        action_r_logits = agent.policy_r(state_embedding)                           # :106-110
        action_t_logits = agent.policy_t(state_embedding)
        action_r_logits = action_r_logits.view(action_r_logits.shape[0], agent.degree_r, agent.config.num_steps)
        action_t_logits = action_t_logits.view(action_t_logits.shape[0], agent.degree_t, agent.config.num_steps)
        value = agent.value(state_embedding).unsqueeze(-1)                          # :112-113
        return action_r_logits, action_t_logits, value

    agent.forward = forward
    agent._cmr_b200_reference_forward = reference_forward
    return agent
