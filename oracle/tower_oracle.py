"""CPU restatement of the agent's 3-D tower in eval mode (TEST INFRASTRUCTURE ONLY).

Reference: ``CMRAgent.forward`` models/CMRAgent.py:92-101 over ``state_3d_embed`` (:25-29), four ``ConvBNReLURes1D``
blocks (models/PointNN.py:260-282).  NO CUDA PATH EXISTS YET for this row: this file pins the algebra the kernel of a
later round will use, against the real reference modules (tests/test_tower_oracle.py, build container only) and against
tests/golden/tower.npz everywhere:

* eval-mode BatchNorm1d folds into the 1 x 1 convolution before it:  W' = W * g / sqrt(var + eps),
  b' = (b - mean) * g / sqrt(var + eps) + beta   (PointNN.py:265-270, 276-279);
* blocks 2-4 see ``cat([feat, max.repeat(N)])`` (CMRAgent.py:97-99): the half of every first-layer product that meets
  the repeated max is the same for all points of an episode, i.e. a per-episode bias ``W[:, f:] @ max`` - but only where
  the input enters a convolution directly (the block's first convolution and its shortcut).

Parity bar (tests/test_tower_oracle.py): 1e-5 of the output's scale - the reference's own float32 evaluation is only
5e-4 accurate elementwise on nearly cancelling outputs, so an elementwise bound cannot be met by any regrouping.

Weights are plain tensors keyed like the reference's ``state_dict`` so that nothing here imports the reference.
"""
import torch

EPS = 1e-5            # nn.BatchNorm1d default (PointNN.py:266)
SLOPE = 0.2           # LeakyReLU(negative_slope=0.2) (PointNN.py:267,272)
TOWER = ((5, 64), (128, 64), (128, 64), (128, 128))   # CMRAgent.py:25-29 with embed_dim = 64 (config/KittiConfig.py:63)


def make_state(seed, cin, cout):
    """Seeded weights of one block (keys of ConvBNReLURes1D.state_dict()), running statistics included."""
    g = torch.Generator().manual_seed(seed)

    def conv(prefix, i, o):
        return {prefix + ".weight": torch.randn(o, i, 1, generator=g) / (i ** 0.5), prefix + ".bias": torch.randn(o, generator=g) * 0.1}

    def bn(prefix, c):
        return {prefix + ".weight": 1.0 + 0.2 * torch.randn(c, generator=g), prefix + ".bias": 0.1 * torch.randn(c, generator=g),
                prefix + ".running_mean": 0.2 * torch.randn(c, generator=g), prefix + ".running_var": 0.5 + torch.rand(c, generator=g),
                prefix + ".num_batches_tracked": torch.tensor(7)}

    sd = {}
    sd.update(conv("net.0", cin, cin)); sd.update(bn("net.1", cin))
    sd.update(conv("net.3", cin, cout)); sd.update(bn("net.4", cout))
    if cin != cout:                                             # PointNN.py:274-279
        sd.update(conv("shortcut.0", cin, cout)); sd.update(bn("shortcut.1", cout))
    return sd


def fold(sd, conv, bn):
    """(W', b') of conv followed by eval-mode BatchNorm."""
    scale = sd[bn + ".weight"] / torch.sqrt(sd[bn + ".running_var"] + EPS)
    W = sd[conv + ".weight"][:, :, 0] * scale[:, None]
    b = (sd[conv + ".bias"] - sd[bn + ".running_mean"]) * scale + sd[bn + ".bias"]
    return W, b


def lrelu(x):
    return torch.where(x >= 0, x, x * SLOPE)


def block(sd, feat, pooled=None):
    """One ConvBNReLURes1D on ``cat([feat, pooled.repeat(N)])`` without building the concatenation.
    feat [B, f, N]; pooled [B, f'] or None (first block).  Returns [B, cout, N]."""
    W1, b1 = fold(sd, "net.0", "net.1")
    W2, b2 = fold(sd, "net.3", "net.4")
    f = feat.shape[1]
    h = torch.einsum("oi,bin->bon", W1[:, :f], feat) + b1[None, :, None]
    if pooled is not None:
        h = h + torch.einsum("oi,bi->bo", W1[:, f:], pooled)[:, :, None]      # the per-episode bias
    y = torch.einsum("oi,bin->bon", W2, lrelu(h)) + b2[None, :, None]
    if "shortcut.0.weight" in sd:
        Ws, bs = fold(sd, "shortcut.0", "shortcut.1")
        sc = torch.einsum("oi,bin->bon", Ws[:, :f], feat) + bs[None, :, None]
        if pooled is not None:
            sc = sc + torch.einsum("oi,bi->bo", Ws[:, f:], pooled)[:, :, None]
    else:                                                       # identity shortcut: the concatenated input itself
        sc = feat if pooled is None else torch.cat([feat, pooled[:, :, None].expand(-1, -1, feat.shape[2])], dim=1)
    return lrelu(y + sc)


def tower(states, obs3d):
    """CMRAgent.py:92-101: obs3d [B, 5, N] -> embed_3d [B, 128]."""
    feat, pooled = obs3d, None
    for sd in states:
        feat = block(sd, feat, pooled)
        pooled = feat.max(dim=2)[0]                             # :96
    return pooled                                               # :101 (the last block's max, [B, 2f])
