"""Oracle: MST++ RGB->HSI network, fp32 torch-CPU functional restatement.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates reference
ml/MST_plus_plus/predict_code/architecture/MST_Plus_Plus.py:88-293 as pure functions over a
state dict that uses the reference's own parameter names, so the same dict can be loaded into the
real `MST_Plus_Plus` module (tools/make_golden.py does that to pin this file) and into the CUDA
path.  The reference ships no weights (`model_zoo` is git-ignored): parity is on seeded random
weights from `make_weights`.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

N_FEAT = 31
STAGES = 3          # MST_Plus_Plus(stage=3): three MST bodies
LEVELS = 2          # MST(stage=2): two encoder / decoder levels


def param_shapes() -> "OrderedDict[str, tuple]":
    """Names and shapes of every tensor in MST_Plus_Plus().state_dict() (227 tensors, 1 619 625
    parameters), in module registration order."""
    sh: "OrderedDict[str, tuple]" = OrderedDict()

    def msab(prefix, dim, heads):
        p = f"{prefix}.blocks.0."
        sh[p + "0.rescale"] = (heads, 1, 1)          # nn.Parameter: listed before sub-modules
        for n in ("to_q", "to_k", "to_v"):
            sh[p + f"0.{n}.weight"] = (N_FEAT * heads, dim)
        sh[p + "0.proj.weight"] = (dim, N_FEAT * heads)
        sh[p + "0.proj.bias"] = (dim,)
        sh[p + "0.pos_emb.0.weight"] = (dim, 1, 3, 3)
        sh[p + "0.pos_emb.2.weight"] = (dim, 1, 3, 3)
        sh[p + "1.fn.net.0.weight"] = (dim * 4, dim, 1, 1)
        sh[p + "1.fn.net.2.weight"] = (dim * 4, 1, 3, 3)
        sh[p + "1.fn.net.4.weight"] = (dim, dim * 4, 1, 1)
        sh[p + "1.norm.weight"] = (dim,)
        sh[p + "1.norm.bias"] = (dim,)
    # (PreNorm registers .fn before .norm)

    sh["conv_in.weight"] = (N_FEAT, 3, 3, 3)
    for s in range(STAGES):
        b = f"body.{s}."
        sh[b + "embedding.weight"] = (N_FEAT, N_FEAT, 3, 3)
        dim = N_FEAT
        for i in range(LEVELS):
            msab(b + f"encoder_layers.{i}.0", dim, dim // N_FEAT)
            sh[b + f"encoder_layers.{i}.1.weight"] = (dim * 2, dim, 4, 4)
            dim *= 2
        msab(b + "bottleneck", dim, dim // N_FEAT)
        for i in range(LEVELS):
            sh[b + f"decoder_layers.{i}.0.weight"] = (dim, dim // 2, 2, 2)   # ConvTranspose2d: (in,out,kh,kw)
            sh[b + f"decoder_layers.{i}.0.bias"] = (dim // 2,)
            sh[b + f"decoder_layers.{i}.1.weight"] = (dim // 2, dim, 1, 1)
            msab(b + f"decoder_layers.{i}.2", dim // 2, (dim // 2) // N_FEAT)
            dim //= 2
        sh[b + "mapping.weight"] = (N_FEAT, N_FEAT, 3, 3)
    sh["conv_out.weight"] = (N_FEAT, N_FEAT, 3, 3)
    return sh


def make_weights(seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic synthetic weights (independent of torch's module init order): fan-in scaled
    normals for matrices / kernels, and LayerNorm / rescale / bias values moved OFF their init
    values (1 / 1 / 0) so that every parameter matters in the parity check."""
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in param_shapes().items():
        if name.endswith("norm.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("rescale"):
            t = 1.0 + 0.25 * torch.rand(shape, generator=g)
        elif name.endswith("bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            if ".decoder_layers." in name and name.endswith(".0.weight"):
                fan_in = shape[0]                       # ConvTranspose2d k=2,s=2: one tap per output
            t = torch.randn(shape, generator=g) * (0.7 / fan_in ** 0.5)   # gain 0.7: |y| stays O(1)
        sd[name] = t.float()
    return sd


# ----------------------------------------------------------------------------- blocks
def ms_msa(x, sd, p, heads):
    """MS_MSA.forward (MST_Plus_Plus.py:110-139). x: [b,h,w,c]."""
    b, h, w, c = x.shape
    n = h * w
    xf = x.reshape(b, n, c)
    q_in = xf @ sd[p + "to_q.weight"].t()
    k_in = xf @ sd[p + "to_k.weight"].t()
    v_in = xf @ sd[p + "to_v.weight"].t()

    def split(t):  # 'b n (h d) -> b h d n'
        return t.reshape(b, n, heads, N_FEAT).permute(0, 2, 3, 1)

    q, k, v = split(q_in), split(k_in), split(v_in)
    q = F.normalize(q, dim=-1, p=2)          # L2 over ALL pixels (:127-128)
    k = F.normalize(k, dim=-1, p=2)
    attn = (k @ q.transpose(-2, -1)) * sd[p + "rescale"]
    attn = attn.softmax(dim=-1)
    y = attn @ v                              # b, heads, d, n
    y = y.permute(0, 3, 1, 2).reshape(b, n, heads * N_FEAT)
    out_c = (y @ sd[p + "proj.weight"].t() + sd[p + "proj.bias"]).view(b, h, w, c)
    vp = v_in.reshape(b, h, w, c).permute(0, 3, 1, 2)
    pe = F.conv2d(vp, sd[p + "pos_emb.0.weight"], padding=1, groups=c)
    pe = F.conv2d(F.gelu(pe), sd[p + "pos_emb.2.weight"], padding=1, groups=c)
    return out_c + pe.permute(0, 2, 3, 1)


def feed_forward(x, sd, p):
    """PreNorm(LayerNorm) + FeedForward (MST_Plus_Plus.py:57-65, :141-158). x: [b,h,w,c]."""
    c = x.shape[-1]
    y = F.layer_norm(x, (c,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5).permute(0, 3, 1, 2)
    y = F.gelu(F.conv2d(y, sd[p + "fn.net.0.weight"]))
    y = F.gelu(F.conv2d(y, sd[p + "fn.net.2.weight"], padding=1, groups=4 * c))
    y = F.conv2d(y, sd[p + "fn.net.4.weight"])
    return y.permute(0, 2, 3, 1)


def msab(x, sd, p, heads):
    """MSAB.forward with num_blocks=1 (MST_Plus_Plus.py:176-186). x: [b,c,h,w]."""
    x = x.permute(0, 2, 3, 1)
    x = ms_msa(x, sd, p + ".blocks.0.0.", heads) + x
    x = feed_forward(x, sd, p + ".blocks.0.1.") + x
    return x.permute(0, 3, 1, 2)


def mst(x, sd, p):
    """MST.forward (MST_Plus_Plus.py:240-268), stage=2, num_blocks=[1,1,1]."""
    fea = F.conv2d(x, sd[p + "embedding.weight"], padding=1)
    skips = []
    heads = 1
    for i in range(LEVELS):
        fea = msab(fea, sd, p + f"encoder_layers.{i}.0", heads)
        skips.append(fea)
        fea = F.conv2d(fea, sd[p + f"encoder_layers.{i}.1.weight"], stride=2, padding=1)
        heads *= 2
    fea = msab(fea, sd, p + "bottleneck", heads)
    for i in range(LEVELS):
        fea = F.conv_transpose2d(fea, sd[p + f"decoder_layers.{i}.0.weight"],
                                 sd[p + f"decoder_layers.{i}.0.bias"], stride=2)
        fea = F.conv2d(torch.cat([fea, skips[LEVELS - 1 - i]], dim=1), sd[p + f"decoder_layers.{i}.1.weight"])
        heads //= 2
        fea = msab(fea, sd, p + f"decoder_layers.{i}.2", heads)
    return F.conv2d(fea, sd[p + "mapping.weight"], padding=1) + x


@torch.no_grad()
def forward(x: torch.Tensor, sd) -> torch.Tensor:
    """MST_Plus_Plus.forward (MST_Plus_Plus.py:279-293). x: [b,3,h,w] float32 in [0,1]."""
    _, _, h_in, w_in = x.shape
    pad_h, pad_w = (8 - h_in % 8) % 8, (8 - w_in % 8) % 8
    x = F.pad(x, [0, pad_w, 0, pad_h], mode="reflect")
    x = F.conv2d(x, sd["conv_in.weight"], padding=1)
    h = x
    for s in range(STAGES):
        h = mst(h, sd, f"body.{s}.")
    h = F.conv2d(h, sd["conv_out.weight"], padding=1) + x
    return h[:, :, :h_in, :w_in]


def rgb_to_hsi(image, sd):
    """The part of predict_rgb_to_hsi_torch worth keeping (predict_torch.py:249-310, :12-19,
    :171-188) without fp16 autocast and OOM tiling: HWC uint8/float frame -> float01 -> centred
    reflect pad to a multiple of 16 -> forward -> crop -> HWC float32 cube."""
    import numpy as np
    a = np.asarray(image)
    if np.issubdtype(a.dtype, np.integer):          # predict_torch.py:12-19
        a = a.astype(np.float32) / 255.0
    else:
        a = a.astype(np.float32)
        if a.max() > 1.001:
            a = np.clip(a / 255.0, 0.0, 1.0)
    H, W = a.shape[:2]
    ph, pw = (16 - H % 16) % 16, (16 - W % 16) % 16
    top, left = ph // 2, pw // 2
    t = torch.from_numpy(a).permute(2, 0, 1).unsqueeze(0)
    t = F.pad(t, [left, pw - left, top, ph - top], mode="reflect")
    y = forward(t, sd)[0].permute(1, 2, 0)
    return y[top:top + H, left:left + W].contiguous().numpy().astype(np.float32)
