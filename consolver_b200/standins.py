"""Denoiser stand-ins — NOT the product.  BASELINE configs[1]/[3] time the sampler around a random-init denoiser
of the right architecture; diffusers is not installed in this image, so the SD1.5 U-Net topology and a FLUX-shaped
DiT block stack are written here in plain PyTorch (library kernels: cuDNN convs, cuBLAS GEMMs, SDPA attention).
They exist so `bench.py --with-denoiser` can report previews/s with the denoiser in the loop and show the solver's
share of a real step; nothing in the solver path depends on them."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, device=t.device, dtype=torch.float32) / half)
    args = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class _Res(nn.Module):
    def __init__(self, cin, cout, temb):
        super().__init__()
        self.n1, self.c1 = nn.GroupNorm(32, cin), nn.Conv2d(cin, cout, 3, padding=1)
        self.t = nn.Linear(temb, cout)
        self.n2, self.c2 = nn.GroupNorm(32, cout), nn.Conv2d(cout, cout, 3, padding=1)
        self.skip = nn.Conv2d(cin, cout, 1) if cin != cout else nn.Identity()

    def forward(self, x, temb):
        h = self.c1(F.silu(self.n1(x)))
        h = h + self.t(F.silu(temb))[:, :, None, None]
        h = self.c2(F.silu(self.n2(h)))
        return self.skip(x) + h


class _Attn(nn.Module):
    def __init__(self, dim, ctx_dim, heads):
        super().__init__()
        self.h = heads
        self.q, self.k, self.v = nn.Linear(dim, dim, bias=False), nn.Linear(ctx_dim, dim, bias=False), \
            nn.Linear(ctx_dim, dim, bias=False)
        self.o = nn.Linear(dim, dim)

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        B, L, C = x.shape
        q = self.q(x).view(B, L, self.h, C // self.h).transpose(1, 2)
        k = self.k(ctx).view(B, ctx.shape[1], self.h, C // self.h).transpose(1, 2)
        v = self.v(ctx).view(B, ctx.shape[1], self.h, C // self.h).transpose(1, 2)
        return self.o(F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, L, C))


class _Transformer2D(nn.Module):
    """GroupNorm -> 1x1 in -> [LN self-attn, LN cross-attn, LN GEGLU-FF] -> 1x1 out + residual (SD1.5 layout)."""

    def __init__(self, dim, ctx_dim, heads):
        super().__init__()
        self.norm = nn.GroupNorm(32, dim, eps=1e-6)
        self.pin, self.pout = nn.Conv2d(dim, dim, 1), nn.Conv2d(dim, dim, 1)
        self.n1, self.n2, self.n3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.a1, self.a2 = _Attn(dim, dim, heads), _Attn(dim, ctx_dim, heads)
        self.ff1, self.ff2 = nn.Linear(dim, dim * 8), nn.Linear(dim * 4, dim)

    def forward(self, x, ctx):
        B, C, H, W = x.shape
        h = self.pin(self.norm(x)).flatten(2).transpose(1, 2)
        h = h + self.a1(self.n1(h))
        h = h + self.a2(self.n2(h), ctx)
        a, g = self.ff1(self.n3(h)).chunk(2, dim=-1)
        h = h + self.ff2(a * F.gelu(g))
        return x + self.pout(h.transpose(1, 2).reshape(B, C, H, W))


class SD15UNet(nn.Module):
    """SD1.5 `UNet2DConditionModel` topology: channels (320, 640, 1280, 1280), 2 res blocks per level, transformer
    blocks (8 heads, cross-attention to 77x768 text states) on the first three levels, mid block, 3 res blocks per
    up level with skip concatenation.  ~860 M parameters, random init."""

    def __init__(self, ch=(320, 640, 1280, 1280), ctx_dim=768, heads=8, in_ch=4):
        super().__init__()
        temb = ch[0] * 4
        self.ch0 = ch[0]
        self.time = nn.Sequential(nn.Linear(ch[0], temb), nn.SiLU(), nn.Linear(temb, temb))
        self.conv_in = nn.Conv2d(in_ch, ch[0], 3, padding=1)
        self.down = nn.ModuleList()
        skips, c = [ch[0]], ch[0]
        for lvl, co in enumerate(ch):
            attn = lvl < len(ch) - 1
            for _ in range(2):
                self.down.append(nn.ModuleList([_Res(c, co, temb), _Transformer2D(co, ctx_dim, heads) if attn else None]))
                c = co
                skips.append(c)
            if lvl < len(ch) - 1:
                self.down.append(nn.ModuleList([nn.Conv2d(c, c, 3, stride=2, padding=1), None]))
                skips.append(c)
        self.mid = nn.ModuleList([_Res(c, c, temb), _Transformer2D(c, ctx_dim, heads), _Res(c, c, temb)])
        self.up = nn.ModuleList()
        for lvl, co in reversed(list(enumerate(ch))):
            attn = lvl < len(ch) - 1
            for j in range(3):
                cs = skips.pop()
                blk = [_Res(c + cs, co, temb), _Transformer2D(co, ctx_dim, heads) if attn else None,
                       nn.Conv2d(co, co, 3, padding=1) if (j == 2 and lvl > 0) else None]
                self.up.append(nn.ModuleList(blk))
                c = co
        self.norm_out, self.conv_out = nn.GroupNorm(32, c), nn.Conv2d(c, in_ch, 3, padding=1)

    def forward(self, x, t, encoder_hidden_states):
        t = torch.as_tensor(t, device=x.device).reshape(-1).expand(x.shape[0])
        temb = self.time(_timestep_embedding(t, self.ch0).to(x.dtype))
        h = self.conv_in(x)
        hs = [h]
        for blk, att in self.down:
            h = blk(h, temb) if isinstance(blk, _Res) else blk(h)
            if att is not None:
                h = att(h, encoder_hidden_states)
            hs.append(h)
        h = self.mid[2](self.mid[1](self.mid[0](h, temb), encoder_hidden_states), temb)
        for res, att, ups in self.up:
            h = res(torch.cat([h, hs.pop()], dim=1), temb)
            if att is not None:
                h = att(h, encoder_hidden_states)
            if ups is not None:
                h = ups(F.interpolate(h, scale_factor=2.0, mode="nearest"))
        return self.conv_out(F.silu(self.norm_out(h)))


class FluxLikeDiT(nn.Module):
    """FLUX-shaped transformer stand-in: packed latents [B, L, 64] (+ an equal number of context-image tokens for
    Kontext) -> hidden 3072 / 24 heads, `depth` pre-LN blocks with adaLN-style timestep modulation, -> [B, L, 64].
    FLUX.1 has 19 double + 38 single blocks (12 B parameters); `depth` is configurable so the stand-in fits a
    benchmark slot — the reported number states the depth used."""

    def __init__(self, in_dim=64, hidden=3072, heads=24, depth=8):
        super().__init__()
        self.inp, self.out = nn.Linear(in_dim, hidden), nn.Linear(hidden, in_dim)
        self.time = nn.Sequential(nn.Linear(256, hidden), nn.SiLU(), nn.Linear(hidden, hidden))
        self.blocks = nn.ModuleList()
        for _ in range(depth):
            self.blocks.append(nn.ModuleDict(dict(mod=nn.Linear(hidden, 6 * hidden), n1=nn.LayerNorm(hidden, elementwise_affine=False),
                                                  attn=_Attn(hidden, hidden, heads), n2=nn.LayerNorm(hidden, elementwise_affine=False),
                                                  f1=nn.Linear(hidden, 4 * hidden), f2=nn.Linear(4 * hidden, hidden))))

    def forward(self, tokens, t):
        temb = self.time(_timestep_embedding(torch.as_tensor(t, device=tokens.device).reshape(-1).expand(tokens.shape[0]),
                                             256).to(tokens.dtype))
        h = self.inp(tokens)
        for b in self.blocks:
            s1, g1, a1, s2, g2, a2 = b["mod"](F.silu(temb))[:, None].chunk(6, dim=-1)
            h = h + a1 * b["attn"](b["n1"](h) * (1 + g1) + s1)
            h = h + a2 * b["f2"](F.gelu(b["f1"](b["n2"](h) * (1 + g2) + s2), approximate="tanh"))
        return self.out(h)
