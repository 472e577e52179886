"""Swin-L backbone on the pfd_b200 kernels — mirrors lib/model_zoo/swin.py (state-dict compatible).

Per block (swin.py:254-310):  LN -> [zero-pad + cyclic shift + window partition] (one gather kernel)
-> q/k/v^T projection GEMMs writing per-head layouts -> QK^T GEMM -> softmax(+rel-pos bias, +shift
mask, reference fp16 rounding) -> PV GEMM -> proj GEMM -> [window reverse + un-shift + crop + residual]
(one scatter kernel) -> LN -> fc1 GEMM(+GELU) -> fc2 GEMM(+residual).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch
import torch.nn as nn

from . import native as nv
from .attention import attend, project_heads
from .modules import Conv2d, LayerNorm, Linear, cached, pk_lin, pk_mat, pk_norm


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def relative_position_index(ws: int) -> torch.Tensor:
    """Index into the (2ws-1)^2 bias table for every (query, key) pair of a window (swin.py:159-169)."""
    ys, xs = np.meshgrid(np.arange(ws), np.arange(ws), indexing="ij")
    flat = np.stack([ys.reshape(-1), xs.reshape(-1)])                  # [2, ws*ws]
    rel = flat[:, :, None] - flat[:, None, :]                           # [2, N, N]
    idx = (rel[0] + ws - 1) * (2 * ws - 1) + (rel[1] + ws - 1)
    return torch.from_numpy(idx.astype(np.int64))


def shift_mask(H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    """[nW, N, N] additive mask (0 / -100) separating the 9 wrap-around regions of the padded,
    cyclically shifted map (swin.py:421-440)."""
    Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
    region = np.zeros((Hp, Wp), dtype=np.float32)
    bounds_h = [(0, Hp - ws), (Hp - ws, Hp - shift), (Hp - shift, Hp)]
    bounds_w = [(0, Wp - ws), (Wp - ws, Wp - shift), (Wp - shift, Wp)]
    cnt = 0
    for h0, h1 in bounds_h:
        for w0, w1 in bounds_w:
            region[h0:h1, w0:w1] = cnt
            cnt += 1
    win = region.reshape(Hp // ws, ws, Wp // ws, ws).transpose(0, 2, 1, 3).reshape(-1, ws * ws)
    diff = win[:, None, :] - win[:, :, None]
    return torch.from_numpy(np.where(diff != 0, -100.0, 0.0).astype(np.float32))


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = Linear(in_features, hidden_features)
        self.fc2 = Linear(hidden_features, in_features)


class WindowAttention(nn.Module):
    def __init__(self, dim, window_size, num_heads, qkv_bias=True):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, window_size, num_heads
        self.scale = (dim // num_heads) ** -0.5
        ws = window_size[0]
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) * (2 * ws - 1), num_heads))
        self.register_buffer("relative_position_index", relative_position_index(ws))
        self.qkv = Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = Linear(dim, dim)

    def bias_dense(self) -> torch.Tensor:
        """[heads, N, N] fp16 relative-position bias (swin.py:190-193), cached."""
        def build():
            n = self.window_size[0] * self.window_size[1]
            t = self.relative_position_bias_table.detach()[self.relative_position_index.view(-1)]
            return t.view(n, n, -1).permute(2, 0, 1).to(torch.float16).contiguous()
        return cached(self, "bias", [self.relative_position_bias_table], build)


class SwinTransformerBlock(nn.Module):
    def __init__(self, dim, num_heads, window_size, shift_size, mlp_ratio=4.0, qkv_bias=True):
        super().__init__()
        self.dim, self.num_heads, self.window_size, self.shift_size = dim, num_heads, window_size, shift_size
        self.norm1 = LayerNorm(dim)
        self.attn = WindowAttention(dim, to_2tuple(window_size), num_heads, qkv_bias)
        self.drop_path = nn.Identity()
        self.norm2 = LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))


class PatchMerging(nn.Module):
    def __init__(self, dim, norm_layer=None):
        super().__init__()
        self.dim = dim
        self.reduction = Linear(4 * dim, 2 * dim, bias=False)
        self.norm = LayerNorm(4 * dim)


class BasicLayer(nn.Module):
    def __init__(self, dim, depth, num_heads, window_size, mlp_ratio, qkv_bias, downsample):
        super().__init__()
        self.window_size, self.shift_size, self.depth = window_size, window_size // 2, depth
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(dim, num_heads, window_size, 0 if i % 2 == 0 else window_size // 2, mlp_ratio, qkv_bias)
            for i in range(depth)])
        self.downsample = PatchMerging(dim) if downsample else None


class PatchEmbed(nn.Module):
    def __init__(self, patch_size=4, in_chans=3, embed_dim=96, patch_norm=True):
        super().__init__()
        self.patch_size = to_2tuple(patch_size)
        self.in_chans, self.embed_dim = in_chans, embed_dim
        self.proj = Conv2d(in_chans, embed_dim, kernel_size=self.patch_size, stride=self.patch_size)
        self.norm = LayerNorm(embed_dim) if patch_norm else None


def run_swin_block(blk: SwinTransformerBlock, x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """x: channel-last [B, H, W, C]."""
    B, H, W, C = x.shape
    ws, shift, heads = blk.window_size, blk.shift_size, blk.num_heads
    d = C // heads
    g, b = pk_norm(blk.norm1)
    n1 = nv.layernorm(x, g, b, blk.norm1.eps)
    win = nv.window_gather(n1, ws, shift)                               # [Bw, ws*ws, C], pad tokens = 0
    Bw, N, _ = win.shape
    a = blk.attn
    w2d = win.reshape(Bw * N, C)
    wq, bq = pk_mat(a.qkv, a.qkv.weight, a.qkv.bias, "q", slice(0, C))
    wk, bk = pk_mat(a.qkv, a.qkv.weight, a.qkv.bias, "k", slice(C, 2 * C))
    wv, bv = pk_mat(a.qkv, a.qkv.weight, a.qkv.bias, "v", slice(2 * C, 3 * C))
    q = project_heads(w2d, wq, bq, Bw, N, heads, d)
    k = project_heads(w2d, wk, bk, Bw, N, heads, d)
    vt = project_heads(w2d, wv, bv, Bw, N, heads, d, transposed=True)
    nW = Bw // B
    o = attend(q, k, vt, B=Bw, heads=heads, Nq=N, Nk=N, scale=a.scale, bias=a.bias_dense(),
               mask=mask if shift > 0 else None, nwin=nW)
    wp, bp = pk_lin(a.proj)
    pr = nv.linear(o.reshape(Bw * N, C), wp, bp)
    x = nv.window_scatter(pr.reshape(Bw, N, C), B, H, W, ws, shift, x)  # + shortcut
    g, b = pk_norm(blk.norm2)
    n2 = nv.layernorm(x, g, b, blk.norm2.eps)
    w1, b1 = pk_lin(blk.mlp.fc1)
    w2, b2 = pk_lin(blk.mlp.fc2)
    h = nv.linear(n2.reshape(B * H * W, C), w1, b1, act=nv.ACT_GELU)
    return nv.linear(h, w2, b2, residual=x.reshape(B * H * W, C)).reshape(B, H, W, C)


class SwinTransformer(nn.Module):
    """swin.py:498-659 constructor signature; forward returns res2..res5 as channel-last tensors."""

    def __init__(self, pretrain_img_size=224, patch_size=4, in_chans=3, embed_dim=96, depths=(2, 2, 6, 2),
                 num_heads=(3, 6, 12, 24), window_size=7, mlp_ratio=4.0, qkv_bias=True, qk_scale=None,
                 drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.2, norm_layer=None, ape=False,
                 patch_norm=True, out_indices=(0, 1, 2, 3), frozen_stages=-1, use_checkpoint=False):
        super().__init__()
        if ape or qk_scale is not None:
            raise NotImplementedError("pfd_b200.SwinTransformer: ape / qk_scale are not used by swin.yaml")
        self.num_layers, self.embed_dim = len(depths), embed_dim
        self.out_indices, self.window_size = tuple(out_indices), window_size
        self.patch_embed = PatchEmbed(patch_size, in_chans, embed_dim, patch_norm)
        self.layers = nn.ModuleList([
            BasicLayer(int(embed_dim * 2 ** i), depths[i], num_heads[i], window_size, mlp_ratio, qkv_bias,
                       downsample=(i < self.num_layers - 1)) for i in range(self.num_layers)])
        self.num_features = [int(embed_dim * 2 ** i) for i in range(self.num_layers)]
        for i in self.out_indices:
            self.add_module(f"norm{i}", LayerNorm(self.num_features[i]))
        self._masks: Dict = {}

    def _mask(self, H, W, device):
        key = (H, W, str(device))
        if key not in self._masks:
            ws = self.window_size
            self._masks[key] = shift_mask(H, W, ws, ws // 2).to(device=device, dtype=torch.float16).contiguous()
        return self._masks[key]

    @torch.no_grad()
    def forward(self, img: torch.Tensor) -> Dict[str, torch.Tensor]:
        pe = self.patch_embed
        P = pe.patch_size[0]
        kp = (pe.in_chans * P * P + 7) // 8 * 8
        col = nv.patchify(img, P, kp)                                   # [B, Wh, Ww, kp]
        B, Wh, Ww, _ = col.shape
        w, b = pk_lin(pe.proj)
        x = nv.linear(col.reshape(B * Wh * Ww, kp), w, b).reshape(B, Wh, Ww, self.embed_dim)
        if pe.norm is not None:
            g, bb = pk_norm(pe.norm)
            x = nv.layernorm(x, g, bb, pe.norm.eps)
        outs = {}
        for i, layer in enumerate(self.layers):
            mask = self._mask(Wh, Ww, x.device)
            for blk in layer.blocks:
                x = run_swin_block(blk, x, mask)
            if i in self.out_indices:
                n = getattr(self, f"norm{i}")
                g, bb = pk_norm(n)
                outs[f"res{i + 2}"] = nv.layernorm(x, g, bb, n.eps)
            if layer.downsample is not None:
                ds = layer.downsample
                m = nv.patch_merge_gather(x)                            # [B, H2, W2, 4C]
                g, bb = pk_norm(ds.norm)
                m = nv.layernorm(m, g, bb, ds.norm.eps)
                Wh, Ww = (Wh + 1) // 2, (Ww + 1) // 2
                wr, _ = pk_lin(ds.reduction)
                x = nv.linear(m.reshape(B * Wh * Ww, m.shape[-1]), wr, None).reshape(B, Wh, Ww, wr.shape[0])
        return outs
