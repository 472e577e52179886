"""SeeCoder (Swin-L -> multi-scale Decoder -> QueryTransformer) on the pfd_b200 kernels.

Mirrors lib/model_zoo/seecoder.py (constructor arguments, registry types and state-dict keys):
  Decoder           seecoder.py:328-428     QueryTransformer  seecoder.py:434-550
  PPE_MLP           seecoder.py:262-310     SemanticContextEncoder seecoder.py:556-578

Note: the reference Decoder feeds [bs, L, C] to a sequence-first
nn.MultiheadAttention, i.e. it attends over the *batch* axis.  For bs == 1 (the only way app.py calls
it, app.py:234-235) that degenerates to out_proj(v_proj(x)), which is what is implemented; bs > 1
would mix different reference images and is rejected loudly rather than silently "fixed".
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import native as nv
from .attention import attend, project_heads
from .modules import (Conv2d, Embedding, GroupNorm, IndexedSequential, LayerNorm, Linear, MultiheadAttention,
                      cached, pk_lin, pk_mat, pk_norm, pk_vec)


class Conv2d_Convenience(Conv2d):
    """seecoder.py:45-58: conv with an attached norm (state-dict key '<name>.norm.*')."""

    def __init__(self, *args, **kwargs):
        norm = kwargs.pop("norm", None)
        kwargs.pop("activation", None)
        super().__init__(*args, **kwargs)
        self.norm = norm


class DecoderLayer(nn.Module):
    def __init__(self, dim, feedforward_dim, n_heads):
        super().__init__()
        self.self_attn = MultiheadAttention(dim, n_heads, dropout=0.0)
        self.norm1 = LayerNorm(dim)
        self.linear1 = Linear(dim, feedforward_dim)
        self.linear2 = Linear(feedforward_dim, dim)
        self.norm2 = LayerNorm(dim)


class DecoderLayerStacked(nn.Module):
    def __init__(self, dim, feedforward_dim, n_heads, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([DecoderLayer(dim, feedforward_dim, n_heads) for _ in range(num_layers)])
        self.num_layers = num_layers


class Decoder(nn.Module):
    """seecoder.py:328-428 (no FPN tags, as in seecoder.yaml)."""

    def __init__(self, inchannels, trans_input_tags, trans_num_layers, trans_dim, trans_nheads, trans_dropout,
                 trans_feedforward_dim):
        super().__init__()
        inchannels = dict(inchannels)
        self.trans_tags = sorted(k for k in inchannels if k in trans_input_tags)
        if sorted(inchannels.keys()) != self.trans_tags:
            raise NotImplementedError("pfd_b200.Decoder: FPN (non-transformer) tags are not used by seecoder.yaml")
        self.all_tags = sorted(inchannels.keys())
        self.trans_dim, self.nheads = trans_dim, trans_nheads
        self.inproj_layers = nn.ModuleDict({
            t: IndexedSequential(Conv2d(inchannels[t], trans_dim, 1), GroupNorm(32, trans_dim)) for t in self.trans_tags})
        self.transformer = DecoderLayerStacked(trans_dim, trans_feedforward_dim, trans_nheads, trans_num_layers)
        self.level_embed = nn.Parameter(torch.zeros(len(self.trans_tags), trans_dim))
        self.lateral_layers = nn.ModuleDict({
            t: Conv2d_Convenience(inchannels[t], trans_dim, 1, bias=False, norm=GroupNorm(32, trans_dim))
            for t in self.all_tags})
        self.output_layers = nn.ModuleDict()

    @torch.no_grad()
    def forward(self, features: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """features: channel-last [1, h, w, C_tag] -> channel-last [1, h, w, trans_dim] per tag."""
        E = self.trans_dim
        nv.gn_reset()                                                   # GroupNorm scratch ring (native.groupnorm)
        tags = self.trans_tags[::-1]                                    # res5, res4, res3 (seecoder.py:397)
        bs = features[tags[0]].shape[0]
        if bs != 1:
            raise NotImplementedError("SeeCoder Decoder: batch > 1 attends across images in the reference "
                                      "(seecoder.py:70,83); encode one image per call as app.py does")
        shapes = {t: features[t].shape[1:3] for t in tags}
        lens = [shapes[t][0] * shapes[t][1] for t in tags]
        L = sum(lens)
        h = torch.empty((L, E), device=features[tags[0]].device, dtype=torch.float16)
        lvl = pk_vec(self, self.level_embed, "level_embed")
        off = 0
        for idx, t in enumerate(tags):
            conv, gn = self.inproj_layers[t][0], self.inproj_layers[t][1]
            w, b = pk_lin(conv)
            hh, ww = shapes[t]
            y = nv.conv1x1(features[t], w, b)
            g, bb = pk_norm(gn)
            seg = h[off:off + lens[idx]]
            nv.groupnorm(y, g, bb, gn.eps, silu=False, out=seg.view(1, hh, ww, E))
            nv.add_rowvec(seg, lvl[idx], out=seg)
            off += lens[idx]
        for layer in self.transformer.layers:
            at = layer.self_attn
            wv, bv = pk_mat(at, at.in_proj_weight, at.in_proj_bias, "v", slice(2 * E, 3 * E))
            wo, bo = pk_lin(at.out_proj)
            a = nv.linear(nv.linear(h, wv, bv), wo, bo)                 # softmax over a single key == 1
            g, bb = pk_norm(layer.norm1)
            h = nv.layernorm(a, g, bb, layer.norm1.eps, residual=h)
            w1, b1 = pk_lin(layer.linear1)
            w2, b2 = pk_lin(layer.linear2)
            f = nv.linear(nv.linear(h, w1, b1, act=nv.ACT_RELU), w2, b2)
            g, bb = pk_norm(layer.norm2)
            h = nv.layernorm(f, g, bb, layer.norm2.eps, residual=h)
        out, off = {}, 0
        for idx, t in enumerate(tags):
            hh, ww = shapes[t]
            lat = self.lateral_layers[t]
            w, _ = pk_lin(lat)
            y = nv.conv1x1(features[t], w, None)
            g, bb = pk_norm(lat.norm)
            y = nv.groupnorm(y, g, bb, lat.norm.eps, silu=False)
            out[t] = nv.axpby(h[off:off + lens[idx]].view(1, hh, ww, E), 1.0, y, 1.0)
            off += lens[idx]
        return out


class _AttnLayer(nn.Module):
    def __init__(self, channels, nhead, attn_name):
        super().__init__()
        setattr(self, attn_name, MultiheadAttention(channels, nhead, dropout=0.0))
        self.norm = LayerNorm(channels)


class SelfAttentionLayer(_AttnLayer):
    def __init__(self, channels, nhead, **_):
        super().__init__(channels, nhead, "self_attn")


class CrossAttentionLayer(_AttnLayer):
    def __init__(self, channels, nhead, **_):
        super().__init__(channels, nhead, "multihead_attn")


class FeedForwardLayer(nn.Module):
    def __init__(self, channels, hidden_channels=2048, **_):
        super().__init__()
        self.linear1 = Linear(channels, hidden_channels)
        self.linear2 = Linear(hidden_channels, channels)
        self.norm = LayerNorm(channels)


class PPE_MLP(nn.Module):
    """seecoder.py:262-310 — importable as lib.model_zoo.seecoder.PPE_MLP for app.py:166-175."""

    def __init__(self, freq_num=20, freq_max=None, out_channel=768, mlp_layer=3):
        super().__init__()
        self.freq_num, self.freq_max, self.out_channel, self.mlp_layer = freq_num, freq_max, out_channel, mlp_layer
        layers, cin = [], freq_num * 4
        for i in range(mlp_layer):
            layers.append(Linear(cin, out_channel, bias=True))
            if i != mlp_layer - 1:
                layers.append(nn.SiLU())
            cin = out_channel
        self.mlp = IndexedSequential(*layers)

    @torch.no_grad()
    def forward(self, x_hw, device) -> torch.Tensor:
        """Positional map for an (h, w) feature grid as tokens [h*w, out_channel] (eval-mode path)."""
        h, w = x_hw
        minlen = min(h, w)
        dt = torch.float16
        he, we = torch.meshgrid(torch.arange(h, device=device), torch.arange(w, device=device), indexing="ij")
        he = ((he + 0.5 - h / 2) / minlen * (2 * math.pi)).to(dt)
        we = ((we + 0.5 - w / 2) / minlen * (2 * math.pi)).to(dt)
        dim_t = torch.linspace(0, 1, self.freq_num, dtype=torch.float32, device=device)
        fmax = self.freq_max if self.freq_max is not None else minlen / 2
        dim_t = fmax ** dim_t.to(dt)
        ph, pw = he[:, :, None] * dim_t, we[:, :, None] * dim_t
        feat = torch.cat((ph.sin(), ph.cos(), pw.sin(), pw.cos()), dim=-1).reshape(h * w, -1).contiguous()
        lins = [m for m in self.mlp if isinstance(m, Linear)]
        for i, lin in enumerate(lins):
            wgt, b = pk_lin(lin)
            feat = nv.linear(feat, wgt, b, act=nv.ACT_SILU if i < len(lins) - 1 else nv.ACT_NONE)
        return feat


class QueryTransformer(nn.Module):
    """seecoder.py:434-550."""

    def __init__(self, in_channels, hidden_dim, num_queries=(8, 144), nheads=8, num_layers=9,
                 feedforward_dim=2048, mask_dim=256, pre_norm=False, num_feature_levels=3,
                 enforce_input_project=False, with_fea2d_pos=True):
        super().__init__()
        if pre_norm or in_channels != hidden_dim or enforce_input_project:
            raise NotImplementedError("pfd_b200.QueryTransformer supports the seecoder.yaml configuration "
                                      "(post-norm, in_channels == hidden_dim)")
        self.pe_layer = PPE_MLP(20, None, hidden_dim, 3) if with_fea2d_pos else None
        self.input_proj = None
        self.num_heads, self.num_layers, self.hidden_dim = nheads, num_layers, hidden_dim
        self.transformer_selfatt_layers = nn.ModuleList([SelfAttentionLayer(hidden_dim, nheads) for _ in range(num_layers)])
        self.transformer_crossatt_layers = nn.ModuleList([CrossAttentionLayer(hidden_dim, nheads) for _ in range(num_layers)])
        self.transformer_feedforward_layers = nn.ModuleList([FeedForwardLayer(hidden_dim, feedforward_dim) for _ in range(num_layers)])
        self.num_queries = list(num_queries)
        nq = sum(self.num_queries)
        self.init_query = Embedding(nq, hidden_dim)
        self.query_pos_embedding = Embedding(nq, hidden_dim)
        self.num_feature_levels = num_feature_levels
        self.level_embed = Embedding(num_feature_levels, hidden_dim)

    def _mha(self, at: MultiheadAttention, q_in, k_in, v_in, Nq, Nk):
        E, H = self.hidden_dim, self.num_heads
        d = E // H
        wq, bq = pk_mat(at, at.in_proj_weight, at.in_proj_bias, "q", slice(0, E))
        wk, bk = pk_mat(at, at.in_proj_weight, at.in_proj_bias, "k", slice(E, 2 * E))
        wv, bv = pk_mat(at, at.in_proj_weight, at.in_proj_bias, "v", slice(2 * E, 3 * E))
        q = project_heads(q_in, wq, bq, 1, Nq, H, d)
        k = project_heads(k_in, wk, bk, 1, Nk, H, d)
        vt = project_heads(v_in, wv, bv, 1, Nk, H, d, transposed=True)
        o = attend(q, k, vt, B=1, heads=H, Nq=Nq, Nk=Nk, scale=d ** -0.5)
        wo, bo = pk_lin(at.out_proj)
        return nv.linear(o.reshape(Nq, E), wo, bo)

    @torch.no_grad()
    def forward(self, x: Sequence[torch.Tensor]) -> torch.Tensor:
        """x: list of channel-last [1, h, w, E] maps (res3, res4, res5) -> [1, sum(num_queries), E]."""
        assert len(x) == self.num_feature_levels
        E = self.hidden_dim
        if x[0].shape[0] != 1:
            raise NotImplementedError("QueryTransformer: one reference image per call (see Decoder note)")
        lvl = pk_vec(self, self.level_embed.weight, "lvl")
        fea, pos = [], []
        for i, xi in enumerate(x):
            _, h, w, _ = xi.shape
            fea.append(nv.add_rowvec(xi.reshape(h * w, E), lvl[i]))
            pos.append(self.pe_layer.forward((h, w), xi.device) if self.pe_layer is not None else None)
        ngq, nlq = self.num_queries
        Q = pk_vec(self, self.init_query.weight, "iq").clone()          # rows [0:ngq) global, [ngq:) local
        P = pk_vec(self, self.query_pos_embedding.weight, "qp")
        for i in range(self.num_layers):
            li = i % self.num_feature_levels
            ca = self.transformer_crossatt_layers[i]
            lq, lqp = Q[ngq:], P[ngq:]
            kv = fea[li]
            k_in = kv if pos[li] is None else nv.axpby(kv, 1.0, pos[li], 1.0)
            a = self._mha(ca.multihead_attn, nv.axpby(lq, 1.0, lqp, 1.0), k_in, kv, nlq, kv.shape[0])
            g, b = pk_norm(ca.norm)
            nv.layernorm(a, g, b, ca.norm.eps, residual=lq, out=lq)      # lquery <- LN(lquery + attn)
            sa = self.transformer_selfatt_layers[i]
            qk = nv.axpby(Q, 1.0, P, 1.0)
            a = self._mha(sa.self_attn, qk, qk, Q, ngq + nlq, ngq + nlq)
            g, b = pk_norm(sa.norm)
            Q1 = nv.layernorm(a, g, b, sa.norm.eps, residual=Q)
            ff = self.transformer_feedforward_layers[i]
            w1, b1 = pk_lin(ff.linear1)
            w2, b2 = pk_lin(ff.linear2)
            f = nv.linear(nv.linear(Q1, w1, b1, act=nv.ACT_RELU), w2, b2)
            g, b = pk_norm(ff.norm)
            Q = nv.layernorm(f, g, b, ff.norm.eps, residual=Q1)
        return Q.reshape(1, ngq + nlq, E)


class SemanticContextEncoder(nn.Module):
    """seecoder.py:556-578: children are built through the model registry like the reference."""

    def __init__(self, imencoder_cfg, imdecoder_cfg, qtransformer_cfg):
        super().__init__()
        from .registry import get_model
        self.imencoder = get_model()(imencoder_cfg)
        self.imdecoder = get_model()(imdecoder_cfg)
        self.qtransformer = get_model()(qtransformer_cfg)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: [1, 3, R, R] raw [0,1] RGB -> [1, 148, 768] (seecoder.py:567-575)."""
        if x.shape[0] != 1:
            raise NotImplementedError(
                "pfd_b200 SeeCoder encodes one reference image per call (app.py:234-235). The reference "
                "Decoder attends across the batch axis for bs > 1 (seecoder.py:70,83), mixing images; "
                "call ctx_encode per image and torch.cat the results.")
        fea = self.imencoder(x)
        hs = self.imdecoder({k: fea[k] for k in ("res3", "res4", "res5")})
        return self.qtransformer([hs["res3"], hs["res4"], hs["res5"]])

    def encode(self, x):
        return self(x)
