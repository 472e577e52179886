"""CPU oracle for the Prompt-Free-Diffusion hot path  —  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional (state-dict in, tensors out) restatement, in plain torch ops on the CPU, of the
reference algorithm for the path named by BASELINE.json: SeeCoder encode -> CFG DDIM loop over the
SD-v1.5 UNet (+ optional ControlNet residuals) -> AutoKL decode.  Every function cites the
reference file:line it follows.  It is importable on a box that has no copy of the reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; the product (pfd_b200/) never does.

Pinning: the reference ships no tests or golden vectors.  The oracle is pinned by
tools/make_golden.py (+ tools/make_golden_configs.py for the BASELINE config sizes), which import the
unmodified reference from /root/reference in the build container (tools/ref_harness.py), load identical
synthetic weights into both and compare every stage (tests/golden/oracle_pin_report.json); the reference's
outputs are committed under tests/golden/ and re-checked by tests/test_oracle_golden.py on any machine.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# ---------------------------------------------------------------------------------------------
# configurations (configs/model/*.yaml of the reference)
# ---------------------------------------------------------------------------------------------
UNET_SD15 = dict(in_channels=4, out_channels=4, model_channels=320, channel_mult=(1, 2, 4, 4),
                 num_res_blocks=(2, 2, 2, 2), attention_resolutions=(4, 2, 1), num_heads=8,
                 context_dim=768)                                     # openai_unet.yaml:24-35
CONTROLNET_SD15 = dict(UNET_SD15, hint_channels=3)                    # controlnet.yaml:1-18
VAE_SD = dict(ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4, out_ch=3,
              in_channels=3, embed_dim=4)                             # autokl.yaml:5-26
SWIN_L = dict(embed_dim=192, depths=(2, 2, 18, 2), num_heads=(6, 12, 24, 48), window_size=12,
              patch_size=4, in_chans=3, mlp_ratio=4.0)                # swin.yaml:18-29
SEECODER_DECODER = dict(inchannels=dict(res3=384, res4=768, res5=1536), dim=768, nheads=8,
                        ffn=1024, layers=6)                           # seecoder.yaml:25-38
QUERY_TRANSFORMER = dict(hidden=768, nheads=8, layers=9, ffn=2048, num_queries=(4, 144), levels=3,
                         with_pos=False)                              # seecoder.yaml:44-57
PFD = dict(beta_linear_start=0.00085, beta_linear_end=0.012, timesteps=1000,
           latent_scale_factor=0.18215)                               # pfd.yaml:6-8,20-21


def sub(sd: SD, prefix: str) -> SD:
    """View of the entries below `prefix` with the prefix stripped."""
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


# ---------------------------------------------------------------------------------------------
# schedules (pfd.py:110-168, diffusion_utils.py:8-59, ddim.py:23-56)
# ---------------------------------------------------------------------------------------------
def make_beta_schedule_linear(n: int, start: float, end: float) -> np.ndarray:
    """diffusion_utils.py:9-12: linear in sqrt(beta), float64."""
    return (torch.linspace(start ** 0.5, end ** 0.5, n, dtype=torch.float64) ** 2).numpy()


def schedule_buffers(cfg=PFD) -> Dict[str, torch.Tensor]:
    """The 12 persistent fp32 buffers registered by pfd.py:110-160 (v_posterior = 0)."""
    betas = make_beta_schedule_linear(cfg["timesteps"], cfg["beta_linear_start"], cfg["beta_linear_end"])
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return {
        "betas": f32(betas),
        "alphas_cumprod": f32(ac),
        "alphas_cumprod_prev": f32(ac_prev),
        "sqrt_alphas_cumprod": f32(np.sqrt(ac)),
        "sqrt_one_minus_alphas_cumprod": f32(np.sqrt(1.0 - ac)),
        "log_one_minus_alphas_cumprod": f32(np.log(1.0 - ac)),
        "sqrt_recip_alphas_cumprod": f32(np.sqrt(1.0 / ac)),
        "sqrt_recipm1_alphas_cumprod": f32(np.sqrt(1.0 / ac - 1)),
        "posterior_variance": f32(post_var),
        "posterior_log_variance_clipped": f32(np.log(np.maximum(post_var, 1e-20))),
        "posterior_mean_coef1": f32(betas * np.sqrt(ac_prev) / (1.0 - ac)),
        "posterior_mean_coef2": f32((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)),
    }


def ddim_timesteps(num_ddim: int, num_ddpm: int = 1000) -> np.ndarray:
    """diffusion_utils.py:32-46 ('uniform'): stride c = T // steps, then +1 (steps=30 -> 31 entries)."""
    c = num_ddpm // num_ddim
    return np.asarray(list(range(0, num_ddpm, c))) + 1


def ddim_schedule(alphas_cumprod: torch.Tensor, steps: int, eta: float = 0.0):
    """ddim.py:23-56 + diffusion_utils.py:48-59.  `alphas_cumprod` is the model buffer in whatever
    precision the model holds it (fp16 after net.half(): App. C #6) — it is up-cast to fp32 first
    exactly like ddim.py:28-30.  Returns (timesteps, alphas, alphas_prev, sigmas, sqrt_one_minus_alphas)."""
    ts = ddim_timesteps(steps, alphas_cumprod.shape[0])
    ac = alphas_cumprod.clone().detach().to(torch.float32).cpu()
    alphas = ac[ts]
    alphas_prev = np.asarray([ac[0]] + ac[ts[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    sqrt_one_minus = np.sqrt(1.0 - alphas)
    return ts, alphas, alphas_prev, sigmas, sqrt_one_minus


def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """diffusion_utils.py:131-151: [cos | sin] of t * exp(-ln(P) * i / half), fp32."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


# ---------------------------------------------------------------------------------------------
# UNet building blocks (openaimodel.py:162-274, attention.py:44-71,159-201,277-371)
# ---------------------------------------------------------------------------------------------
def _gn(x, sd, name, eps):
    return F.group_norm(x, 32, sd[name + ".weight"], sd[name + ".bias"], eps)


def _conv(x, sd, name, stride=1, padding=1):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)


def _lin(x, sd, name):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _ln(x, sd, name, eps=1e-5):
    w = sd[name + ".weight"]
    return F.layer_norm(x, (w.shape[0],), w, sd[name + ".bias"], eps)


def resblock(sd: SD, p: str, x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """openaimodel.py:254-274 with use_scale_shift_norm=False, no up/down."""
    h = _conv(F.silu(_gn(x, sd, p + "in_layers.0", 1e-5)), sd, p + "in_layers.2")
    e = _lin(F.silu(emb), sd, p + "emb_layers.1").type(h.dtype)
    h = h + e[:, :, None, None]
    h = _conv(F.silu(_gn(h, sd, p + "out_layers.0", 1e-5)), sd, p + "out_layers.3")
    if (p + "skip_connection.weight") in sd:
        x = _conv(x, sd, p + "skip_connection", padding=0)
    return x + h


def cross_attention(sd: SD, p: str, x: torch.Tensor, context: Optional[torch.Tensor], heads: int):
    """attention.py:178-201: bias-free q/k/v, sim = (q k^T) * d^-0.5 in working dtype, softmax(-1)."""
    ctx = x if context is None else context
    q, k, v = _lin(x, sd, p + "to_q"), _lin(ctx, sd, p + "to_k"), _lin(ctx, sd, p + "to_v")
    b, n, c = q.shape
    d = c // heads

    def split(t):
        return t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(b * heads, t.shape[1], d)

    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bid,bjd->bij", q, k) * (d ** -0.5)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bij,bjd->bid", attn, v)
    out = out.reshape(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, c)
    return _lin(out, sd, p + "to_out.0")


def transformer_block(sd: SD, p: str, x: torch.Tensor, context: torch.Tensor, heads: int):
    """attention.py:302-306 (BasicTransformerBlock) with GEGLU feed-forward (attention.py:44-71)."""
    x = cross_attention(sd, p + "attn1.", _ln(x, sd, p + "norm1"), None, heads) + x
    x = cross_attention(sd, p + "attn2.", _ln(x, sd, p + "norm2"), context, heads) + x
    h = _lin(_ln(x, sd, p + "norm3"), sd, p + "ff.net.0.proj")
    a, gate = h.chunk(2, dim=-1)
    x = _lin(a * F.gelu(gate), sd, p + "ff.net.2") + x
    return x


def spatial_transformer(sd: SD, p: str, x: torch.Tensor, context: torch.Tensor, heads: int):
    """attention.py:352-371 (use_linear=False, depth 1, GroupNorm eps 1e-6)."""
    b, c, h, w = x.shape
    x_in = x
    t = _conv(_gn(x, sd, p + "norm", 1e-6), sd, p + "proj_in", padding=0)
    t = t.permute(0, 2, 3, 1).reshape(b, h * w, c)
    t = transformer_block(sd, p + "transformer_blocks.0.", t, context, heads)
    t = t.reshape(b, h, w, c).permute(0, 3, 1, 2)
    return _conv(t, sd, p + "proj_out", padding=0) + x_in


def unet_plan(cfg) -> Tuple[List[dict], List[dict], List[str], List[str], List[str]]:
    """Layer list of UNetModel2D_Next.__init__ (openaimodel.py:2575-2749): data blocks, context
    blocks and the three order lists ('d', 'c', 'save_hidden_feature', 'load_hidden_feature')."""
    mc, mults, nres = cfg["model_channels"], cfg["channel_mult"], cfg["num_res_blocks"]
    attn_res, heads = cfg["attention_resolutions"], cfg["num_heads"]
    data, ctxs, order = [], [], []

    def add_d(b):
        data.append(b)
        order.append("d")

    def add_c(ch):
        ctxs.append(dict(ch=ch, heads=heads, dhead=ch // heads))
        order.append("c")

    add_d(dict(kind="conv", cin=cfg["in_channels"], cout=mc))
    order.append("save_hidden_feature")
    chans = [mc]
    ch, ds = mc, 1
    for level, mult in enumerate(mults):
        for _ in range(nres[level]):
            add_d(dict(kind="res", cin=ch, cout=mult * mc))
            ch = mult * mc
            if ds in attn_res:
                add_c(ch)
            chans.append(ch)
            order.append("save_hidden_feature")
        if level != len(mults) - 1:
            add_d(dict(kind="down", cin=ch, cout=ch))
            chans.append(ch)
            order.append("save_hidden_feature")
            ds *= 2
    i_order, order = order, []
    add_d(dict(kind="res", cin=ch, cout=ch))
    add_c(ch)
    add_d(dict(kind="res", cin=ch, cout=ch))
    m_order, order = order, []
    for level, mult in list(enumerate(mults))[::-1]:
        for _ in range(nres[level] + 1):
            order.append("load_hidden_feature")
            ich = chans.pop()
            add_d(dict(kind="res", cin=ch + ich, cout=mc * mult))
            ch = mc * mult
            if ds in attn_res:
                add_c(ch)
        if level != 0:
            add_d(dict(kind="up", cin=ch, cout=ch))
            ds //= 2
    add_d(dict(kind="out", cin=ch, cout=cfg["out_channels"]))
    return data, ctxs, i_order, m_order, order


def _data_block(sd, p, blk, h, emb):
    k = blk["kind"]
    if k == "conv":
        return _conv(h, sd, p + "0")
    if k == "res":
        return resblock(sd, p + "0.", h, emb)
    if k == "down":
        return _conv(h, sd, p + "0.op", stride=2)                      # openaimodel.py:150,157-159
    if k == "up":
        h = F.interpolate(h, scale_factor=2, mode="nearest")           # openaimodel.py:114
        return _conv(h, sd, p + "0.conv")
    if k == "out":                                                      # openaimodel.py:2732-2736
        return _conv(F.silu(_gn(h, sd, p + "0.0", 1e-5)), sd, p + "0.2")
    raise ValueError(k)


def unet_apply(sd: SD, cfg, x: torch.Tensor, t: torch.Tensor, context: torch.Tensor,
               control: Optional[List[torch.Tensor]] = None,
               mixed: Optional[Sequence[Tuple[torch.Tensor, float]]] = None) -> torch.Tensor:
    """pfd.py:314-365 / 466-528: walk i/m/o orders; `control` = ControlNet outputs (popped from the end).
    mixed = [(context_i, ratio_i)]: pfd.py:367-439 apply_model_multicontext with 'attention' mixing - every context
    block output is sum_i block(h, context_i) * (ratio_i / sum ratio) (pfd.py:374-379)."""
    data, ctxs, i_order, m_order, o_order = unet_plan(cfg)
    dtype = x.dtype
    t_emb = timestep_embedding(t, cfg["model_channels"]).to(dtype)      # pfd.py:486
    emb = _lin(F.silu(_lin(t_emb, sd, "time_embed.0")), sd, "time_embed.2")
    ccs = list(control) if control is not None else None
    di, ci = 0, 0
    hs = []
    h = x

    def step(ltype):
        nonlocal di, ci, h
        if ltype == "d":
            h = _data_block(sd, f"data_blocks.{di}.", data[di], h, emb)
            di += 1
        elif ltype == "c":
            c = ctxs[ci]
            if mixed is None:
                h = spatial_transformer(sd, f"context_blocks.{ci}.0.", h, context, c["heads"])
            else:
                rs = np.array([r for _, r in mixed], dtype=np.float64)
                rs = rs / rs.sum()
                acc = None
                for (cm, _), r in zip(mixed, rs):
                    hi = spatial_transformer(sd, f"context_blocks.{ci}.0.", h, cm, c["heads"]) * r
                    acc = hi if acc is None else acc + hi
                h = acc
            ci += 1

    for lt in i_order:
        if lt == "save_hidden_feature":
            hs.append(h)
        else:
            step(lt)
    for lt in m_order:
        step(lt)
    if ccs is not None:
        h = h + ccs.pop()                                               # pfd.py:515
    for lt in o_order:
        if lt == "load_hidden_feature":
            skip = hs.pop()
            if ccs is not None:
                skip = skip + ccs.pop()                                 # pfd.py:519
            h = torch.cat([h, skip], dim=1)
        else:
            step(lt)
    return h


# ---------------------------------------------------------------------------------------------
# ControlNet (controlnet.py:66-324)
# ---------------------------------------------------------------------------------------------
HINT_STEM = [(16, 1), (16, 1), (32, 2), (32, 1), (96, 2), (96, 1), (256, 2)]  # controlnet.py:165-180


def controlnet_plan(cfg) -> List[List[dict]]:
    """input_blocks of ControlNet.__init__ (controlnet.py:154-262): list of per-block layer lists."""
    mc, mults, nres = cfg["model_channels"], cfg["channel_mult"], cfg["num_res_blocks"]
    heads, attn_res = cfg["num_heads"], cfg["attention_resolutions"]
    blocks = [[dict(kind="conv", cin=cfg["in_channels"], cout=mc)]]
    ch, ds = mc, 1
    for level, mult in enumerate(mults):
        for _ in range(nres[level]):
            layers = [dict(kind="res", cin=ch, cout=mult * mc)]
            ch = mult * mc
            if ds in attn_res:
                layers.append(dict(kind="attn", ch=ch, heads=heads))
            blocks.append(layers)
        if level != len(mults) - 1:
            blocks.append([dict(kind="down", cin=ch, cout=ch)])
            ds *= 2
    return blocks


def controlnet_hint(sd: SD, cfg, hint: torch.Tensor) -> torch.Tensor:
    """input_hint_block (controlnet.py:165-181): 7 conv+SiLU, then a (zero-init) conv to model_channels."""
    h = hint
    for i, (_, s) in enumerate(HINT_STEM):
        h = F.silu(_conv(h, sd, f"input_hint_block.{2 * i}", stride=s))
    return _conv(h, sd, f"input_hint_block.{2 * len(HINT_STEM)}")


def controlnet_apply(sd: SD, cfg, x, hint, t, context) -> List[torch.Tensor]:
    """controlnet.py:302-324 -> 13 residuals (12 zero-conv outputs + middle_block_out)."""
    t_emb = timestep_embedding(t, cfg["model_channels"]).to(x.dtype)
    emb = _lin(F.silu(_lin(t_emb, sd, "time_embed.0")), sd, "time_embed.2")
    guided = controlnet_hint(sd, cfg, hint)
    outs = []
    h = x
    heads = cfg["num_heads"]
    for bi, layers in enumerate(controlnet_plan(cfg)):
        for li, l in enumerate(layers):
            p = f"input_blocks.{bi}.{li}."
            if l["kind"] == "conv":
                h = _conv(h, sd, p[:-1])
            elif l["kind"] == "res":
                h = resblock(sd, p, h, emb)
            elif l["kind"] == "attn":
                h = spatial_transformer(sd, p, h, context, heads)
            elif l["kind"] == "down":
                h = _conv(h, sd, p + "op", stride=2)
        if guided is not None:
            h = h + guided                                               # controlnet.py:315 (broadcast over batch)
            guided = None
        outs.append(_conv(h, sd, f"zero_convs.{bi}.0", padding=0))
    h = resblock(sd, "middle_block.0.", h, emb)
    h = spatial_transformer(sd, "middle_block.1.", h, context, heads)
    h = resblock(sd, "middle_block.2.", h, emb)
    outs.append(_conv(h, sd, "middle_block_out.0", padding=0))
    return outs


# ---------------------------------------------------------------------------------------------
# AutoKL decoder (autokl.py:44-54, autokl_modules.py:82-202, 462-568)
# ---------------------------------------------------------------------------------------------
def vae_resnet(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """autokl_modules.py:121-141 with temb=None, GroupNorm eps 1e-6."""
    h = _conv(F.silu(_gn(x, sd, p + "norm1", 1e-6)), sd, p + "conv1")
    h = _conv(F.silu(_gn(h, sd, p + "norm2", 1e-6)), sd, p + "conv2")
    if (p + "nin_shortcut.weight") in sd:
        x = _conv(x, sd, p + "nin_shortcut", padding=0)
    return x + h


def vae_attn(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """autokl_modules.py:178-202: single head, scale c^-0.5, softmax over keys."""
    h_ = _gn(x, sd, p + "norm", 1e-6)
    q, k, v = (_conv(h_, sd, p + n, padding=0) for n in ("q", "k", "v"))
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + _conv(h_, sd, p + "proj_out", padding=0)


def vae_decode(sd: SD, cfg, z: torch.Tensor, scale_factor: Optional[float] = PFD["latent_scale_factor"]):
    """pfd.py:275-282 + autokl.py:44-54 + Decoder.forward (autokl_modules.py:535-568); sd = 'vae.image.*'."""
    if scale_factor is not None:
        z = 1.0 / scale_factor * z
    h = _conv(z, sd, "post_quant_conv", padding=0)
    h = _conv(h, sd, "decoder.conv_in")
    h = vae_resnet(sd, "decoder.mid.block_1.", h)
    h = vae_attn(sd, "decoder.mid.attn_1.", h)
    h = vae_resnet(sd, "decoder.mid.block_2.", h)
    nlev = len(cfg["ch_mult"])
    for lvl in reversed(range(nlev)):
        for bi in range(cfg["num_res_blocks"] + 1):
            h = vae_resnet(sd, f"decoder.up.{lvl}.block.{bi}.", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(h, sd, f"decoder.up.{lvl}.upsample.conv")
    h = _conv(F.silu(_gn(h, sd, "decoder.norm_out", 1e-6)), sd, "decoder.conv_out")
    return torch.clamp((h + 1) / 2, 0, 1)


def vae_encode_moments(sd: SD, cfg, x: torch.Tensor):
    """autokl.py:33-42 + Encoder.forward (autokl_modules.py:436-459) + Downsample (autokl_modules.py:69-76:
    F.pad(x, (0,1,0,1)) then a stride-2 conv with padding 0) + DiagonalGaussianDistribution.__init__
    (distributions.py:24-31).  x: [B,3,H,W] in [0,1]; returns (mean, logvar clamped to [-30, 20]); sd = 'vae.image.*'."""
    h = _conv(x * 2 - 1, sd, "encoder.conv_in")
    nlev = len(cfg["ch_mult"])
    for lvl in range(nlev):
        for bi in range(cfg["num_res_blocks"]):
            h = vae_resnet(sd, f"encoder.down.{lvl}.block.{bi}.", h)
        if lvl != nlev - 1:
            h = _conv(F.pad(h, (0, 1, 0, 1), mode="constant", value=0), sd, f"encoder.down.{lvl}.downsample.conv",
                      stride=2, padding=0)
    h = vae_resnet(sd, "encoder.mid.block_1.", h)
    h = vae_attn(sd, "encoder.mid.attn_1.", h)
    h = vae_resnet(sd, "encoder.mid.block_2.", h)
    h = _conv(F.silu(_gn(h, sd, "encoder.norm_out", 1e-6)), sd, "encoder.conv_out")
    mean, logvar = torch.chunk(_conv(h, sd, "quant_conv", padding=0), 2, dim=1)
    return mean, torch.clamp(logvar, -30.0, 20.0)


def vae_encode(sd: SD, cfg, x: torch.Tensor, noise: torch.Tensor,
               scale_factor: Optional[float] = PFD["latent_scale_factor"]) -> torch.Tensor:
    """pfd.py:266-273: scale * (mean + exp(0.5*logvar) * noise) with the caller-supplied posterior noise
    (the reference draws torch.randn(mean.shape) on the CPU, distributions.py:36)."""
    mean, logvar = vae_encode_moments(sd, cfg, x)
    z = mean + torch.exp(0.5 * logvar) * noise
    return z * scale_factor if scale_factor is not None else z


# ---------------------------------------------------------------------------------------------
# Swin-L backbone (swin.py)
# ---------------------------------------------------------------------------------------------
def relative_position_index(ws: int) -> torch.Tensor:
    """swin.py:159-169."""
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij"))
    flat = torch.flatten(coords, 1)
    rel = (flat[:, :, None] - flat[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def _win_partition(x, ws):
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def _win_reverse(win, ws, H, W):
    B = int(win.shape[0] / (H * W / ws / ws))
    x = win.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def swin_shift_mask(H: int, W: int, ws: int, shift: int, dtype) -> torch.Tensor:
    """swin.py:421-440: region ids on the padded map -> (0 / -100) mask per window."""
    Hp, Wp = int(np.ceil(H / ws)) * ws, int(np.ceil(W / ws)) * ws
    img = torch.zeros((1, Hp, Wp, 1), dtype=dtype)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    mw = _win_partition(img, ws).view(-1, ws * ws)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, float(-100.0)).masked_fill(am == 0, float(0.0))


def swin_block(sd: SD, p: str, x: torch.Tensor, H: int, W: int, ws: int, shift: int, heads: int,
               mask: torch.Tensor, rpi: torch.Tensor) -> torch.Tensor:
    """swin.py:254-310 (block) + 179-210 (window attention)."""
    B, L, C = x.shape
    shortcut = x
    x = _ln(x, sd, p + "norm1").view(B, H, W, C)
    pad_r, pad_b = (ws - W % ws) % ws, (ws - H % ws) % ws
    x = F.pad(x, (0, 0, 0, pad_r, 0, pad_b))                           # pad AFTER norm1: zeros enter qkv
    Hp, Wp = x.shape[1], x.shape[2]
    if shift > 0:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
    xw = _win_partition(x, ws).view(-1, ws * ws, C)
    Bw, N, _ = xw.shape
    d = C // heads
    qkv = _lin(xw, sd, p + "attn.qkv").reshape(Bw, N, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (d ** -0.5), qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    bias = sd[p + "attn.relative_position_bias_table"][rpi.view(-1)].view(N, N, -1).permute(2, 0, 1).contiguous()
    attn = attn + bias.unsqueeze(0)
    if shift > 0:
        nW = mask.shape[0]
        attn = attn.view(Bw // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, heads, N, N)
    attn = attn.softmax(dim=-1)
    xo = (attn @ v).transpose(1, 2).reshape(Bw, N, C)
    xo = _lin(xo, sd, p + "attn.proj").view(-1, ws, ws, C)
    x = _win_reverse(xo, ws, Hp, Wp)
    if shift > 0:
        x = torch.roll(x, shifts=(shift, shift), dims=(1, 2))
    if pad_r > 0 or pad_b > 0:
        x = x[:, :H, :W, :].contiguous()
    x = shortcut + x.view(B, H * W, C)
    h = _lin(F.gelu(_lin(_ln(x, sd, p + "norm2"), sd, p + "mlp.fc1")), sd, p + "mlp.fc2")
    return x + h


def swin_patch_merge(sd: SD, p: str, x: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """swin.py:325-351."""
    B, L, C = x.shape
    x = x.view(B, H, W, C)
    if (H % 2 == 1) or (W % 2 == 1):
        x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
    x = x.view(B, -1, 4 * C)
    return F.linear(_ln(x, sd, p + "norm"), sd[p + "reduction.weight"])


def swin_forward(sd: SD, cfg, img: torch.Tensor) -> Dict[str, torch.Tensor]:
    """swin.py:623-653 (+ PatchEmbed 479-495, BasicLayer 414-453): returns res2..res5 NCHW."""
    ps, ws = cfg["patch_size"], cfg["window_size"]
    _, _, H, W = img.shape
    if W % ps:
        img = F.pad(img, (0, ps - W % ps))
    if H % ps:
        img = F.pad(img, (0, 0, 0, ps - H % ps))
    x = F.conv2d(img, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=ps)
    Wh, Ww = x.shape[2], x.shape[3]
    x = _ln(x.flatten(2).transpose(1, 2), sd, "patch_embed.norm")
    rpi = relative_position_index(ws).to(x.device)
    outs = {}
    nl = len(cfg["depths"])
    for i in range(nl):
        C = cfg["embed_dim"] * 2 ** i
        mask = swin_shift_mask(Wh, Ww, ws, ws // 2, x.dtype).to(x.device)
        for j in range(cfg["depths"][i]):
            x = swin_block(sd, f"layers.{i}.blocks.{j}.", x, Wh, Ww, ws, 0 if j % 2 == 0 else ws // 2,
                           cfg["num_heads"][i], mask, rpi)
        xo = _ln(x, sd, f"norm{i}")
        outs[f"res{i + 2}"] = xo.view(-1, Wh, Ww, C).permute(0, 3, 1, 2).contiguous()
        if i < nl - 1:
            x = swin_patch_merge(sd, f"layers.{i}.downsample.", x, Wh, Ww)
            Wh, Ww = (Wh + 1) // 2, (Ww + 1) // 2
    return outs


# ---------------------------------------------------------------------------------------------
# SeeCoder decoder + query transformer (seecoder.py)
# ---------------------------------------------------------------------------------------------
def _mha(sd: SD, p: str, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int) -> torch.Tensor:
    """nn.MultiheadAttention forward (batch_first=False): q [Lq,B,E], k/v [Lk,B,E]."""
    E = q.shape[-1]
    w, b = sd[p + "in_proj_weight"], sd[p + "in_proj_bias"]
    out, _ = F.multi_head_attention_forward(
        q, k, v, E, heads, w, b, None, None, False, 0.0, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"],
        training=False, need_weights=False)
    return out


def seecoder_decoder(sd: SD, cfg, feats: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """seecoder.py:394-428.  NOTE the batch-axis quirk (App. C #1): tokens [bs, L, C] are fed to a
    seq-first MultiheadAttention, so attention runs over the *batch* axis (seecoder.py:70,83)."""
    tags = sorted(cfg["inchannels"].keys())
    xs, shapes = [], {}
    for idx, tag in enumerate(tags[::-1]):
        xi = _conv(feats[tag], sd, f"inproj_layers.{tag}.0", padding=0)
        xi = F.group_norm(xi, 32, sd[f"inproj_layers.{tag}.1.weight"], sd[f"inproj_layers.{tag}.1.bias"], 1e-5)
        bs, _, h, w = xi.shape
        shapes[tag] = (h, w)
        xs.append(xi.flatten(2).transpose(1, 2) + sd["level_embed"][idx].view(1, 1, -1))
    lens = [t.shape[1] for t in xs]
    h = torch.cat(xs, 1)
    for l in range(cfg["layers"]):
        p = f"transformer.layers.{l}."
        h = _ln(h + _mha(sd, p + "self_attn.", h, h, h, cfg["nheads"]), sd, p + "norm1")
        h2 = _lin(F.relu(_lin(h, sd, p + "linear1")), sd, p + "linear2")
        h = _ln(h + h2, sd, p + "norm2")
    ys = torch.split(h, lens, dim=1)
    out = {}
    for idx, tag in enumerate(tags[::-1]):
        hh, ww = shapes[tag]
        out[tag] = ys[idx].transpose(1, 2).reshape(bs, -1, hh, ww)
    for tag in tags[::-1]:
        lat = F.conv2d(feats[tag], sd[f"lateral_layers.{tag}.weight"], None)
        lat = F.group_norm(lat, 32, sd[f"lateral_layers.{tag}.norm.weight"], sd[f"lateral_layers.{tag}.norm.bias"], 1e-5)
        out[tag] = out[tag] + lat
    return out


def ppe_mlp(sd: SD, p: str, x: torch.Tensor, freq_num: int = 20) -> torch.Tensor:
    """PPE_MLP.forward in eval mode (seecoder.py:285-310): [1, C, h, w] positional map."""
    h, w = x.shape[-2:]
    minlen = min(h, w)
    he, we = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    he = ((he + 0.5 - h / 2) / minlen * (2 * math.pi)).to(x.device).to(x.dtype)
    we = ((we + 0.5 - w / 2) / minlen * (2 * math.pi)).to(x.device).to(x.dtype)
    dim_t = torch.linspace(0, 1, freq_num, dtype=torch.float32, device=x.device)
    dim_t = (minlen / 2) ** dim_t.to(x.dtype)
    ph, pw = he[:, :, None] * dim_t, we[:, :, None] * dim_t
    pos = torch.cat((ph.sin(), ph.cos(), pw.sin(), pw.cos()), dim=-1)
    pos = _lin(F.silu(_lin(F.silu(_lin(pos, sd, p + "mlp.0")), sd, p + "mlp.2")), sd, p + "mlp.4")
    return pos.permute(2, 0, 1)[None]


def query_transformer(sd: SD, cfg, feats: Sequence[torch.Tensor]) -> torch.Tensor:
    """seecoder.py:500-550: 9 x [cross-attn(local queries -> level i%3), self-attn(148), FFN], post-norm."""
    heads = cfg["nheads"]
    with_pos = "pe_layer.mlp.0.weight" in sd
    fea, pos = [], []
    for i in range(cfg["levels"]):
        pi = ppe_mlp(sd, "pe_layer.", feats[i]).flatten(2).transpose(1, 2) if with_pos else None
        xi = feats[i].flatten(2) + sd["level_embed.weight"][i][None, :, None]
        fea.append(xi.transpose(1, 2))
        pos.append(pi)
    bs = fea[0].shape[0]
    ngq, nlq = cfg["num_queries"]
    iq, qp = sd["init_query.weight"], sd["query_pos_embedding.weight"]
    gq, lq = iq[:ngq].unsqueeze(0).repeat(bs, 1, 1), iq[ngq:].unsqueeze(0).repeat(bs, 1, 1)
    gqp, lqp = qp[:ngq].unsqueeze(0).repeat(bs, 1, 1), qp[ngq:].unsqueeze(0).repeat(bs, 1, 1)
    for i in range(cfg["layers"]):
        lvl = i % cfg["levels"]
        p = f"transformer_crossatt_layers.{i}."
        kv = fea[lvl]
        kk = kv if pos[lvl] is None else kv + pos[lvl]
        h1 = _mha(sd, p + "multihead_attn.", (lq + lqp).transpose(0, 1), kk.transpose(0, 1), kv.transpose(0, 1), heads)
        lq = _ln(lq + h1.transpose(0, 1), sd, p + "norm")
        p = f"transformer_selfatt_layers.{i}."
        qkv = torch.cat([gq, lq], 1)
        qk = (qkv + torch.cat([gqp, lqp], 1)).transpose(0, 1)
        h1 = _mha(sd, p + "self_attn.", qk, qk, qkv.transpose(0, 1), heads)
        q = _ln(qkv + h1.transpose(0, 1), sd, p + "norm")
        p = f"transformer_feedforward_layers.{i}."
        q = _ln(q + _lin(F.relu(_lin(q, sd, p + "linear1")), sd, p + "linear2"), sd, p + "norm")
        gq, lq = q[:, :ngq], q[:, ngq:]
    return torch.cat([gq, lq], 1)


def seecoder_encode(sd: SD, img: torch.Tensor, swin_cfg=SWIN_L, dec_cfg=SEECODER_DECODER,
                    qt_cfg=QUERY_TRANSFORMER) -> torch.Tensor:
    """seecoder.py:567-575: raw [0,1] RGB -> [B, 148, 768]; sd = 'ctx.image.*'."""
    fea = swin_forward(sub(sd, "imencoder."), swin_cfg, img)
    hs = seecoder_decoder(sub(sd, "imdecoder."), dec_cfg, {k: fea[k] for k in ("res3", "res4", "res5")})
    return query_transformer(sub(sd, "qtransformer."), qt_cfg, [hs["res3"], hs["res4"], hs["res5"]])


# ---------------------------------------------------------------------------------------------
# DDIM sampler (ddim.py:58-172)
# ---------------------------------------------------------------------------------------------
def ddim_update(x, e_t, a_t, a_prev, sigma_t, sqrt_one_minus_at, noise=None, temperature=1.0):
    """ddim.py:159-171; coefficients are torch.full(..., dtype=x.dtype) tensors exactly as in the reference.
    noise=None: the eta = 0 case (sigma_t * noise == 0)."""
    b = x.shape[0]
    ext = [b] + [1] * (x.dim() - 1)
    mk = lambda v: torch.full(ext, float(v), dtype=x.dtype, device=x.device)
    a_t, a_prev, sigma_t, s1m = mk(a_t), mk(a_prev), mk(sigma_t), mk(sqrt_one_minus_at)
    pred_x0 = (x - s1m * e_t) / a_t.sqrt()
    dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e_t
    if noise is not None:
        return a_prev.sqrt() * pred_x0 + dir_xt + sigma_t * noise * temperature, pred_x0
    return a_prev.sqrt() * pred_x0 + dir_xt, pred_x0


def ddim_sample(unet_sd: SD, unet_cfg, alphas_cumprod: torch.Tensor, *, steps: int, x_T: torch.Tensor,
                cond: torch.Tensor, uncond: Optional[torch.Tensor], guidance: float,
                ctl_sd: Optional[SD] = None, ctl_cfg=None, hint: Optional[torch.Tensor] = None,
                trace: Optional[list] = None, max_evals: Optional[int] = None,
                n_forward: Optional[int] = None, eta: float = 0.0, noises: Optional[Sequence[torch.Tensor]] = None,
                temperature: float = 1.0, mixed: Optional[Sequence[Tuple[torch.Tensor, Optional[torch.Tensor], float]]] = None
                ) -> torch.Tensor:
    """ddim.py:81-172 with x_T supplied (the reference draws it with torch.randn, :105).
    n_forward: the img2img branch (ddim.py:94-101) - x_T is x0 already noised to timestep ts[n_forward] and only
    the first n_forward timesteps are walked.  eta > 0: `noises[i]` is the noise_like() draw of step i
    (ddim.py:168).  mixed = [(cond_i, uncond_i, ratio_i)]: sample_multicontext (ddim.py:174-299); `cond` /
    `uncond` are then ignored."""
    ts, alphas, alphas_prev, sigmas, s1m = ddim_schedule(alphas_cumprod, steps, eta)
    if n_forward is not None:
        ts = ts[:n_forward]
    x = x_T
    b = x.shape[0]
    total = ts.shape[0]
    for i, step in enumerate(np.flip(ts)):
        if max_evals is not None and i >= max_evals:
            break
        index = total - i - 1
        t = torch.full((b,), int(step), dtype=torch.long, device=x.device)
        if mixed is not None:
            if guidance == 1.0:
                e_t = unet_apply(unet_sd, unet_cfg, x, t, None, mixed=[(c, r) for c, _, r in mixed])
            else:
                x_in, t_in = torch.cat([x] * 2), torch.cat([t] * 2)
                e_u, e_c = unet_apply(unet_sd, unet_cfg, x_in, t_in, None,
                                      mixed=[(torch.cat([u, c]), r) for c, u, r in mixed]).chunk(2)
                e_t = e_u + guidance * (e_c - e_u)
        elif guidance == 1.0 or uncond is None:
            control = controlnet_apply(ctl_sd, ctl_cfg, x, hint, t, cond) if hint is not None else None
            e_t = unet_apply(unet_sd, unet_cfg, x, t, cond, control) * guidance
        else:
            x_in, t_in = torch.cat([x] * 2), torch.cat([t] * 2)
            c_in = torch.cat([uncond, cond])
            control = controlnet_apply(ctl_sd, ctl_cfg, x_in, hint, t_in, c_in) if hint is not None else None
            e_u, e_c = unet_apply(unet_sd, unet_cfg, x_in, t_in, c_in, control).chunk(2)
            e_t = e_u + guidance * (e_c - e_u)
        if trace is not None:
            trace.append(dict(x=x.clone(), e_t=e_t.clone(), t=int(step)))
        nz = noises[i] if (noises is not None and eta != 0.0) else None
        x, _ = ddim_update(x, e_t, alphas[index], alphas_prev[index], sigmas[index], s1m[index], nz, temperature)
    return x
