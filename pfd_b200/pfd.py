"""PromptFreeDiffusion pipeline composite — mirrors lib/model_zoo/pfd.py:28-168, 266-289, 314-365, 458-528.

Keeps the attributes / methods app.py and DDIMSampler rely on: `vae`, `ctx`, `diffuser`
ModuleDicts, the 12 fp32 schedule buffers, `to(device)` that records `self.device` and returns None,
`ctx_encode`, `vae_decode`, `apply_model(x_info, timesteps, c_info)`.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import native as nv
from .graphs import GraphCache, GraphedFunction, weights_signature
from .registry import get_model


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    """diffusion_utils.py:8-30 ('linear' is the only schedule the pfd configs use)."""
    if schedule != "linear":
        raise NotImplementedError(f"beta schedule '{schedule}' is not used by the pfd configs")
    return (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64, device="cpu") ** 2).numpy()


class PromptFreeDiffusion(nn.Module):
    def __init__(self, vae_cfg_list, ctx_cfg_list, diffuser_cfg_list, global_layer_ptr=None,
                 parameterization="eps", timesteps=1000, use_ema=False, beta_schedule="linear",
                 beta_linear_start=1e-4, beta_linear_end=2e-2, given_betas=None, cosine_s=8e-3,
                 loss_type="l2", l_simple_weight=1.0, l_elbo_weight=0.0, v_posterior=0.0,
                 learn_logvar=False, logvar_init=0, latent_scale_factor=None):
        super().__init__()
        assert parameterization in ["eps", "x0"]
        if use_ema:
            raise NotImplementedError("EMA weights are a training feature (pfd.yaml: use_ema false)")
        self.parameterization = parameterization
        self.vae = self.get_model_list(vae_cfg_list)
        self.ctx = self.get_model_list(ctx_cfg_list)
        self.diffuser = self.get_model_list(diffuser_cfg_list)
        self.global_layer_ptr = global_layer_ptr
        self.use_ema = use_ema
        self.v_posterior = v_posterior
        self.register_schedule(given_betas, beta_schedule, timesteps, beta_linear_start, beta_linear_end, cosine_s)
        self.latent_scale_factor = {} if latent_scale_factor is None else dict(latent_scale_factor)
        self.parameter_group = {}
        for n, d in self.diffuser.items():
            self.parameter_group.update({f"diffuser_{n}_{k}": v for k, v in d.parameter_group.items()})
        self._hint_cache = None
        self.use_cuda_graphs = True          # replay captured graphs for ctx_encode / vae_decode (see graphs.py)
        self._graphs = GraphCache(max_entries=6)

    def to(self, device):
        self.device = device
        super().to(device)

    def get_model_list(self, cfg_list):
        net = nn.ModuleDict()
        for name, cfg in cfg_list:
            net[name] = get_model()(cfg)
        return net

    def register_schedule(self, given_betas=None, beta_schedule="linear", timesteps=1000, linear_start=1e-4,
                          linear_end=2e-2, cosine_s=8e-3):
        """pfd.py:110-168: float64 numpy schedule -> 12 persistent fp32 buffers (+ lvlb_weights)."""
        betas = given_betas if given_betas is not None else make_beta_schedule(
            beta_schedule, timesteps, linear_start, linear_end, cosine_s)
        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1.0, ac[:-1])
        self.num_timesteps = int(betas.shape[0])
        self.linear_start, self.linear_end = linear_start, linear_end
        f32 = lambda a: torch.tensor(a, dtype=torch.float32, device="cpu")
        v = self.v_posterior
        post_var = (1 - v) * betas * (1.0 - ac_prev) / (1.0 - ac) + v * betas
        for name, val in [
            ("betas", betas), ("alphas_cumprod", ac), ("alphas_cumprod_prev", ac_prev),
            ("sqrt_alphas_cumprod", np.sqrt(ac)), ("sqrt_one_minus_alphas_cumprod", np.sqrt(1.0 - ac)),
            ("log_one_minus_alphas_cumprod", np.log(1.0 - ac)), ("sqrt_recip_alphas_cumprod", np.sqrt(1.0 / ac)),
            ("sqrt_recipm1_alphas_cumprod", np.sqrt(1.0 / ac - 1)), ("posterior_variance", post_var),
            ("posterior_log_variance_clipped", np.log(np.maximum(post_var, 1e-20))),
            ("posterior_mean_coef1", betas * np.sqrt(ac_prev) / (1.0 - ac)),
            ("posterior_mean_coef2", (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)),
        ]:
            self.register_buffer(name, f32(val))
        lvlb = self.betas ** 2 / (2 * self.posterior_variance * f32(alphas) * (1 - self.alphas_cumprod))
        lvlb[0] = lvlb[1]
        self.register_buffer("lvlb_weights", lvlb, persistent=False)

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def vae_decode(self, z, which, **kwargs):
        """pfd.py:275-282: z / scale -> AutoencoderKL.decode -> [B,3,H,W] in [0,1]."""
        scale = self.latent_scale_factor.get(which, None) if self.latent_scale_factor is not None else None
        pre = (1.0 / scale) if scale is not None else 1.0
        vae = self.vae[which]
        if not (self.use_cuda_graphs and z.is_cuda) or kwargs:
            return vae.decode(z, pre_scale=pre, **kwargs)
        z = z.to(torch.float16)
        key = ("vae", which, tuple(z.shape), pre, weights_signature(vae.decoder, vae.post_quant_conv))
        gf, hit = self._graphs.get(key, lambda: GraphedFunction(lambda t: vae.decode(t, pre_scale=pre), [z]))
        return (gf(z) if hit else gf.first_out).clone()

    @torch.no_grad()
    def vae_encode(self, x, which, **kwargs):
        """pfd.py:266-273: AutoencoderKL.encode -> scale * z."""
        scale = self.latent_scale_factor.get(which, None) if self.latent_scale_factor is not None else None
        if kwargs.get("out_posterior", False):
            return self.vae[which].encode(x, **kwargs)
        return self.vae[which].encode(x, post_scale=1.0 if scale is None else float(scale), **kwargs)

    @torch.no_grad()
    def q_sample(self, x_start, t, noise=None):
        """pfd.py:204-207: sqrt(acp[t]) * x0 + sqrt(1-acp[t]) * noise (per-sample timestep tensor t)."""
        noise = torch.randn_like(x_start) if noise is None else noise
        x0 = x_start.to(torch.float16).contiguous()
        nz = noise.to(torch.float16).contiguous()
        out = torch.empty_like(x0)
        tl = [int(v) for v in t.reshape(-1).tolist()]
        for i, ti in enumerate(tl):
            nv.axpby(x0[i], float(self.sqrt_alphas_cumprod[ti]), nz[i], float(self.sqrt_one_minus_alphas_cumprod[ti]),
                     out=out[i])
        return out

    @torch.no_grad()
    def ctx_encode(self, x, which, **kwargs):
        """pfd.py:284-289."""
        if which.find("vae_") == 0:
            return self.vae[which[4:]].encode(x, **kwargs)
        enc = self.ctx[which]
        if not (self.use_cuda_graphs and x.is_cuda) or kwargs:
            return enc.encode(x, **kwargs)
        if x.dtype not in (torch.float16, torch.float32):
            x = x.to(torch.float16)
        key = ("ctx", which, tuple(x.shape), x.dtype, weights_signature(enc))
        gf, hit = self._graphs.get(key, lambda: GraphedFunction(lambda t: enc.encode(t), [x]))
        return (gf(x) if hit else gf.first_out).clone()

    def check_diffuser(self):
        orders = [d.layer_order for d in self.diffuser.values()]
        return all(o == orders[0] for o in orders)

    def prepare_context(self, c: torch.Tensor, c_type: str = "image") -> Dict:
        """Step-invariant cross-attention K/V of every transformer block for context `c` (the sampler
        calls this once per `sample`; the reference recomputes to_k/to_v every step)."""
        c = c.to(torch.float16).contiguous()
        return {"c": c, "unet": self.diffuser[c_type].prepare_context(c)}

    @torch.no_grad()
    def apply_model(self, x_info, timesteps, c_info):
        """pfd.py:314-365: one UNet evaluation.  Returns NCHW eps in fp16."""
        x_type, x = x_info["type"], x_info["x"]
        c_type, c = c_info["type"], c_info["c"]
        prep = c_info.get("_pfd_prepared", None)
        kv = prep["unet"] if (prep is not None and prep["c"].data_ptr() == c.data_ptr()) else None
        gl = x_type if self.global_layer_ptr is None else self.global_layer_ptr
        assert gl == x_type == c_type, "pfd_b200 runs single-modality pipelines (image/image)"
        return self.diffuser[x_type].apply(x, timesteps, c, control=None, kv=kv)

    @torch.no_grad()
    def apply_model_multicontext(self, x_info, timesteps, c_info_list, mixing_type="attention"):
        """pfd.py:367-439: every context block output is the ratio-weighted sum over the contexts
        (`context_mixing`, 'attention' mixing; the stochastic 'layer' mixing is not used by the sampler)."""
        if mixing_type != "attention":
            raise NotImplementedError("only 'attention' context mixing is used (ddim.py:273,278)")
        x_type, x = x_info["type"], x_info["x"]
        ratios = np.array([float(c["ratio"]) for c in c_info_list], dtype=np.float64)
        ratios = ratios / ratios.sum()
        for c in c_info_list:
            assert c["type"] == x_type, "pfd_b200 runs single-modality pipelines (image/image)"
        ctxs = [(c["c"].to(torch.float16).contiguous(), float(r)) for c, r in zip(c_info_list, ratios)]
        return self.diffuser[x_type].apply(x, timesteps, ctxs[0][0], mixed_contexts=ctxs)

    def get_device(self):
        return next(self.parameters()).device

    def get_dtype(self):
        return next(self.parameters()).dtype


class PromptFreeDiffusion_with_control(PromptFreeDiffusion):
    """pfd.py:458-528."""

    def __init__(self, *args, **kwargs):
        ctl_cfg = kwargs.pop("ctl_cfg")
        super().__init__(*args, **kwargs)
        self.ctl = get_model()(ctl_cfg)
        self.control_scales = [1.0] * 13
        self.parameter_group["ctl"] = [self.ctl]

    def prepare_context(self, c, c_type="image"):
        prep = super().prepare_context(c, c_type)
        prep["ctl"] = self.ctl.prepare_context(prep["c"])
        return prep

    def _hint(self, cc: torch.Tensor) -> torch.Tensor:
        key = (cc.data_ptr(), cc._version, tuple(cc.shape))
        if self._hint_cache is None or self._hint_cache[0] != key:
            self._hint_cache = (key, self.ctl.hint_features(cc))
        return self._hint_cache[1]

    @torch.no_grad()
    def apply_model(self, x_info, timesteps, c_info):
        x_type, x = x_info["type"], x_info["x"]
        c_type, c = c_info["type"], c_info["c"]
        cc = c_info.get("control", None)
        prep = c_info.get("_pfd_prepared", None)
        if prep is not None and prep["c"].data_ptr() != c.data_ptr():
            prep = None
        control = None
        if cc is not None:
            hint_feat = prep.get("hint") if prep is not None else None
            control = self.ctl(x, hint=cc, timesteps=timesteps, context=c,
                               kv=prep["ctl"] if prep is not None else None,
                               hint_feat=hint_feat if hint_feat is not None else self._hint(cc))
        return self.diffuser[x_type].apply(x, timesteps, c, control=control,
                                           kv=prep["unet"] if prep is not None else None)
