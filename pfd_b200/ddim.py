"""DDIM sampler — mirrors lib/model_zoo/ddim.py:9-172 (`DDIMSampler(net).sample(steps, shape, x_info,
c_info, eta, ...) -> (x, intermediates)`), with the schedule maths reproduced operation for
operation on the host (including the fp16-rounded `alphas_cumprod` after `net.half()`, SURVEY.md
App. C #6) and the per-step CFG combine + x_{t-1} update fused into one CUDA kernel that reproduces
the reference's fp16 rounding sequence (pfd_ddim_step_f16).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import native as nv
from .graphs import weights_signature


def eta_is_zero(sigmas) -> bool:
    return not np.any(np.asarray(sigmas, dtype=np.float64) != 0.0)


def make_ddim_timesteps(num_ddim_timesteps, num_ddpm_timesteps):
    """diffusion_utils.py:32-46, 'uniform': stride T//S then +1 (steps=30 yields 31 evaluations)."""
    c = num_ddpm_timesteps // num_ddim_timesteps
    return np.asarray(list(range(0, num_ddpm_timesteps, c))) + 1


class DDIMSampler(object):
    def __init__(self, model, schedule="linear", **kwargs):
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        self.use_cuda_graph = kwargs.get("use_cuda_graph", True)
        self._states = {}

    def register_buffer(self, name, attr):
        setattr(self, name, attr)

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0.0, verbose=True):
        """ddim.py:23-56 — same host ops in the same order."""
        if ddim_discretize != "uniform":
            raise NotImplementedError("only the 'uniform' DDIM discretisation is used by app.py")
        self.ddim_timesteps = make_ddim_timesteps(ddim_num_steps, self.ddpm_num_timesteps)
        ac = self.model.alphas_cumprod
        assert ac.shape[0] == self.ddpm_num_timesteps
        to32 = lambda x: x.clone().detach().to(torch.float32).cpu()
        self.betas = to32(self.model.betas)
        self.alphas_cumprod = to32(ac)
        self.alphas_cumprod_prev = to32(self.model.alphas_cumprod_prev)
        acc = self.alphas_cumprod
        ts = self.ddim_timesteps
        alphas = acc[ts]                                                 # fp32 torch tensor
        alphas_prev = np.asarray([acc[0]] + acc[ts[:-1]].tolist())       # float64 numpy (ddim.py via diffusion_utils:51)
        sigmas = ddim_eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
        self.ddim_sigmas, self.ddim_alphas, self.ddim_alphas_prev = sigmas, alphas, alphas_prev
        self.ddim_sqrt_one_minus_alphas = np.sqrt(1.0 - alphas)

    def _coef_table(self, device) -> torch.Tensor:
        """[steps, 4] fp32 table of the per-step coefficients already rounded to fp16 the way
        torch.full(..., dtype=float16) rounds them (ddim.py:160-163)."""
        n = self.ddim_timesteps.shape[0]
        f16 = lambda v: float(torch.as_tensor(v).to(torch.float64).to(torch.float16)) if not torch.is_tensor(v) \
            else float(v.to(torch.float16))
        rows = [[f16(self.ddim_alphas[i]), f16(self.ddim_alphas_prev[i]), f16(self.ddim_sigmas[i]),
                 f16(self.ddim_sqrt_one_minus_alphas[i])] for i in range(n)]
        return torch.tensor(rows, dtype=torch.float32, device=device)

    @torch.no_grad()
    def sample(self, steps, shape, x_info, c_info, eta=0.0, temperature=1.0, noise_dropout=0.0, verbose=True,
               log_every_t=100):
        self.make_schedule(ddim_num_steps=steps, ddim_eta=eta, verbose=verbose)
        return self.ddim_sampling(shape, x_info=x_info, c_info=c_info, noise_dropout=noise_dropout,
                                  temperature=temperature, log_every_t=log_every_t)

    @torch.no_grad()
    def ddim_sampling(self, shape, x_info, c_info, noise_dropout=0.0, temperature=1.0, log_every_t=100):
        """ddim.py:81-127.  The per-step work (CFG batch -> UNet [+ControlNet] -> fused CFG combine + DDIM
        update, in place on a static latent buffer) and the step-invariant preparation (cross-attention
        K/V of the context, ControlNet hint stem) are captured into CUDA graphs that are cached across
        calls, keyed on shapes + a signature of the weights they baked in."""
        model = self.model
        device = model.device
        if noise_dropout > 0.0:
            raise NotImplementedError("noise_dropout is a training-time option not used by app.py")
        bs = shape[0]
        timesteps = self.ddim_timesteps
        if x_info.get("xt", None) is not None:
            x_T = x_info["xt"].to(device=device, dtype=torch.float16)
        elif x_info.get("x0", None) is not None:
            # img2img branch (ddim.py:94-101): noise x0 forward to the n-th DDIM timestep (q_sample, pfd.py:204-207,
            # same torch.randn_like RNG call) and run only the first n timesteps of the schedule
            n_fwd = int(x_info["x0_forward_timesteps"])
            x0 = x_info["x0"].to(device=device, dtype=torch.float16).contiguous()
            t_fwd = int(timesteps[n_fwd])
            timesteps = timesteps[:n_fwd]
            noise = torch.randn_like(x0)
            x_T = nv.axpby(x0, float(model.sqrt_alphas_cumprod[t_fwd]), noise,
                           float(model.sqrt_one_minus_alphas_cumprod[t_fwd]))
        else:
            # same RNG call as ddim.py:105 (dtype of the conditioning; fp16 on the GPU path)
            x_T = torch.randn(shape, device=device, dtype=c_info["conditioning"].dtype).to(torch.float16)
        guidance = float(c_info["unconditional_guidance_scale"])
        cond = c_info["conditioning"]
        uncond = c_info.get("unconditional_conditioning", None)
        use_cfg = not (guidance == 1.0 or uncond is None)
        c_full = (torch.cat([uncond, cond]) if use_cfg else cond).to(torch.float16).contiguous()   # ddim.py:147
        cc = c_info.get("control", None)
        total = timesteps.shape[0]
        nb = 2 * bs if use_cfg else bs
        eta0 = eta_is_zero(self.ddim_sigmas)

        key = (tuple(x_T.shape), tuple(c_full.shape), use_cfg, guidance, c_info["type"], x_info["type"],
               None if cc is None else (tuple(cc.shape), cc.dtype), total, weights_signature(model))
        st = self._states.get(key) if self.use_cuda_graph else None
        if st is None:
            st = _SamplerState(model, x_T, c_full, cc, nb, total, use_cfg, guidance, x_info["type"], c_info["type"],
                               capture=self.use_cuda_graph and eta0 and total > 1)
            if self.use_cuda_graph:
                if len(self._states) >= 2:
                    self._states.pop(next(iter(self._states)))
                self._states[key] = st
        st.load_request(x_T, c_full, cc, self._coef_table(device)[:total])
        x = st.x
        intermediates = {"pred_xt": [], "pred_x0": []}
        for i, step in enumerate(np.flip(timesteps)):
            index = total - i - 1
            st.t_in.fill_(int(step))
            st.step_idx.fill_(index)
            st.step()
            sigma = float(self.ddim_sigmas[index])
            if sigma != 0.0:
                noise = torch.randn_like(x)
                nv.axpby(x, 1.0, noise, sigma * temperature, out=x)
            if index % log_every_t == 0 or index == total - 1:
                intermediates["pred_xt"].append(x.clone())
                intermediates["pred_x0"].append(st.pred_x0.clone())
        out = x.clone()
        x_info["x"] = out
        c_info["c"] = c_full
        return out, intermediates


class _SamplerState:
    """Static buffers + captured graphs of one sampling configuration."""

    def __init__(self, model, x_T, c_full, cc, nb, total, use_cfg, guidance, x_type, c_type, capture):
        dev = x_T.device
        self.model, self.use_cfg, self.guidance = model, use_cfg, guidance
        self.x = torch.empty_like(x_T)
        self.pred_x0 = torch.empty_like(x_T)
        self.c = torch.empty_like(c_full)
        self.cc = None if cc is None else torch.empty_like(cc)
        self.t_in = torch.zeros((nb,), device=dev, dtype=torch.long)
        self.step_idx = torch.zeros(1, dtype=torch.int32, device=dev)
        self.coef = torch.zeros((total, 4), dtype=torch.float32, device=dev)
        self.x_info = {"type": x_type}
        self.c_info = {"type": c_type, "control": self.cc}
        self.prep_graph = self.step_graph = None
        self.n_prep = self.n_step = 0
        # eager pass first: builds every packed-weight cache and validates the launch sequence
        self.x.copy_(x_T)
        self.c.copy_(c_full)
        if cc is not None:
            self.cc.copy_(cc)
        self._prepare()
        if capture:
            self.t_in.fill_(1)
            self._one_step()                       # warm-up on scratch state (x is re-loaded per request)
            torch.cuda.synchronize()
            self.prep_graph = torch.cuda.CUDAGraph()
            n0 = nv.launch_count()
            with torch.cuda.graph(self.prep_graph):
                self._prepare()
            self.n_prep = nv.launch_count() - n0
            self.step_graph = torch.cuda.CUDAGraph()
            n0 = nv.launch_count()
            with torch.cuda.graph(self.step_graph):
                self._one_step()
            self.n_step = nv.launch_count() - n0

    def _prepare(self):
        prep = self.model.prepare_context(self.c, self.c_info["type"])
        if self.cc is not None and hasattr(self.model, "ctl"):
            prep["hint"] = self.model.ctl.hint_features(self.cc)
        self.c_info["c"] = prep["c"]
        self.c_info["_pfd_prepared"] = prep

    def _one_step(self):
        # CFG batch (ddim.py:145-150) -> UNet (+ControlNet) -> fused CFG combine + DDIM update, in place on x
        x = self.x
        self.x_info["x"] = torch.cat([x, x]) if self.use_cfg else x
        eps = self.model.apply_model(self.x_info, self.t_in, self.c_info)
        if not self.use_cfg:                                             # e_t = eps * scale (ddim.py:143-144)
            eps = torch.cat([torch.zeros_like(eps), eps])
        nv.ddim_step(eps, x, self.guidance, self.coef, self.step_idx, x, self.pred_x0)

    def load_request(self, x_T, c_full, cc, coef):
        self.x.copy_(x_T)
        self.c.copy_(c_full)
        if cc is not None:
            self.cc.copy_(cc)
        self.coef.copy_(coef)
        if self.prep_graph is not None:
            self.prep_graph.replay()
            nv.note_replay(self.n_prep)
        else:
            self._prepare()

    def step(self):
        if self.step_graph is not None:
            self.step_graph.replay()
            nv.note_replay(self.n_step)
        else:
            self._one_step()
