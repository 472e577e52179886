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
        self._graph = None

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
        """ddim.py:81-127."""
        model = self.model
        device = model.device
        if noise_dropout > 0.0:
            raise NotImplementedError("noise_dropout is a training-time option not used by app.py")
        bs = shape[0]
        timesteps = self.ddim_timesteps
        if x_info.get("xt", None) is not None:
            x = x_info["xt"].to(device=device, dtype=torch.float16).clone()
        elif x_info.get("x0", None) is not None:
            raise NotImplementedError("img2img (x0) sampling is outside the pfd_b200 hot path (SURVEY.md §8f)")
        else:
            # same RNG call as ddim.py:105 (dtype of the conditioning; fp16 on the GPU path)
            x = torch.randn(shape, device=device, dtype=c_info["conditioning"].dtype).to(torch.float16)
        x_info["x"] = x
        guidance = float(c_info["unconditional_guidance_scale"])
        cond = c_info["conditioning"]
        uncond = c_info.get("unconditional_conditioning", None)
        use_cfg = not (guidance == 1.0 or uncond is None)
        c_full = torch.cat([uncond, cond]) if use_cfg else cond          # ddim.py:147
        prep = model.prepare_context(c_full, c_info["type"])
        c_info["c"] = prep["c"]
        c_info["_pfd_prepared"] = prep
        coef = self._coef_table(device)
        step_idx = torch.zeros(1, dtype=torch.int32, device=device)
        total = timesteps.shape[0]
        intermediates = {"pred_xt": [], "pred_x0": []}
        pred_x0 = torch.empty_like(x)
        nb = 2 * bs if use_cfg else bs
        t_in = torch.zeros((nb,), device=device, dtype=torch.long)
        x = x.contiguous()

        def one_step():
            # CFG batch (ddim.py:145-150) -> UNet (+ControlNet) -> fused CFG combine + DDIM update, in place on x
            x_info["x"] = torch.cat([x, x]) if use_cfg else x
            eps = model.apply_model(x_info, t_in, c_info)
            if not use_cfg:                                              # e_t = eps * scale (ddim.py:143-144)
                eps = torch.cat([torch.zeros_like(eps), eps])
            nv.ddim_step(eps, x, guidance, coef, step_idx, x, pred_x0)

        graph = None
        use_graph = self.use_cuda_graph and eta_is_zero(self.ddim_sigmas) and total > 2
        for i, step in enumerate(np.flip(timesteps)):
            index = total - i - 1
            t_in.fill_(int(step))
            step_idx.fill_(index)
            if use_graph and i == 1:
                # step 0 ran eagerly (it also built every packed-weight / hint cache); capture the
                # identical launch sequence once and replay it for the remaining steps.
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                n_before = nv.launch_count()
                with torch.cuda.graph(graph):
                    one_step()
                n_nodes = nv.launch_count() - n_before
            if graph is not None:
                graph.replay()
                nv.note_replay(n_nodes)
            else:
                one_step()
            sigma = float(self.ddim_sigmas[index])
            if sigma != 0.0:
                noise = torch.randn_like(x)
                nv.axpby(x, 1.0, noise, sigma * temperature, out=x)
            x_info["x"] = x
            if index % log_every_t == 0 or index == total - 1:
                intermediates["pred_xt"].append(x.clone())
                intermediates["pred_x0"].append(pred_x0.clone())
        if graph is not None:
            torch.cuda.synchronize()
            del graph
        c_info.pop("_pfd_prepared", None)
        return x, intermediates
