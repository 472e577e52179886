"""Deterministic synthetic weights for the Prompt-Free-Diffusion pipeline.

No pretrained checkpoints exist offline, so parity and throughput are measured on seeded random
weights of the reference architecture.  Each tensor is generated from a seed derived
from its state-dict *name*, so the reference modules, the CPU oracle and the CUDA path can all be
filled identically, independent of construction order.  Zero-initialised tensors of the reference
(zero_module, diffusion_utils.py:153-159: ResBlock out conv, proj_out, ControlNet zero convs, ...)
get non-zero values too — otherwise the UNet output would be identically zero and parity vacuous.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Iterable, Mapping, Sequence, Tuple

import torch

# last layer of a residual branch: damped so activations stay O(1) through ~60 residual adds
_BRANCH_OUT = ("out_layers.3.weight", "proj_out.weight", "to_out.0.weight", "ff.net.2.weight",
               "conv2.weight", "mlp.fc2.weight", "attn.proj.weight", "linear2.weight",
               "out_proj.weight", "zero_convs", "middle_block_out", "input_hint_block.14.weight")
_UNIT_STD = ("init_query.weight", "query_pos_embedding.weight", "level_embed")


def synth_tensor(name: str, shape: Sequence[int], seed: int = 0, dtype=torch.float32) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    shape = tuple(shape)
    x = torch.randn(shape, generator=g, dtype=torch.float32)
    if name.endswith("relative_position_bias_table"):
        x = x * 0.5
    elif any(s in name for s in _UNIT_STD):
        pass
    elif len(shape) >= 2:
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        gain = 0.35 if any(s in name for s in _BRANCH_OUT) else 1.0
        x = x * (gain / math.sqrt(fan_in))
    elif name.endswith("weight"):            # 1-D weights are always norm scales on this path
        x = 1.0 + 0.1 * x
    else:                                    # biases
        x = 0.05 * x
    return x.to(dtype)


def synth_state_dict(shapes: Mapping[str, Tuple[Sequence[int], str]], seed: int = 0,
                     prefix: str = "", dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """`shapes`: name -> (shape, dtype-name).  Integer buffers (relative_position_index) and the
    schedule buffers are not random and are skipped (callers keep their own values)."""
    out = {}
    for name, (shape, dt) in shapes.items():
        if not name.startswith(prefix):
            continue
        if dt not in ("float32", "float16", "bfloat16"):
            continue
        out[name] = synth_tensor(name, shape, seed, dtype)
    return out


def fill_module_(module: torch.nn.Module, seed: int = 0, prefix: str = "", skip: Iterable[str] = ()) -> None:
    """Overwrite every floating-point entry of module.state_dict() with its synthetic value (in place).
    `prefix` is prepended to the local names so that sub-modules get the same values they would get
    inside the full pipeline (e.g. prefix='diffuser.image.')."""
    skip = tuple(skip)
    with torch.no_grad():
        for name, t in module.state_dict().items():
            if not t.dtype.is_floating_point or any(name.startswith(s) for s in skip):
                continue
            t.copy_(synth_tensor(prefix + name, t.shape, seed).to(t.dtype))


SCHEDULE_BUFFERS = ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
                    "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod",
                    "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
                    "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2")
