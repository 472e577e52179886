"""pfd_b200 — B200-native (sm_100a) implementation of the Prompt-Free-Diffusion inference hot path.

Public surface (mirrors the reference's lib.model_zoo / lib.cfg_helper plugin API):
    from pfd_b200 import get_model, register, model_cfg_bank, DDIMSampler
    net = get_model()(model_cfg_bank()('pfd_seecoder_with_controlnet')); net.to('cuda')
    c = net.ctx_encode(img, 'image'); x, _ = DDIMSampler(net).sample(...); im = net.vae_decode(x, 'image')
All arithmetic runs in the hand-written CUDA kernels behind include/pfd_b200.h (pfd_b200/native.py);
there is no torch / CPU fallback.
"""
from .registry import AttrDict, get_model, register, install_into_reference  # noqa: F401
from .configs import model_cfg_bank  # noqa: F401


def __getattr__(name):
    if name == "DDIMSampler":
        from .ddim import DDIMSampler
        return DDIMSampler
    raise AttributeError(name)
