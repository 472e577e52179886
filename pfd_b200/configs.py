"""Model configurations of the Prompt-Free-Diffusion pipeline as plain dicts.

Same names, types and argument values as the reference's yaml bank (configs/model/pfd.yaml,
openai_unet.yaml, seecoder.yaml, swin.yaml, autokl.yaml, controlnet.yaml), with `MODEL(x)` references
and `super_cfg` inheritance already resolved, so `model_cfg_bank()('pfd_seecoder_with_controlnet')`
returns the same structure lib/cfg_helper.py:107-146 produces (minus the absent VAE checkpoint path).
"""
from __future__ import annotations

import copy

from .registry import AttrDict

_SWIN_LARGE = dict(type="swin", strict_sd=False, args=dict(
    embed_dim=192, depths=[2, 2, 18, 2], num_heads=[6, 12, 24, 48], window_size=12, ape=False,
    drop_path_rate=0.3, patch_norm=True))
_SEECODER_DECODER = dict(type="seecoder_decoder", args=dict(
    inchannels=dict(res3=384, res4=768, res5=1536), trans_input_tags=["res3", "res4", "res5"], trans_dim=768,
    trans_dropout=0.1, trans_nheads=8, trans_feedforward_dim=1024, trans_num_layers=6))
_QT = dict(type="seecoder_query_transformer", args=dict(
    in_channels=768, hidden_dim=768, num_queries=[4, 144], nheads=8, num_layers=9, feedforward_dim=2048,
    pre_norm=False, num_feature_levels=3, enforce_input_project=False, with_fea2d_pos=False))
_QT_PA = copy.deepcopy(_QT)
_QT_PA["args"]["with_fea2d_pos"] = True
_SEECODER = dict(type="seecoder", args=dict(imencoder_cfg=_SWIN_LARGE, imdecoder_cfg=_SEECODER_DECODER,
                                            qtransformer_cfg=_QT))
_SEECODER_PA = dict(type="seecoder", args=dict(imencoder_cfg=_SWIN_LARGE, imdecoder_cfg=_SEECODER_DECODER,
                                               qtransformer_cfg=_QT_PA))
_UNET = dict(type="openai_unet_2d_next", args=dict(
    in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
    num_res_blocks=[2, 2, 2, 2], channel_mult=[1, 2, 4, 4], num_heads=8, context_dim=768,
    use_checkpoint=False, parts=["global", "data", "context"]))
_AUTOKL = dict(type="autoencoderkl", args=dict(embed_dim=4, lossconfig=None, ddconfig=dict(
    double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
    num_res_blocks=2, attn_resolutions=[], dropout=0.0)))
_CONTROLNET = dict(type="controlnet", args=dict(
    image_size=32, in_channels=4, hint_channels=3, model_channels=320, attention_resolutions=[4, 2, 1],
    num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True,
    transformer_depth=1, context_dim=768, use_checkpoint=True, legacy=False))
_PFD_ARGS = dict(beta_linear_start=0.00085, beta_linear_end=0.012, timesteps=1000, use_ema=False,
                 vae_cfg_list=[["image", _AUTOKL]], ctx_cfg_list=[["image", _SEECODER]],
                 diffuser_cfg_list=[["image", _UNET]], latent_scale_factor=dict(image=0.18215))

MODEL_BANK = {
    "swin_large": _SWIN_LARGE,
    "seecoder_decoder": _SEECODER_DECODER,
    "seecoder_query_transformer": _QT,
    "seecoder_query_transformer_position_aware": _QT_PA,
    "seecoder": _SEECODER,
    "seecoder_pa": _SEECODER_PA,
    "openai_unet_2d_v1": _UNET,
    "autokl_v2": _AUTOKL,
    "controlnet": _CONTROLNET,
    "pfd_seecoder": dict(type="pfd", args=_PFD_ARGS),
    "pfd_seecoder_with_controlnet": dict(type="pfd_with_control", args=dict(_PFD_ARGS, ctl_cfg=_CONTROLNET)),
}


class model_cfg_bank(object):
    """lib/cfg_helper.py:102-146 surface: model_cfg_bank()(name) -> attr-dict {type, args, name}."""

    def __call__(self, name):
        if name not in MODEL_BANK:
            raise ValueError(f"No model named {name}")
        cfg = AttrDict(copy.deepcopy(MODEL_BANK[name]))
        cfg["name"] = name
        return cfg
