"""Registers the pfd_b200 networks under the reference's type names (get_model.py:72-85)."""
from .registry import get_model
from .autokl import AutoencoderKL
from .controlnet import ControlNet
from .pfd import PromptFreeDiffusion, PromptFreeDiffusion_with_control
from .seecoder import Decoder, QueryTransformer, SemanticContextEncoder
from .swin import SwinTransformer
from .unet import UNetModel2D_Next

for _name, _cls in {
    "pfd": PromptFreeDiffusion,
    "pfd_with_control": PromptFreeDiffusion_with_control,
    "autoencoderkl": AutoencoderKL,
    "openai_unet_2d_next": UNetModel2D_Next,
    "controlnet": ControlNet,
    "seecoder": SemanticContextEncoder,
    "seecoder_decoder": Decoder,
    "seecoder_query_transformer": QueryTransformer,
    "swin": SwinTransformer,
}.items():
    get_model().register(_cls, _name)
