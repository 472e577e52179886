"""GPU parity of the full pfd_b200 pipeline (through its public, reference-shaped API and therefore
through the C-ABI) against (a) golden outputs produced by the UNMODIFIED reference in fp32
(tests/golden/reference_outputs.npz, tools/make_golden.py) and (b) the CPU oracle run on this box.

Tolerance: the path computes in fp16 with fp32 accumulation; BASELINE.json's bar is latent
MSE < 1e-3.  We assert both absolute MSE < 1e-3 and relative rms error < 2e-2 per stage (the
reference's own fp16-vs-fp32 disagreement is ~1e-3 relative rms per UNet call).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _metrics(out, ref):
    out, ref = out.detach().float().cpu(), torch.as_tensor(ref).float()
    mse = (out - ref).pow(2).mean().item()
    rel = (mse / max(ref.pow(2).mean().item(), 1e-20)) ** 0.5
    return mse, rel


def _check(name, out, ref, mse_tol=1e-3, rel_tol=2e-2):
    mse, rel = _metrics(out, ref)
    print(f"[parity] {name}: mse={mse:.3e} rel_rms={rel:.3e}")
    assert np.isfinite(mse) and mse < mse_tol and rel < rel_tol, f"{name}: mse={mse:.3e} rel_rms={rel:.3e}"


@pytest.fixture(scope="module")
def env():
    from oracle.golden_inputs import golden_inputs
    from pfd_b200 import get_model, model_cfg_bank
    from pfd_b200.weights import SCHEDULE_BUFFERS, fill_module_
    net = get_model()(model_cfg_bank()("pfd_seecoder_with_controlnet"))
    fill_module_(net, seed=0, skip=SCHEDULE_BUFFERS)
    net = net.half()
    net.to("cuda")
    net.eval()
    gold = dict(np.load(os.path.join(GOLD, "reference_outputs.npz")))
    inp = {k: v.cuda() for k, v in golden_inputs().items()}
    return net, gold, inp


def test_unet_eps_matches_reference(env):
    net, gold, inp = env
    eps = net.apply_model({"type": "image", "x": inp["x"].half()}, inp["t"],
                          {"type": "image", "c": inp["ctx"].half(), "control": None})
    _check("unet eps", eps, gold["unet_eps"])


def test_controlnet_and_controlled_unet_match_reference(env):
    net, gold, inp = env
    outs = net.ctl(inp["x"].half(), hint=inp["hint"].half(), timesteps=inp["t"], context=inp["ctx"].half())
    assert len(outs) == 13
    for i, o in enumerate(outs):
        nchw = o.permute(0, 3, 1, 2).float().cpu()
        sub = nchw.reshape(-1)[::37]
        _check(f"controlnet out[{i}]", sub, gold[f"ctl_{i}_sub"])
    eps = net.apply_model({"type": "image", "x": inp["x"].half()}, inp["t"],
                          {"type": "image", "c": inp["ctx"].half(), "control": inp["hint"].half()})
    _check("unet eps (control)", eps, gold["unet_eps_control"])


def test_vae_decode_matches_reference(env):
    net, gold, inp = env
    im = net.vae_decode(inp["z"].half(), "image")
    assert im.shape == (1, 3, 64, 64) and im.min() >= 0 and im.max() <= 1
    _check("vae image", im, gold["vae_image"])


def test_seecoder_matches_reference(env):
    net, gold, inp = env
    fea = net.ctx["image"].imencoder(inp["img"])
    for k in ("res3", "res4", "res5"):
        nchw = fea[k].permute(0, 3, 1, 2).contiguous().float().cpu()
        _check(f"swin {k}", nchw.reshape(-1)[::11], gold[f"swin_{k}_sub"])
    c = net.ctx_encode(inp["img"], "image")
    assert c.shape == (1, 148, 768)
    _check("seecoder ctx", c, gold["seecoder_ctx"].astype(np.float32))


def test_seecoder_position_aware_matches_reference(env):
    net, gold, inp = env
    from pfd_b200.seecoder import PPE_MLP
    from pfd_b200.weights import fill_module_
    pe = PPE_MLP(freq_num=20, freq_max=None, out_channel=768, mlp_layer=3)
    fill_module_(pe, seed=0, prefix="ctx.image.qtransformer.pe_layer.")
    qt = net.ctx["image"].qtransformer
    qt.pe_layer = pe.half().cuda()                                       # app.py:166-175 hot-swap
    try:
        c = net.ctx_encode(inp["img"], "image")
    finally:
        qt.pe_layer = None
    _check("seecoder ctx (PA)", c, gold["seecoder_ctx_pa"].astype(np.float32))


def test_ddim_sampler_matches_reference(env):
    net, gold, inp = env
    from pfd_b200 import DDIMSampler
    sampler = DDIMSampler(net)
    cond = inp["cond"].half()
    x, inter = sampler.sample(steps=4, x_info={"type": "image", "xt": inp["x_T"].half()},
                              c_info={"type": "image", "conditioning": cond,
                                      "unconditional_conditioning": torch.zeros_like(cond),
                                      "unconditional_guidance_scale": 2.0, "control": None},
                              shape=[1, 4, 16, 16], verbose=False, eta=0.0)
    assert x.shape == (1, 4, 16, 16) and len(inter["pred_x0"]) >= 1
    _check("ddim 4-step latent", x, gold["ddim4_latent"])


def test_unet_matches_oracle_on_this_box():
    """Fresh seeded inputs at a size not in the goldens (24x24 latents, batch 3) vs the CPU oracle."""
    from oracle import pfd_oracle as O
    from pfd_b200 import get_model, model_cfg_bank
    from pfd_b200.weights import fill_module_
    unet = get_model()(model_cfg_bank()("openai_unet_2d_v1"))
    fill_module_(unet, seed=3, prefix="diffuser.image.")
    g = torch.Generator().manual_seed(77)
    x, ctx = torch.randn((3, 4, 24, 24), generator=g), 0.5 * torch.randn((3, 148, 768), generator=g)
    t = torch.tensor([981, 981, 981])
    sd = {k: v.detach() for k, v in unet.state_dict().items()}
    with torch.no_grad():
        ref = O.unet_apply(sd, O.UNET_SD15, x, t, ctx)
    unet = unet.half().cuda()
    out = unet.apply(x.cuda().half(), t.cuda(), ctx.cuda().half())
    _check("unet eps vs oracle (24x24, B=3)", out, ref)


def test_unet_768_matches_oracle():
    """BASELINE configs[4] geometry: 96x96 latents (N = 9216 / 2304 / 576 / 144 tokens, widths that are not
    powers of two) — one CFG pair vs the CPU oracle."""
    from oracle import pfd_oracle as O
    from pfd_b200 import get_model, model_cfg_bank
    from pfd_b200.weights import fill_module_
    unet = get_model()(model_cfg_bank()("openai_unet_2d_v1"))
    fill_module_(unet, seed=0, prefix="diffuser.image.")
    g = torch.Generator().manual_seed(78)
    x, ctx = torch.randn((1, 4, 96, 96), generator=g), 0.5 * torch.randn((1, 148, 768), generator=g)
    x_in, c_in, t = torch.cat([x, x]), torch.cat([torch.zeros_like(ctx), ctx]), torch.tensor([321, 321])
    sd = {k: v.detach() for k, v in unet.state_dict().items()}
    with torch.no_grad():
        ref = O.unet_apply(sd, O.UNET_SD15, x_in, t, c_in)
    unet = unet.half().cuda()
    out = unet.apply(x_in.cuda().half(), t.cuda(), c_in.cuda().half())
    _check("unet eps vs oracle (96x96 latents)", out, ref)


def test_vae_decode_512_matches_oracle():
    """Full-size decode of one 64x64 latent to 512x512 (GroupNorm over 64 MB tensors, N=4096 mid attention)."""
    from oracle import pfd_oracle as O
    from pfd_b200 import get_model, model_cfg_bank
    from pfd_b200.weights import fill_module_
    vae = get_model()(model_cfg_bank()("autokl_v2"))
    fill_module_(vae, seed=0, prefix="vae.image.")
    z = torch.randn((1, 4, 64, 64), generator=torch.Generator().manual_seed(79))
    sd = {k: v.detach() for k, v in vae.state_dict().items()}
    with torch.no_grad():
        ref = O.vae_decode(sd, O.VAE_SD, z)
    vae = vae.half().cuda()
    out = vae.decode(z.cuda().half(), pre_scale=1.0 / 0.18215)
    assert out.shape == (1, 3, 512, 512)
    _check("vae decode 512x512 vs oracle", out, ref)


def test_seecoder_512_and_sampler_graph_reuse(env):
    """512x512 reference image through SeeCoder (feature maps 128/64/32/16 -> padded windows), then two
    back-to-back sampler calls with different seeds: the second replays the cached CUDA graphs and must
    give a different (seed-dependent) but finite latent, and repeating seed 1 must reproduce run 1 (to rounding)."""
    net, gold, inp = env
    from pfd_b200 import DDIMSampler
    img = torch.rand((1, 3, 512, 512), generator=torch.Generator().manual_seed(80)).cuda()
    c = net.ctx_encode(img, "image")
    c2 = net.ctx_encode(img, "image")                                   # graph replay path
    # GroupNorm statistics are combined with atomics, so runs agree to rounding, not bit for bit
    assert c.shape == (1, 148, 768) and torch.isfinite(c.float()).all()
    assert (c.float() - c2.float()).abs().max().item() < 2e-2
    sampler = DDIMSampler(net)

    def run(seed):
        torch.manual_seed(seed)
        x, _ = sampler.sample(steps=4, x_info={"type": "image"},
                              c_info={"type": "image", "conditioning": c.repeat(2, 1, 1),
                                      "unconditional_conditioning": torch.zeros_like(c.repeat(2, 1, 1)),
                                      "unconditional_guidance_scale": 2.0, "control": None},
                              shape=[2, 4, 32, 32], verbose=False, eta=0.0)
        return x
    a, b, a2 = run(1), run(2), run(1)
    assert torch.isfinite(a.float()).all() and (a.float() - b.float()).abs().max().item() > 0.1
    assert (a.float() - a2.float()).abs().max().item() < 5e-2


def test_sampler_without_cfg_and_with_eta(env):
    """guidance == 1.0 skips the CFG batch (ddim.py:142-144) -> compare with the oracle sampler; eta > 0 adds
    sigma_t * noise every step (ddim.py:168) -> must stay finite and differ from the eta = 0 result."""
    net, gold, inp = env
    from oracle import pfd_oracle as O
    from pfd_b200 import DDIMSampler
    cond = inp["cond"].half()
    sampler = DDIMSampler(net)
    x, _ = sampler.sample(steps=4, x_info={"type": "image", "xt": inp["x_T"].half()},
                          c_info={"type": "image", "conditioning": cond, "unconditional_conditioning": None,
                                  "unconditional_guidance_scale": 1.0, "control": None},
                          shape=[1, 4, 16, 16], verbose=False, eta=0.0)
    sd = {k[len("diffuser.image."):]: v.detach().float().cpu() for k, v in net.state_dict().items()
          if k.startswith("diffuser.image.")}
    with torch.no_grad():
        ref = O.ddim_sample(sd, O.UNET_SD15, O.schedule_buffers()["alphas_cumprod"], steps=4, x_T=inp["x_T"].cpu(),
                            cond=inp["cond"].cpu().half().float(), uncond=None, guidance=1.0)
    _check("ddim 4-step latent, no CFG", x, ref, mse_tol=1e-3, rel_tol=2e-2)
    torch.manual_seed(3)
    xe, _ = sampler.sample(steps=4, x_info={"type": "image", "xt": inp["x_T"].half()},
                           c_info={"type": "image", "conditioning": cond, "unconditional_conditioning": None,
                                   "unconditional_guidance_scale": 1.0, "control": None},
                           shape=[1, 4, 16, 16], verbose=False, eta=0.5)
    assert torch.isfinite(xe.float()).all() and (xe.float() - x.float()).abs().max().item() > 1e-2


def test_sampler_x0_img2img_branch(env):
    """x_info['x0'] + 'x0_forward_timesteps' (ddim.py:94-101): x0 is noised forward with q_sample (same RNG call as
    the reference) and only the first n DDIM timesteps are walked -> compare with the oracle on the same noise."""
    net, gold, inp = env
    from oracle import pfd_oracle as O
    from pfd_b200 import DDIMSampler
    cond = inp["cond"].half()
    x0 = (inp["x_T"] * 0.5).half()
    n_fwd, steps = 3, 6
    torch.manual_seed(11)
    x, _ = DDIMSampler(net).sample(steps=steps, x_info={"type": "image", "x0": x0, "x0_forward_timesteps": n_fwd},
                                   c_info={"type": "image", "conditioning": cond, "unconditional_conditioning": None,
                                           "unconditional_guidance_scale": 1.0, "control": None},
                                   shape=[1, 4, 16, 16], verbose=False, eta=0.0)
    torch.manual_seed(11)
    noise = torch.randn_like(x0)
    bufs = O.schedule_buffers()
    ac16 = bufs["alphas_cumprod"].half().float()                      # the buffers are fp16 after net.half()
    t_fwd = int(O.ddim_timesteps(steps)[n_fwd])
    sa = float(net.sqrt_alphas_cumprod[t_fwd])
    sb = float(net.sqrt_one_minus_alphas_cumprod[t_fwd])
    x_T = sa * x0.float().cpu() + sb * noise.float().cpu()
    sd = {k[len("diffuser.image."):]: v.detach().float().cpu() for k, v in net.state_dict().items()
          if k.startswith("diffuser.image.")}
    with torch.no_grad():
        ref = O.ddim_sample(sd, O.UNET_SD15, bufs["alphas_cumprod"], steps=steps, x_T=x_T,
                            cond=inp["cond"].cpu().half().float(), uncond=None, guidance=1.0, n_forward=n_fwd)
    assert abs(float(ac16[t_fwd]) ** 0.5 - sa) < 2e-3
    _check("ddim img2img branch (x0 noised to step 3 of 6)", x, ref, mse_tol=1e-3, rel_tol=2e-2)


def test_native_library_is_what_ran():
    from pfd_b200 import native
    assert native.launch_count() > 0
