"""Pin the oracle against the UNMODIFIED reference and write the golden fixtures.

Runs only in the build container (needs /root/reference).  For every stage of the hot path it
 1. builds the reference pipeline (lib/model_zoo via tools/ref_harness.py) and fills it with the
    name-seeded synthetic weights of pfd_b200/weights.py,
 2. runs the reference and the oracle (oracle/pfd_oracle.py) on identical seeded inputs in fp32 on
    the CPU and asserts they agree (max |diff| printed per stage),
 3. stores the REFERENCE outputs under tests/golden/ (small .npz files) together with the
    state-dict key/shape table the product's module tree must reproduce.

    python tools/make_golden.py            # ~6 min on 8 cores
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
GOLD = os.path.join(ROOT, "tests", "golden")

import ref_harness as rh  # noqa: E402
from oracle import pfd_oracle as O  # noqa: E402
from pfd_b200.weights import SCHEDULE_BUFFERS, fill_module_  # noqa: E402


from oracle.golden_inputs import golden_inputs  # noqa: E402


def sub_sample(t, stride=37):
    f = t.detach().float().reshape(-1)
    return f[::stride].numpy().astype(np.float32)


def report(name, ref, ora):
    d = (ref.float() - ora.float()).abs().max().item()
    s = ref.float().abs().max().item()
    print(f"  {name:28s} max|ref-oracle| = {d:.3e}   (ref max {s:.3e}, rms {ref.float().pow(2).mean().sqrt().item():.3e})")
    assert d <= 2e-4 * max(1.0, s), f"oracle deviates from the reference at stage {name}"
    return d


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_grad_enabled(False)
    t0 = time.time()
    net, cfgm = rh.build_reference_net()
    print(f"reference net built in {time.time() - t0:.0f}s")
    shapes = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in net.state_dict().items()}
    json.dump(shapes, open(os.path.join(GOLD, "state_dict_shapes.json"), "w"), indent=0, sort_keys=True)
    fill_module_(net, seed=0, skip=SCHEDULE_BUFFERS)
    net.device = "cpu"
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    inp = golden_inputs()
    out = {}
    devs = {}

    # ---- schedule buffers (closed form)
    buf = O.schedule_buffers()
    for k in SCHEDULE_BUFFERS:
        assert torch.equal(buf[k], sd[k]), k
    print("  schedule buffers: bit-identical")

    # ---- UNet (no control)
    usd = O.sub(sd, "diffuser.image.")
    ref = net.apply_model({"type": "image", "x": inp["x"]}, inp["t"], {"type": "image", "c": inp["ctx"], "control": None})
    ora = O.unet_apply(usd, O.UNET_SD15, inp["x"], inp["t"], inp["ctx"])
    devs["unet"] = report("unet eps", ref, ora)
    out["unet_eps"] = ref.numpy()

    # ---- ControlNet + UNet with control
    csd = O.sub(sd, "ctl.")
    ref_c = net.ctl(inp["x"], hint=inp["hint"], timesteps=inp["t"], context=inp["ctx"])
    ora_c = O.controlnet_apply(csd, O.CONTROLNET_SD15, inp["x"], inp["hint"], inp["t"], inp["ctx"])
    for i, (a, b) in enumerate(zip(ref_c, ora_c)):
        devs[f"ctl{i}"] = report(f"controlnet out[{i}] {tuple(a.shape)}", a, b)
        out[f"ctl_{i}_sub"] = sub_sample(a)
        out[f"ctl_{i}_stats"] = np.array([a.mean().item(), a.std().item(), a.abs().max().item()], dtype=np.float32)
    ref = net.apply_model({"type": "image", "x": inp["x"]}, inp["t"], {"type": "image", "c": inp["ctx"], "control": inp["hint"]})
    ora = O.unet_apply(usd, O.UNET_SD15, inp["x"], inp["t"], inp["ctx"], control=ora_c)
    devs["unet_ctl"] = report("unet eps (with control)", ref, ora)
    out["unet_eps_control"] = ref.numpy()

    # ---- VAE decode
    vsd = O.sub(sd, "vae.image.")
    ref = net.vae_decode(inp["z"], "image")
    ora = O.vae_decode(vsd, O.VAE_SD, inp["z"])
    devs["vae"] = report("vae decode", ref, ora)
    out["vae_image"] = ref.numpy()

    # ---- SeeCoder
    ssd = O.sub(sd, "ctx.image.")
    fea_ref = net.ctx["image"].imencoder(inp["img"])
    fea_ora = O.swin_forward(O.sub(ssd, "imencoder."), O.SWIN_L, inp["img"])
    for k in ("res3", "res4", "res5"):
        devs["swin_" + k] = report(f"swin {k}", fea_ref[k], fea_ora[k])
        out[f"swin_{k}_sub"] = sub_sample(fea_ref[k], 11)
    ref = net.ctx_encode(inp["img"], "image")
    ora = O.seecoder_encode(ssd, inp["img"])
    devs["seecoder"] = report("seecoder context", ref, ora)
    out["seecoder_ctx"] = ref.numpy().astype(np.float16)

    # ---- position-aware variant (app.py:166-175 installs PPE_MLP at run time)
    from lib.model_zoo.seecoder import PPE_MLP
    pe = PPE_MLP(freq_num=20, freq_max=None, out_channel=768, mlp_layer=3)
    fill_module_(pe, seed=0, prefix="ctx.image.qtransformer.pe_layer.")
    pe.eval()                                                            # app.py:174
    net.ctx["image"].qtransformer.pe_layer = pe
    ref = net.ctx_encode(inp["img"], "image")
    ssd_pa = dict(ssd)
    ssd_pa.update({"qtransformer.pe_layer." + k: v for k, v in pe.state_dict().items()})
    ora = O.seecoder_encode(ssd_pa, inp["img"])
    devs["seecoder_pa"] = report("seecoder context (PA)", ref, ora)
    out["seecoder_ctx_pa"] = ref.numpy().astype(np.float16)
    net.ctx["image"].qtransformer.pe_layer = None

    # ---- DDIM sampler, 4 steps, CFG 2.0, fixed x_T (torch.randn patched for the reference call)
    sampler = rh.cpu_sampler(net)
    real_randn = torch.randn
    torch.randn = lambda *a, **k: inp["x_T"].clone()
    try:
        ref, _ = sampler.sample(steps=4, x_info={"type": "image"},
                                c_info={"type": "image", "conditioning": inp["cond"],
                                        "unconditional_conditioning": torch.zeros_like(inp["cond"]),
                                        "unconditional_guidance_scale": 2.0, "control": None},
                                shape=[1, 4, 16, 16], verbose=False, eta=0.0)
    finally:
        torch.randn = real_randn
    ora = O.ddim_sample(usd, O.UNET_SD15, sd["alphas_cumprod"], steps=4, x_T=inp["x_T"], cond=inp["cond"],
                        uncond=torch.zeros_like(inp["cond"]), guidance=2.0)
    devs["ddim4"] = report("ddim 4-step latent", ref, ora)
    out["ddim4_latent"] = ref.numpy()
    from lib.model_zoo.diffusion_utils import make_ddim_timesteps
    for n in (10, 30, 50):
        assert list(O.ddim_timesteps(n)) == list(make_ddim_timesteps("uniform", n, 1000, verbose=False))
    assert len(O.ddim_timesteps(30)) == 31

    np.savez_compressed(os.path.join(GOLD, "reference_outputs.npz"), **out)
    json.dump({"max_abs_dev_reference_vs_oracle": devs, "torch": torch.__version__,
               "generated_by": "tools/make_golden.py", "weights": "pfd_b200.weights.synth_tensor(seed=0)"},
              open(os.path.join(GOLD, "oracle_pin_report.json"), "w"), indent=1)
    print(f"done in {time.time() - t0:.0f}s; golden files in {GOLD}")


if __name__ == "__main__":
    main()
