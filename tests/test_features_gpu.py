"""GPU tests of the rows either side of the hot path and of the cache-invalidation /
graph-capture behaviour the boundary depends on."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).pow(2).mean() / b.pow(2).mean().clamp_min(1e-20)).sqrt().item()


@pytest.fixture(scope="module")
def net():
    from pfd_b200 import get_model, model_cfg_bank
    from pfd_b200.weights import SCHEDULE_BUFFERS, fill_module_
    n = get_model()(model_cfg_bank()("pfd_seecoder_with_controlnet"))
    fill_module_(n, seed=0, skip=SCHEDULE_BUFFERS)
    n = n.half()
    n.to("cuda")
    n.eval()
    return n


# ------------------------------------------------------------------------------------------------ f1: Canny
@pytest.mark.parametrize("shape,dtype", [((1, 3, 512, 512), torch.float16), ((2, 3, 131, 97), torch.float32),
                                         ((1, 3, 768, 640), torch.float32), ((1, 3, 33, 31), torch.float16)])
def test_canny_preprocess_bit_exact(net, shape, dtype):
    """ControlNet.preprocess(type='canny') on the GPU == cv2.Canny through the reference's tensor path
    (controlnet.py:332-360), bit for bit; images are smooth random fields with structure at several scales."""
    from oracle.canny_oracle import preprocess_canny
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.rand(shape, generator=g)
    x = 0.5 * torch.nn.functional.avg_pool2d(x, 9, 1, 4) + 0.5 * torch.nn.functional.avg_pool2d(x, 3, 1, 1)
    x = ((x - x.min()) / (x.max() - x.min())).to(dtype)
    for low, high in ((100, 200), (20, 50)):
        ref = preprocess_canny(x.cpu(), low, high)
        out = net.ctl.preprocess(x.cuda(), type="canny", low_threshold=low, high_threshold=high, size=list(shape[2:]))
        assert out.dtype == torch.float32 and out.shape == ref.shape and out.is_cuda
        nd = int((out.cpu() != ref).sum())
        assert nd == 0, f"{nd} of {ref.numel()} output values differ from cv2.Canny (edges: {int(ref.sum())})"
    assert ref.sum() > 0                                                    # the low thresholds do find edges
    try:
        import cv2
        from oracle.canny_oracle import to_pil_u8
        assert np.array_equal(cv2.Canny(to_pil_u8(x[0].cpu()), 20, 50), (out[0, 0].cpu().numpy() * 255).astype(np.uint8))
    except ImportError:
        pass


def test_canny_long_chain_needs_many_sweeps():
    """A one-pixel-wide weak spiral seeded by a single strong pixel: hysteresis must walk the whole chain across tiles."""
    from oracle.canny_oracle import preprocess_canny
    from pfd_b200 import native as nv
    H = W = 256
    img = torch.zeros((1, 3, H, W))
    y = x = 8
    dy, dx, seg = 0, 1, W - 16
    val = 0.15                                                              # |gradient| 4*38 = 152: between the thresholds
    while seg > 8:
        for _ in range(seg):
            img[0, :, y, x] = val
            y, x = y + dy, x + dx
        dy, dx = dx, -dy
        seg -= 6
    img[0, :, 8, 8:12] = 1.0                                                # strong seed at the start of the chain
    ref = preprocess_canny(img, 100, 200)
    out, sweeps = nv.canny(img.cuda(), 100, 200)
    assert torch.equal(out.cpu(), ref)
    assert sweeps >= 2
    print(f"[canny] spiral: {int(ref[0, 0].sum())} edge pixels, {sweeps} hysteresis sweeps")


def test_preprocess_input_and_none(net):
    x = torch.rand((1, 3, 40, 56), generator=torch.Generator().manual_seed(1))
    assert net.ctl.preprocess(x.cuda(), type="none") is None
    out = net.ctl.preprocess(x.cuda().half(), type="input")
    ref = x.half().mul(255).byte().float().div(255)
    assert out.dtype == torch.float32 and torch.equal(out.cpu(), ref)
    with pytest.raises(NotImplementedError):
        net.ctl.preprocess(x.cuda(), type="depth")


# ------------------------------------------------------------------------------------------------ f2: hot swap
def _sample(net, sampler, seed=7, steps=4, control=None):
    g = torch.Generator().manual_seed(seed)
    xT = torch.randn((1, 4, 16, 16), generator=g).cuda().half()
    cond = (0.5 * torch.randn((1, 148, 768), generator=g)).cuda().half()
    x, _ = sampler.sample(steps=steps, x_info={"type": "image", "xt": xT},
                          c_info={"type": "image", "conditioning": cond, "unconditional_conditioning": torch.zeros_like(cond),
                                  "unconditional_guidance_scale": 2.0, "control": control},
                          shape=[1, 4, 16, 16], verbose=False, eta=0.0)
    return x


def test_weight_hot_swap_invalidates_graphs_and_packs(net):
    """app.py:137-195: strict load_state_dict of another diffuser / ControlNet / VAE into the live net.  The cached
    CUDA graphs and packed weights must be rebuilt: outputs change, and equal the un-graphed path on the new weights."""
    from pfd_b200 import DDIMSampler
    from pfd_b200.weights import synth_tensor
    sampler = DDIMSampler(net)
    a = _sample(net, sampler)
    a2 = _sample(net, sampler)                                             # cached-graph replay
    # (GroupNorm statistics are combined with atomics: runs agree to fp16 rounding flips, not bit for bit)
    assert _rel(a2, a) < 5e-3
    old = {k: v.clone() for k, v in net.diffuser.state_dict().items()}
    new = {k: (synth_tensor("diffuser." + k, v.shape, seed=5).to(v.dtype) if v.dtype.is_floating_point else v)
           for k, v in old.items()}
    net.diffuser.load_state_dict(new, strict=True)                         # action_load_diffuser (app.py:153)
    b = _sample(net, sampler)
    assert _rel(b, a) > 0.05, "stale CUDA graph / packed weights were replayed after load_state_dict"
    b_eager = _sample(net, DDIMSampler(net, use_cuda_graph=False))
    assert _rel(b, b_eager) < 5e-3
    net.diffuser.load_state_dict(old, strict=True)
    c = _sample(net, sampler)
    assert _rel(c, a) < 5e-3
    # VAE decode graph
    z = torch.randn((1, 4, 16, 16), generator=torch.Generator().manual_seed(2)).cuda().half()
    im0 = net.vae_decode(z, "image")
    im0b = net.vae_decode(z, "image")
    assert _rel(im0b, im0) < 5e-3
    vold = {k: v.clone() for k, v in net.vae.state_dict().items()}
    net.vae.load_state_dict({k: (synth_tensor("vae." + k, v.shape, seed=9).to(v.dtype) if v.dtype.is_floating_point else v)
                             for k, v in vold.items()}, strict=True)
    im1 = net.vae_decode(z, "image")
    assert _rel(im1, im0) > 0.01
    net.vae.load_state_dict(vold, strict=True)
    assert _rel(net.vae_decode(z, "image"), im0) < 5e-3
    # ControlNet (net.ctl.load_state_dict, app.py:161)
    hint = (torch.rand((1, 3, 128, 128), generator=torch.Generator().manual_seed(3)) > 0.9).half().cuda()
    e0 = _sample(net, sampler, control=hint)
    cold = {k: v.clone() for k, v in net.ctl.state_dict().items()}
    net.ctl.load_state_dict({k: (synth_tensor("ctl." + k, v.shape, seed=11).to(v.dtype) if v.dtype.is_floating_point else v)
                             for k, v in cold.items()}, strict=True)
    e1 = _sample(net, sampler, control=hint)
    assert _rel(e1, e0) > 1e-3
    net.ctl.load_state_dict(cold, strict=True)
    assert _rel(_sample(net, sampler, control=hint), e0) < 5e-3


def test_explicit_invalidate_after_data_write(net):
    """Writes that bypass autograd's version counter (p.data.copy_) are invisible to the signatures: the documented
    escape hatch is pfd_b200.graphs.invalidate(module)."""
    from pfd_b200 import DDIMSampler
    from pfd_b200.graphs import invalidate
    sampler = DDIMSampler(net)
    a = _sample(net, sampler)
    p = net.diffuser["image"].data_blocks[0][0].weight
    saved = p.data.clone()
    p.data.mul_(1.5)
    invalidate(net)
    b = _sample(net, sampler)
    assert _rel(b, a) > 1e-3
    p.data.copy_(saved)
    invalidate(net)
    assert _rel(_sample(net, sampler), a) < 5e-3


# ------------------------------------------------------------------------------------------------ capture behaviour
def test_split_k_is_the_same_in_graph_replay_and_eager():
    """the split-K workspace used to be missing on the capture stream, so replayed graphs silently
    ran a different (unsplit) configuration than the eagerly tested one.  fp32 partial sums are deterministic, so the
    same configuration gives bit-identical outputs."""
    from pfd_b200 import native as nv
    g = torch.Generator().manual_seed(0)
    x = torch.randn((8, 8, 8, 1280), generator=g).cuda().half()
    w = (torch.randn((1280, 9 * 1280), generator=g) * 0.01).cuda().half()
    b = torch.randn((1280,), generator=g).cuda().half()
    eager = nv.conv3x3(x, w, b)
    torch.cuda.synchronize()
    out = torch.empty_like(eager)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        nv.conv3x3(x, w, b, out=out)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager)
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().reshape(1280, 3, 3, 1280).permute(0, 3, 1, 2),
                                     b.float(), padding=1).permute(0, 2, 3, 1)
    assert _rel(out, ref) < 2e-3


def test_whole_loop_graph_equals_step_graphs(net):
    """One captured graph for all DDIM steps (device-side step header) == one graph per step == eager."""
    from pfd_b200 import DDIMSampler
    a = _sample(net, DDIMSampler(net), steps=8)
    b = _sample(net, DDIMSampler(net, steps_per_graph=1), steps=8)
    c = _sample(net, DDIMSampler(net, steps_per_graph=4), steps=8)
    d = _sample(net, DDIMSampler(net, use_cuda_graph=False), steps=8)
    for o in (b, c, d):
        assert _rel(o, a) < 5e-3
    s = DDIMSampler(net)
    x1 = _sample(net, s, steps=1)                                          # single-step schedules are captured too
    assert torch.isfinite(x1.float()).all()


# ------------------------------------------------------------------------------------------------ f3: 1536 x 1536
def test_flash_attention_1536_level0_shape():
    """app.py:201-207 allows 1536x1536 requests: N = 192*192 = 36864 tokens at UNet level 0 (d = 40)."""
    from pfd_b200 import native as nv
    B, heads, N, d = 1, 2, 36864, 40
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn((B * heads, N, d), generator=g).cuda().half() for _ in range(3))
    vt = v.transpose(1, 2).contiguous()
    out = torch.empty((B, N, heads * d), device="cuda", dtype=torch.float16)
    nv.flash_attn(q, k, vt, B=B, heads=heads, Nq=N, Nk=N, scale=d ** -0.5, out=out)
    rows = torch.arange(0, N, 97, device="cuda")
    s = (q[:, rows].float() @ k.float().transpose(1, 2)).half().float() * d ** -0.5
    ref = (torch.softmax(s.half().float(), -1) @ v.float()).permute(1, 0, 2).reshape(len(rows), heads * d)
    assert _rel(out[0, rows], ref) < 5e-3


def test_unet_and_vae_at_1536(net):
    """One UNet evaluation at 192x192 latents (no CFG, batch 1) against the oracle run in fp16 on this GPU, and the
    VAE mid-block attention path at N = 36864 (the [B,N,N] score tensor is 2.7 GB per sample)."""
    from oracle import pfd_oracle as O
    g = torch.Generator().manual_seed(4)
    x = torch.randn((1, 4, 192, 192), generator=g).cuda().half()
    c = (0.5 * torch.randn((1, 148, 768), generator=g)).cuda().half()
    t = torch.tensor([501], device="cuda")
    e = net.apply_model({"type": "image", "x": x}, t, {"type": "image", "c": c, "control": None})
    sd = {k[len("diffuser.image."):]: v.detach() for k, v in net.state_dict().items() if k.startswith("diffuser.image.")}
    with torch.no_grad():
        ref = O.unet_apply(sd, O.UNET_SD15, x, t, c)
    r = _rel(e, ref)
    print(f"[parity] UNet eps at 192x192 latents vs oracle fp16 eager on GPU: rel_rms={r:.3e}")
    assert r < 1e-2
    del ref
    torch.cuda.empty_cache()
    z = torch.randn((1, 4, 192, 192), generator=g).cuda().half()
    im = net.vae["image"].decode(z, pre_scale=1.0 / 0.18215)
    assert im.shape == (1, 3, 1536, 1536) and torch.isfinite(im.float()).all()
    vsd = {k[len("vae.image."):]: v.detach() for k, v in net.state_dict().items() if k.startswith("vae.image.")}
    with torch.no_grad():
        ref = O.vae_decode(vsd, O.VAE_SD, z)
    r = _rel(im, ref)
    print(f"[parity] VAE decode 1536x1536 vs oracle fp16 eager on GPU: rel_rms={r:.3e}")
    assert r < 1e-2
