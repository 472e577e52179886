"""Seeded inputs of the golden fixtures (shared by tools/make_golden.py and the tests). TEST INFRASTRUCTURE."""
import torch


def seeded(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


def golden_inputs():
    g = torch.Generator().manual_seed(1234)
    return dict(
        x=seeded((2, 4, 16, 16), 11),
        t=torch.tensor([501, 501], dtype=torch.long),
        ctx=seeded((2, 148, 768), 12, 0.5),
        hint=(torch.rand((1, 3, 128, 128), generator=g) > 0.9).float(),
        z=seeded((1, 4, 8, 8), 13),
        img=torch.rand((1, 3, 128, 128), generator=torch.Generator().manual_seed(14)),
        x_T=seeded((1, 4, 16, 16), 15),
        cond=seeded((1, 148, 768), 16, 0.5),
    )


def _rand(shape, seed):
    return torch.rand(shape, generator=torch.Generator().manual_seed(seed))


def config_inputs():
    """Seeded inputs of the BASELINE-config-size fixtures (tests/golden/config_outputs.npz, written by
    tools/make_golden_configs.py from the UNMODIFIED reference)."""
    anime_ug = torch.cat([seeded((77, 768), 33, 0.5), torch.zeros((71, 768))], 0)[None]      # app.py:238-241
    edge = (_rand((1, 1, 512, 512), 35) > 0.9).float().repeat(1, 3, 1, 1)                   # canny-like binary map
    return dict(
        # config 1: 256x256 reference image -> SeeCoder -> 10 DDIM steps CFG 2.0 @ [1,4,64,64] -> VAE 512x512
        c1_img=_rand((1, 3, 256, 256), 1), c1_xT=seeded((1, 4, 64, 64), 20),
        # config 2 size: teacher-forced eps, B=4 (CFG batch 8) at 64x64, t in {981, 501, 1}
        c2_x=seeded((4, 4, 64, 64), 31), c2_cond=seeded((1, 148, 768), 32, 0.5), c2_t=(981, 501, 1),
        # config 3: unconditional context = [77,768] tensor zero-padded to 148 tokens
        c3_x=seeded((2, 4, 64, 64), 34), c3_cond=seeded((1, 148, 768), 39, 0.5), c3_uncond=anime_ug, c3_t=501,
        # config 4: ControlNet at 64x64 latents with a 512x512 hint, B=2 (CFG batch 4)
        c4_x=seeded((2, 4, 64, 64), 36), c4_cond=seeded((1, 148, 768), 37, 0.5), c4_hint=edge, c4_t=501,
        # config 5: PPE_MLP + 768x768 reference image + 96x96 latents, 31-entry schedule, 2 teacher-forced steps
        c5_img=_rand((1, 3, 768, 768), 23), c5_xT=seeded((1, 4, 96, 96), 38),
        # SeeCoder at 512x512 (padded windows at every level)
        c6_img=_rand((1, 3, 512, 512), 40),
        # eta = 0.5: x_T + per-step noise tensors, 4 steps at 16x16, CFG 2.0
        c7_xT=seeded((1, 4, 16, 16), 41), c7_cond=seeded((1, 148, 768), 42, 0.5),
        c7_noise=[seeded((1, 4, 16, 16), 50 + i) for i in range(4)],
        # VAE encode (img2img / image variation): 256x256 image -> posterior [1,4,32,32]
        c8_img=_rand((1, 3, 256, 256), 60),
        # sample_multicontext: two contexts mixed 0.3 / 0.7, 4 steps at 16x16, CFG 2.0
        c9_xT=seeded((1, 4, 16, 16), 61), c9_cond_a=seeded((1, 148, 768), 62, 0.5),
        c9_cond_b=seeded((1, 148, 768), 63, 0.5),
    )
