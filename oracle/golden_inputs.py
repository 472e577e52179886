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
