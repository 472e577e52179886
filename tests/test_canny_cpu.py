"""Pins the Canny oracle (oracle/canny_oracle.py) against the reference's own third-party call, cv2.Canny
(lib/model_zoo/controlnet_annotator/canny/__init__.py:4-5), bit-exactly — CPU only."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _images():
    rng = np.random.RandomState(0)
    yield "noise", rng.randint(0, 256, (97, 131, 3), dtype=np.uint8)
    smooth = cv2.GaussianBlur(rng.randint(0, 256, (160, 200, 3), dtype=np.uint8), (0, 0), 3)
    yield "smooth", smooth
    yy, xx = np.mgrid[0:128, 0:160]
    disk = ((yy - 60) ** 2 + (xx - 80) ** 2 < 40 ** 2).astype(np.uint8)
    img = np.stack([disk * 200 + 20, disk * 90 + 60, 255 - disk * 180], -1).astype(np.uint8)
    img[20:50, 30:120, 1] = 250                                          # a bar visible in one channel only
    yield "shapes", img
    grad = np.tile(np.linspace(0, 255, 256).astype(np.uint8)[None, :, None], (64, 1, 3))
    grad[:, 100:, 0] //= 2
    yield "ramp", grad
    yield "tiny", rng.randint(0, 256, (3, 5, 3), dtype=np.uint8)
    yield "flat", np.full((40, 40, 3), 128, dtype=np.uint8)


@pytest.mark.parametrize("low,high", [(100, 200), (50, 120), (200, 100)])
def test_oracle_matches_cv2_canny(low, high):
    from oracle.canny_oracle import canny_u8
    for name, img in _images():
        ref = cv2.Canny(img, low, high)
        out = canny_u8(img, low, high)
        assert out.shape == ref.shape and np.array_equal(out, ref), \
            f"{name}: {int((out != ref).sum())} of {ref.size} pixels differ (edges ref {int((ref > 0).sum())})"


def test_preprocess_canny_tensor_path_matches_reference_formula():
    import torch
    from oracle.canny_oracle import preprocess_canny, to_pil_u8
    x = torch.rand((2, 3, 64, 80), generator=torch.Generator().manual_seed(3))
    x = torch.nn.functional.avg_pool2d(x, 5, 1, 2)
    out = preprocess_canny(x, 20, 60)
    assert out.shape == (2, 3, 64, 80) and out.dtype == torch.float32
    for b in range(2):
        ref = cv2.Canny(to_pil_u8(x[b]), 20, 60)
        assert np.array_equal(out[b, 0].numpy() * 255, ref.astype(np.float32))
        assert torch.equal(out[b, 0], out[b, 1]) and torch.equal(out[b, 0], out[b, 2])
    assert out.sum() > 0
