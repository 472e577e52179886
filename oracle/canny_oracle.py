"""CPU restatement of the Canny pre-processing step of ControlNet.preprocess.  TEST INFRASTRUCTURE — only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Reference path: lib/model_zoo/controlnet.py:332-360 (tensor -> ToPILImage -> apply_canny -> ToTensor ->
repeat(1,3,1,1) -> float32) and lib/model_zoo/controlnet_annotator/canny/__init__.py:4-5, which is one call into a
third-party dependency that is NOT vendored in /root/reference: `cv2.Canny(img, low, high)` of opencv-python
(requirements.txt:8 `opencv-python==4.7.0.72`; this image ships cv2 4.13.0).  The published algorithm
(OpenCV modules/imgproc/src/canny.cpp, aperture 3, L2gradient=false) is restated below in numpy; it is PINNED
against cv2.Canny itself on random and structured images by tests/test_canny_cpu.py (bit-exact).
"""
import numpy as np

TG22 = 13573          # (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5)
SHIFT = 15


def to_pil_u8(x):
    """torchvision ToPILImage on a float [3,H,W] tensor: pic.mul(255).byte() -> HxWx3 uint8 (controlnet.py:339)."""
    import torch
    return x.mul(255).byte().permute(1, 2, 0).contiguous().cpu().numpy()


def _sobel(img):
    """3x3 Sobel dx, dy of an HxWxC uint8 image with BORDER_REPLICATE -> int32 HxWxC."""
    p = np.pad(img.astype(np.int32), ((1, 1), (1, 1), (0, 0)), mode="edge")
    dx = (p[:-2, 2:] - p[:-2, :-2]) + 2 * (p[1:-1, 2:] - p[1:-1, :-2]) + (p[2:, 2:] - p[2:, :-2])
    dy = (p[2:, :-2] - p[:-2, :-2]) + 2 * (p[2:, 1:-1] - p[:-2, 1:-1]) + (p[2:, 2:] - p[:-2, 2:])
    return dx, dy


def canny_u8(img, low=100, high=200):
    """cv2.Canny(img, low, high) for an HxW or HxWxC uint8 image -> HxW uint8 in {0, 255}."""
    if img.ndim == 2:
        img = img[:, :, None]
    low, high = int(np.floor(low)), int(np.floor(high))
    if low > high:
        low, high = high, low
    H, W, C = img.shape
    dx, dy = _sobel(img)
    mag = np.abs(dx) + np.abs(dy)
    best = np.argmax(mag, axis=2)                      # first maximum on ties, like the `>` scan in canny.cpp
    ii, jj = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    dx, dy, mag = dx[ii, jj, best], dy[ii, jj, best], mag[ii, jj, best]
    mp = np.pad(mag, 1)                                # magnitude is zero outside the image
    c = mp[1:-1, 1:-1]
    ax, ay = np.abs(dx).astype(np.int64), np.abs(dy).astype(np.int64) << SHIFT
    tg22 = ax * TG22
    tg67 = tg22 + (ax << (SHIFT + 1))
    horiz = ay < tg22
    vert = (~horiz) & (ay > tg67)
    diag = ~(horiz | vert)
    s_neg = (dx ^ dy) < 0                              # s = -1: compare (y-1, x+1) and (y+1, x-1)
    left, right = mp[1:-1, :-2], mp[1:-1, 2:]
    up, down = mp[:-2, 1:-1], mp[2:, 1:-1]
    ul, ur, dl, dr = mp[:-2, :-2], mp[:-2, 2:], mp[2:, :-2], mp[2:, 2:]
    cand = np.where(horiz, (c > left) & (c >= right),
                    np.where(vert, (c > up) & (c >= down),
                             np.where(s_neg, (c > ur) & (c > dl), (c > ul) & (c > dr))))
    cand &= c > low
    strong = cand & (c > high)
    # hysteresis: candidates 8-connected to a strong pixel
    out = strong.copy()
    stack = list(zip(*np.nonzero(strong)))
    while stack:
        y, x = stack.pop()
        for yy in (y - 1, y, y + 1):
            for xx in (x - 1, x, x + 1):
                if 0 <= yy < H and 0 <= xx < W and cand[yy, xx] and not out[yy, xx]:
                    out[yy, xx] = True
                    stack.append((yy, xx))
    return (out.astype(np.uint8) * 255)


def preprocess_canny(x, low=100, high=200):
    """ControlNet.preprocess(x, type='canny') for a float [B,3,H,W] torch tensor -> float32 [B,3,H,W]."""
    import torch
    ys = [torch.from_numpy(canny_u8(to_pil_u8(xi), low, high)).float().div(255)[None] for xi in x]
    return torch.stack(ys).repeat(1, 3, 1, 1).to(torch.float32)
