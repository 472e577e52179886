"""Degree-3 polynomial for 2^f on [-0.5, 0.5] evaluated in fp16 Horner form (attention.cu: exp2_poly_h2)."""
import numpy as np

f = np.cos(np.pi * (np.arange(4000) + 0.5) / 4000) * 0.5
c = np.linalg.lstsq(np.vander(f, 4, increasing=True), 2.0 ** f, rcond=None)[0]
print("coefficients c0..c3:", c)
ff = np.linspace(-0.5, 0.5, 20001).astype(np.float16)
p = (np.float16(c[3]) * ff + np.float16(c[2])).astype(np.float16)
p = (p * ff + np.float16(c[1])).astype(np.float16)
p = (p * ff + np.float16(c[0])).astype(np.float16)
rel = np.abs(p.astype(np.float64) / 2.0 ** ff.astype(np.float64) - 1)
print(f"fp16 Horner: max rel err {rel.max():.2e}, rms {np.sqrt((rel ** 2).mean()):.2e}")
