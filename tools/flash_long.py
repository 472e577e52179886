"""Long-sequence check of the flash-attention kernel: level-0 self-attention of a 1536x1536 request (N = 36864
tokens, d = 40) - parity of one head against fp32 torch (chunked over queries) + device time
of one CFG pair (2 x 8 heads)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfd_b200 import native as nv
from tools.gemm_perf import timeit

torch.manual_seed(0)
N, d, heads, B = 36864, 40, 8, 2
q = torch.randn(B * heads, N, d, device="cuda").half()
k = torch.randn(B * heads, N, d, device="cuda").half()
vt = torch.randn(B * heads, d, N, device="cuda").half()
out = torch.empty(B, N, heads * d, device="cuda", dtype=torch.float16)
scale = d ** -0.5
nv.flash_attn(q, k, vt, B=B, heads=heads, Nq=N, Nk=N, scale=scale, out=out)
torch.cuda.synchronize()
# reference for (b=1, h=3), queries in chunks of 2048
b, h = 1, 3
qq, kk, vv = q[b * heads + h].float(), k[b * heads + h].float(), vt[b * heads + h].float().t()
ref = torch.empty(N, d, device="cuda")
for s in range(0, N, 2048):
    p = torch.softmax((qq[s:s + 2048] @ kk.t()) * scale, -1)
    ref[s:s + 2048] = p @ vv
got = out[b, :, h * d:(h + 1) * d].float()
err = (got - ref).abs().max().item()
rel = ((got - ref).pow(2).mean() / ref.pow(2).mean()).sqrt().item()
ms = timeit(lambda: nv.flash_attn(q, k, vt, B=B, heads=heads, Nq=N, Nk=N, scale=scale, out=out), n=3)
print(json.dumps(dict(N=N, d=d, bh=B * heads, max_abs_err=err, rel_rms=rel, ms=ms,
                      tflops=4.0 * B * heads * N * N * d / ms / 1e9)))
assert rel < 5e-3
