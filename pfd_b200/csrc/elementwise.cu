// HBM-bound kernels of the Prompt-Free-Diffusion hot path (channel-last fp16, fp32 math):
// GroupNorm(+SiLU, optional two-source concat), LayerNorm(+residual), row softmax with the
// reference's fp16 score rounding, timestep embedding, nearest 2x upsample, layout converts,
// small-Cin im2col, fused CFG + DDIM update, Swin window gather/scatter and patch-merge gather.
// All loads/stores are 128-bit vectorised where the layout allows; reductions use warp shuffles.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/pfd_b200.h"
#include "common.h"

namespace pfd {

__device__ __forceinline__ void pdl_enter() {
  // programmatic dependent launch: block until the producer kernel has completed, then let the
  // consumer kernel start scheduling its CTAs (see common.h: launch_k)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
__device__ __forceinline__ float rh(float v) { return __half2float(__float2half_rn(v)); }

// ------------------------------------------------------------------------------------ GroupNorm
// Thread mapping shared by both passes: blockDim is a multiple of the number of 8-channel vectors
// (vecs = C/8) so a thread always owns the same channel vector and strides over pixels ->
// consecutive threads read consecutive 16-byte vectors of one pixel (fully coalesced), per-channel
// partial sums / affine coefficients live in registers, and nothing is recomputed per element.
//   pass 1 (gn_stats): per-(image, group) sum / sum of squares -> fp64 atomics (one per group per CTA)
//   pass 2 (gn_apply): y = x * a[c] + b[c] (a = rstd*gamma, b = beta - mean*a) [+ SiLU] -> fp16
constexpr int GN_MAX_GROUPS = 32;

__device__ __forceinline__ uint4 gn_load(const __half* __restrict__ x1, int c1, const __half* __restrict__ x2,
                                         int c2, long long pixn, int c) {
  if (c < c1) return __ldg(reinterpret_cast<const uint4*>(x1 + pixn * c1 + c));
  return __ldg(reinterpret_cast<const uint4*>(x2 + pixn * c2 + (c - c1)));
}

__global__ void __launch_bounds__(320)
gn_stats_kernel(const __half* __restrict__ x1, int c1, const __half* __restrict__ x2, int c2,
                long long HW, int groups, long long pix_per_cta, double* __restrict__ ws) {
  pdl_enter();
  const int C = c1 + c2;
  const int cpg = C / groups;
  const int vecs = C / 8;
  const int n = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * pix_per_cta;
  long long p1 = p0 + pix_per_cta;
  if (p1 > HW) p1 = HW;
  __shared__ float s_sum[GN_MAX_GROUPS];
  __shared__ float s_sq[GN_MAX_GROUPS];
  if (threadIdx.x < GN_MAX_GROUPS) {
    s_sum[threadIdx.x] = 0.f;
    s_sq[threadIdx.x] = 0.f;
  }
  __syncthreads();
  const int lanes = blockDim.x / vecs;      // pixel lanes per CTA (blockDim % vecs == 0, or vecs > blockDim)
  if (lanes >= 1) {
    const int v = threadIdx.x % vecs;
    const int c = v * 8;
    float sm[8], sq[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sm[i] = sq[i] = 0.f;
    long long pix = p0 + threadIdx.x / vecs;
    for (; pix + 3 * lanes < p1; pix += 4 * lanes) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = gn_load(x1, c1, x2, c2, (long long)n * HW + pix + k * lanes, c);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float f[8];
        unpack8(u[k], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          sm[i] += f[i];
          sq[i] += f[i] * f[i];
        }
      }
    }
    for (; pix < p1; pix += lanes) {
      float f[8];
      unpack8(gn_load(x1, c1, x2, c2, (long long)n * HW + pix, c), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sm[i] += f[i];
        sq[i] += f[i] * f[i];
      }
    }
    // fold the 8 channels into (at most two) group bins, then one shared atomic per bin
    int g_prev = c / cpg;
    float as = 0.f, aq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int g = (c + i) / cpg;
      if (g != g_prev) {
        atomicAdd(&s_sum[g_prev], as);
        atomicAdd(&s_sq[g_prev], aq);
        as = aq = 0.f;
        g_prev = g;
      }
      as += sm[i];
      aq += sq[i];
    }
    atomicAdd(&s_sum[g_prev], as);
    atomicAdd(&s_sq[g_prev], aq);
  } else {
    // very wide rows (vecs > blockDim): a thread walks several vectors of each pixel
    for (long long pix = p0; pix < p1; ++pix) {
      for (int v = threadIdx.x; v < vecs; v += blockDim.x) {
        const int c = v * 8;
        float f[8];
        unpack8(gn_load(x1, c1, x2, c2, (long long)n * HW + pix, c), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int g = (c + i) / cpg;
          atomicAdd(&s_sum[g], f[i]);
          atomicAdd(&s_sq[g], f[i] * f[i]);
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    atomicAdd(&ws[((long long)n * groups + threadIdx.x) * 2 + 0], (double)s_sum[threadIdx.x]);
    atomicAdd(&ws[((long long)n * groups + threadIdx.x) * 2 + 1], (double)s_sq[threadIdx.x]);
  }
}

__global__ void __launch_bounds__(320)
gn_apply_kernel(const __half* __restrict__ x1, int c1, const __half* __restrict__ x2, int c2,
                long long HW, int groups, const __half* __restrict__ gamma,
                const __half* __restrict__ beta, float eps, int silu,
                const double* __restrict__ ws, __half* __restrict__ out, long long pix_per_cta,
                double inv_cnt) {
  pdl_enter();
  const int C = c1 + c2;
  const int cpg = C / groups;
  const int vecs = C / 8;
  const int n = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * pix_per_cta;
  long long p1 = p0 + pix_per_cta;
  if (p1 > HW) p1 = HW;
  const int lanes = blockDim.x / vecs;
  const int vstep = lanes >= 1 ? vecs : blockDim.x;
  const int pstep = lanes >= 1 ? lanes : 1;
  for (int v = threadIdx.x % vstep; v < vecs; v += vstep) {
    const int c = v * 8;
    float a8[8], b8[8];
    {
      float g8[8], be8[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(gamma + c)), g8);
      unpack8(__ldg(reinterpret_cast<const uint4*>(beta + c)), be8);
      int gprev = -1;
      float mean = 0.f, rstd = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int g = (c + i) / cpg;
        if (g != gprev) {
          // fp64 only for the cancellation-prone E[x^2] - mean^2 (3 DP ops, no DP division)
          const double m = ws[((long long)n * groups + g) * 2 + 0] * inv_cnt;
          double var = ws[((long long)n * groups + g) * 2 + 1] * inv_cnt - m * m;
          if (var < 0) var = 0;
          mean = (float)m;
          rstd = rsqrtf((float)var + eps);
          gprev = g;
        }
        a8[i] = rstd * g8[i];
        b8[i] = be8[i] - mean * a8[i];
      }
    }
    auto emit = [&](long long pixn, const uint4& u) {
      float f[8];
      unpack8(u, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float y = fmaf(f[i], a8[i], b8[i]);
        if (silu) {
          y = rh(y);  // reference materialises the GroupNorm output in fp16 before SiLU
          y = __fdividef(y, 1.f + __expf(-y));
        }
        f[i] = y;
      }
      *reinterpret_cast<uint4*>(out + pixn * C + c) = pack8(f);
    };
    long long pix = p0 + (lanes >= 1 ? threadIdx.x / vecs : 0);
    const long long base = (long long)n * HW;
    for (; pix + 3 * pstep < p1; pix += 4 * pstep) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = gn_load(x1, c1, x2, c2, base + pix + k * pstep, c);
#pragma unroll
      for (int k = 0; k < 4; ++k) emit(base + pix + k * pstep, u[k]);
    }
    for (; pix < p1; pix += pstep) emit(base + pix, gn_load(x1, c1, x2, c2, base + pix, c));
  }
}

// Single-pass GroupNorm: each CTA keeps its pixel chunk in shared memory between the statistics phase and
// the apply phase, so the activation is read from L2/HBM ONCE (the two-kernel version reads it twice) and
// one launch disappears.  The phases are separated by a per-image arrival counter in global memory; this is
// only used when every CTA of the grid can be co-resident (host checks with the occupancy API), and the
// dependent kernel is not allowed to start launching before the barrier has been passed.
__global__ void __launch_bounds__(320)
gn_fused_kernel(const __half* __restrict__ x1, int c1, const __half* __restrict__ x2, int c2, long long HW,
                int groups, const __half* __restrict__ gamma, const __half* __restrict__ beta, float eps,
                int silu, double* __restrict__ ws, unsigned int* __restrict__ counter, __half* __restrict__ out,
                long long pix_per_cta, double inv_cnt) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  extern __shared__ uint4 gn_tile[];
  const int C = c1 + c2;
  const int cpg = C / groups;
  const int vecs = C / 8;
  const int n = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * pix_per_cta;
  long long p1 = p0 + pix_per_cta;
  if (p1 > HW) p1 = HW;
  __shared__ float s_sum[GN_MAX_GROUPS];
  __shared__ float s_sq[GN_MAX_GROUPS];
  if (threadIdx.x < GN_MAX_GROUPS) {
    s_sum[threadIdx.x] = 0.f;
    s_sq[threadIdx.x] = 0.f;
  }
  __syncthreads();
  const int lanes = blockDim.x / vecs;          // host guarantees lanes >= 1 for this kernel
  const int v = threadIdx.x % vecs;
  const int c = v * 8;
  const long long base = (long long)n * HW;
  const bool active = threadIdx.x < lanes * vecs;
  if (active) {
    float sm[8], sq[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sm[i] = sq[i] = 0.f;
    long long pix = p0 + threadIdx.x / vecs;
    for (; pix + 3 * lanes < p1; pix += 4 * lanes) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = gn_load(x1, c1, x2, c2, base + pix + k * lanes, c);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        gn_tile[(pix + k * lanes - p0) * vecs + v] = u[k];
        float f[8];
        unpack8(u[k], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          sm[i] += f[i];
          sq[i] += f[i] * f[i];
        }
      }
    }
    for (; pix < p1; pix += lanes) {
      const uint4 u = gn_load(x1, c1, x2, c2, base + pix, c);
      gn_tile[(pix - p0) * vecs + v] = u;
      float f[8];
      unpack8(u, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sm[i] += f[i];
        sq[i] += f[i] * f[i];
      }
    }
    int g_prev = c / cpg;
    float as = 0.f, aq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int g = (c + i) / cpg;
      if (g != g_prev) {
        atomicAdd(&s_sum[g_prev], as);
        atomicAdd(&s_sq[g_prev], aq);
        as = aq = 0.f;
        g_prev = g;
      }
      as += sm[i];
      aq += sq[i];
    }
    atomicAdd(&s_sum[g_prev], as);
    atomicAdd(&s_sq[g_prev], aq);
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    atomicAdd(&ws[((long long)n * groups + threadIdx.x) * 2 + 0], (double)s_sum[threadIdx.x]);
    atomicAdd(&ws[((long long)n * groups + threadIdx.x) * 2 + 1], (double)s_sq[threadIdx.x]);
  }
  // ---- per-image barrier over the gridDim.x CTAs of image n
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(&counter[n], 1u);
    unsigned int spins = 0;
    while (*reinterpret_cast<volatile unsigned int*>(&counter[n]) < gridDim.x) {
      __nanosleep(64);
      if (++spins > (1u << 22)) asm volatile("trap;");
    }
    __threadfence();
  }
  __syncthreads();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (!active) return;
  // ---- apply from shared memory
  float a8[8], b8[8];
  {
    float g8[8], be8[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(gamma + c)), g8);
    unpack8(__ldg(reinterpret_cast<const uint4*>(beta + c)), be8);
    int gprev = -1;
    float mean = 0.f, rstd = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int g = (c + i) / cpg;
      if (g != gprev) {
        const double m = __ldcg(&ws[((long long)n * groups + g) * 2 + 0]) * inv_cnt;
        double var = __ldcg(&ws[((long long)n * groups + g) * 2 + 1]) * inv_cnt - m * m;
        if (var < 0) var = 0;
        mean = (float)m;
        rstd = rsqrtf((float)var + eps);
        gprev = g;
      }
      a8[i] = rstd * g8[i];
      b8[i] = be8[i] - mean * a8[i];
    }
  }
  for (long long pix = p0 + threadIdx.x / vecs; pix < p1; pix += lanes) {
    float f[8];
    unpack8(gn_tile[(pix - p0) * vecs + v], f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y = fmaf(f[i], a8[i], b8[i]);
      if (silu) {
        y = rh(y);
        y = __fdividef(y, 1.f + __expf(-y));
      }
      f[i] = y;
    }
    *reinterpret_cast<uint4*>(out + (base + pix) * C + c) = pack8(f);
  }
}

// ------------------------------------------------------------------------------------ LayerNorm
// one warp per row; C % 8 == 0; row cached in registers (C <= 8*32*MAXV).
template <int MAXV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __half* __restrict__ x, const __half* __restrict__ res, long long rows, int C,
                 const __half* __restrict__ gamma, const __half* __restrict__ beta, float eps,
                 __half* __restrict__ out) {
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const int vecs = C / 8;
  float f[MAXV][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int v = lane + k * 32;
    if (v < vecs) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + row * C + v * 8)), f[k]);
      if (res) {
        float r8[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(res + row * C + v * 8)), r8);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[k][i] = rh(f[k][i] + r8[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) s += f[k][i];
    }
  }
  s = warp_sum(s);
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int v = lane + k * 32;
    if (v < vecs) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = f[k][i] - mean;
        q += d * d;
      }
    }
  }
  q = warp_sum(q);
  const float rstd = rsqrtf(q / C + eps);
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int v = lane + k * 32;
    if (v < vecs) {
      float g8[8], b8[8], o[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(gamma + v * 8)), g8);
      unpack8(__ldg(reinterpret_cast<const uint4*>(beta + v * 8)), b8);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (f[k][i] - mean) * rstd * g8[i] + b8[i];
      *reinterpret_cast<uint4*>(out + row * C + v * 8) = pack8(o);
    }
  }
}

// ------------------------------------------------------------------------------------ softmax
// one warp per row (cols <= 32*MAXE) or one CTA per row for long rows.
__global__ void __launch_bounds__(256)
softmax_rows_kernel(__half* __restrict__ s, long long batch, int rows, int cols, long long ld,
                    float scale, const __half* __restrict__ bias, int nheads,
                    const __half* __restrict__ mask, int nwin) {
  pdl_enter();
  // CTA per row, threads stride over columns; values cached in shared memory as fp32.
  extern __shared__ float sv[];
  const long long r = blockIdx.x;  // global row = b*rows + i
  const long long b = r / rows;
  const int i = (int)(r % rows);
  __half* row = s + r * ld;
  const __half* brow = bias ? bias + ((long long)(b % nheads) * rows + i) * cols : nullptr;
  const __half* mrow = mask ? mask + ((long long)((b / nheads) % nwin) * rows + i) * cols : nullptr;
  __shared__ float red[32];
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float v = __half2float(row[c]);
    v = rh(v * scale);  // reference: fp16 score * scale stays an fp16 tensor
    if (brow) v = rh(v + __half2float(brow[c]));
    if (mrow) v = rh(v + __half2float(mrow[c]));
    sv[c] = v;
    mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
    t = warp_max(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  mx = red[0];
  __syncthreads();
  float sum = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float e = __expf(sv[c] - mx);
    sv[c] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  const float inv = 1.f / red[0];
  for (int c = threadIdx.x; c < cols; c += blockDim.x) row[c] = __float2half_rn(sv[c] * inv);
}

// ------------------------------------------------------------------------------------ misc
__global__ void timestep_embedding_kernel(const long long* __restrict__ t, int n, int dim,
                                          float max_period, __half* __restrict__ out) {
  pdl_enter();
  const int half_d = dim / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * half_d) return;
  const int b = idx / half_d, i = idx % half_d;
  // diffusion_utils.py:141-146: freqs = exp(-ln(max_period) * i / half) in fp32
  const float freq = expf(-logf(max_period) * (float)i / (float)half_d);
  const float arg = (float)t[b] * freq;
  out[(long long)b * dim + i] = __float2half_rn(cosf(arg));
  out[(long long)b * dim + half_d + i] = __float2half_rn(sinf(arg));
  if ((dim & 1) && i == 0) out[(long long)b * dim + dim - 1] = __float2half_rn(0.f);
}

__global__ void upsample2x_kernel(const uint4* __restrict__ x, int NB, int H, int W, int vecs,
                                  uint4* __restrict__ out) {
  pdl_enter();
  const long long total = (long long)NB * (2 * H) * (2 * W) * vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long p = i / vecs;
    const int ox = (int)(p % (2 * W));
    p /= (2 * W);
    const int oy = (int)(p % (2 * H));
    const int n = (int)(p / (2 * H));
    out[i] = __ldg(&x[(((long long)n * H + (oy >> 1)) * W + (ox >> 1)) * vecs + v]);
  }
}

template <typename T>
__global__ void nchw_to_nhwc_kernel(const T* __restrict__ x, int NB, int C, int H, int W, int Cpad,
                                    __half* __restrict__ out) {
  pdl_enter();
  const long long total = (long long)NB * H * W * Cpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cpad);
    long long p = i / Cpad;
    const int xw = (int)(p % W);
    p /= W;
    const int y = (int)(p % H);
    const int n = (int)(p / H);
    float v = 0.f;
    if (c < C) v = (float)x[(((long long)n * C + c) * H + y) * W + xw];
    out[i] = __float2half_rn(v);
  }
}

__global__ void nhwc_to_nchw_kernel(const __half* __restrict__ x, int NB, int C, int H, int W,
                                    int Cpad, float mul, float add, float lo, float hi,
                                    __half* __restrict__ out) {
  pdl_enter();
  const long long total = (long long)NB * C * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int xw = (int)(i % W);
    long long p = i / W;
    const int y = (int)(p % H);
    p /= H;
    const int c = (int)(p % C);
    const int n = (int)(p / C);
    float v = __half2float(x[(((long long)n * H + y) * W + xw) * Cpad + c]);
    v = rh(v * mul + add);
    v = fminf(fmaxf(v, lo), hi);
    out[i] = __float2half_rn(v);
  }
}

__global__ void im2col3x3_kernel(const __half* __restrict__ x, int NB, int H, int W, int C,
                                 int stride, int Ho, int Wo, int Kpad, __half* __restrict__ out) {
  pdl_enter();
  const long long total = (long long)NB * Ho * Wo * Kpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kpad);
    long long p = i / Kpad;
    const int ox = (int)(p % Wo);
    p /= Wo;
    const int oy = (int)(p % Ho);
    const int n = (int)(p / Ho);
    __half v = __float2half_rn(0.f);
    if (k < 9 * C) {
      const int tap = k / C, c = k % C;
      const int iy = oy * stride + tap / 3 - 1;
      const int ix = ox * stride + tap % 3 - 1;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[(((long long)n * H + iy) * W + ix) * C + c];
    }
    out[i] = v;
  }
}

__global__ void axpby_kernel(const __half* __restrict__ a, float sa, const __half* __restrict__ b,
                             float sb, long long n, __half* __restrict__ out) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float v = __half2float(a[i]) * sa;
    if (b) v += __half2float(b[i]) * sb;
    out[i] = __float2half_rn(v);
  }
}

__global__ void add_rowvec_kernel(const uint4* __restrict__ a, const uint4* __restrict__ row,
                                  long long rows, int vecs, uint4* __restrict__ out) {
  pdl_enter();
  const long long total = rows * vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    float f[8], r[8];
    unpack8(__ldg(&a[i]), f);
    unpack8(__ldg(&row[i % vecs]), r);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] += r[k];
    out[i] = pack8(f);
  }
}

// CFG combine + DDIM update with the reference's fp16 rounding sequence (ddim.py:150-171).
__global__ void ddim_step_kernel(const __half* __restrict__ eps, const __half* __restrict__ x,
                                 long long half_n, float guidance, const float* __restrict__ coef,
                                 const int* __restrict__ step, __half* __restrict__ x_prev,
                                 __half* __restrict__ pred_x0) {
  pdl_enter();
  const int st = step ? *step : 0;
  // torch.full(..., dtype=fp16) rounds each coefficient to fp16 first (ddim.py:160-163)
  const float a_t = rh(coef[st * 4 + 0]);
  const float a_prev = rh(coef[st * 4 + 1]);
  const float sigma = rh(coef[st * 4 + 2]);
  const float s1m = rh(coef[st * 4 + 3]);
  const float sqrt_at = rh(sqrtf(a_t));
  const float sqrt_ap = rh(sqrtf(a_prev));
  const float dir_c = rh(sqrtf(rh(rh(1.f - a_prev) - rh(sigma * sigma))));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < half_n;
       i += (long long)gridDim.x * blockDim.x) {
    const float eu = __half2float(eps[i]);
    const float ec = __half2float(eps[half_n + i]);
    // e_t = e_u + s * (e_c - e_u), each op rounded to fp16 like the eager reference
    const float e = rh(eu + rh(guidance * rh(ec - eu)));
    const float xv = __half2float(x[i]);
    const float p0 = rh(rh(xv - rh(s1m * e)) / sqrt_at);
    const float dir = rh(dir_c * e);
    const float xp = rh(rh(sqrt_ap * p0) + dir);
    x_prev[i] = __float2half_rn(xp);
    if (pred_x0) pred_x0[i] = __float2half_rn(p0);
  }
}

// Swin: pad to multiples of ws, cyclic shift by -shift, partition into windows (swin.py:269-287).
// out[(b*nWh + wy)*nWw + wx][iy*ws+ix][c] = xpad[b, (wy*ws+iy+shift)%Hp, (wx*ws+ix+shift)%Wp, c]
__global__ void window_gather_kernel(const uint4* __restrict__ x, int B, int H, int W, int vecs,
                                     int ws, int shift, int Hp, int Wp, uint4* __restrict__ out) {
  pdl_enter();
  const long long total = (long long)B * Hp * Wp * vecs;
  const int nWw = Wp / ws, nWh = Hp / ws;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long p = i / vecs;
    const int t = (int)(p % (ws * ws));
    p /= (ws * ws);
    const int wx = (int)(p % nWw);
    p /= nWw;
    const int wy = (int)(p % nWh);
    const int b = (int)(p / nWh);
    const int sy = (wy * ws + t / ws + shift) % Hp;
    const int sx = (wx * ws + t % ws + shift) % Wp;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (sy < H && sx < W) u = __ldg(&x[(((long long)b * H + sy) * W + sx) * vecs + v]);
    out[i] = u;
  }
}
// inverse: out[b,y,x,:] = residual[b,y,x,:] + win[window(y,x)]   (swin.py:289-304,307)
__global__ void window_scatter_kernel(const uint4* __restrict__ win, int B, int H, int W, int vecs,
                                      int ws, int shift, int Hp, int Wp,
                                      const uint4* __restrict__ residual, uint4* __restrict__ out) {
  pdl_enter();
  const long long total = (long long)B * H * W * vecs;
  const int nWw = Wp / ws, nWh = Hp / ws;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long p = i / vecs;
    const int xw = (int)(p % W);
    p /= W;
    const int y = (int)(p % H);
    const int b = (int)(p / H);
    // x[y] = shifted[(y - shift) mod Hp]
    const int sy = (y - shift + Hp) % Hp;
    const int sx = (xw - shift + Wp) % Wp;
    const int wy = sy / ws, iy = sy % ws, wx = sx / ws, ix = sx % ws;
    const long long widx = (((long long)b * nWh + wy) * nWw + wx) * (ws * ws) + iy * ws + ix;
    float f[8];
    unpack8(__ldg(&win[widx * vecs + v]), f);
    if (residual) {
      float r[8];
      unpack8(__ldg(&residual[i]), r);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] += r[k];
    }
    out[i] = pack8(f);
  }
}

// PatchMerging gather (swin.py:336-346): pad odd H/W with zeros, concat [x0,x1,x2,x3] =
// [(0,0),(1,0),(0,1),(1,1)] (dy,dx) along channels.
__global__ void patch_merge_kernel(const uint4* __restrict__ x, int B, int H, int W, int vecs,
                                   uint4* __restrict__ out) {
  pdl_enter();
  const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
  const long long total = (long long)B * H2 * W2 * 4 * vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long p = i / vecs;
    const int qd = (int)(p % 4);
    p /= 4;
    const int ox = (int)(p % W2);
    p /= W2;
    const int oy = (int)(p % H2);
    const int b = (int)(p / H2);
    const int dy = qd & 1, dx = qd >> 1;  // x0:(0,0) x1:(1,0) x2:(0,1) x3:(1,1)
    const int y = 2 * oy + dy, xw = 2 * ox + dx;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (y < H && xw < W) u = __ldg(&x[(((long long)b * H + y) * W + xw) * vecs + v]);
    out[i] = u;
  }
}


// PatchEmbed gather (swin.py:479-489): zero-pad H,W to multiples of P, then
// out[b, py, px, (c*P + dy)*P + dx] = img[b, c, py*P+dy, px*P+dx]  (K order == flattened conv weight)
template <typename T>
__global__ void patchify_kernel(const T* __restrict__ x, int B, int C, int H, int W, int P, int Kpad,
                                __half* __restrict__ out) {
  pdl_enter();
  const int Hp = (H + P - 1) / P, Wp = (W + P - 1) / P;
  const long long total = (long long)B * Hp * Wp * Kpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kpad);
    long long p = i / Kpad;
    const int px = (int)(p % Wp);
    p /= Wp;
    const int py = (int)(p % Hp);
    const int b = (int)(p / Hp);
    float v = 0.f;
    if (k < C * P * P) {
      const int dx = k % P, dy = (k / P) % P, c = k / (P * P);
      const int y = py * P + dy, xx = px * P + dx;
      if (y < H && xx < W) v = (float)x[(((long long)b * C + c) * H + y) * W + xx];
    }
    out[i] = __float2half_rn(v);
  }
}

// ------------------------------------------------------------------------------------------------
// Single-pass GroupNorm(+SiLU), CS = 8: on a thread-block cluster.  One cluster of GNC_CS CTAs owns (image n, G consecutive
// groups): CTA r keeps pixels [r*npc, (r+1)*npc) x (G*cpg channels) in shared memory, the group statistics are
// reduced across the cluster through distributed shared memory (mean first, then the centred second moment from
// the cached slice: numerically the textbook two-pass form in fp32), and the slice is normalised straight out of
// shared memory.  One HBM read + one write and ONE launch per GroupNorm (the stats + apply pair cost two ~5 us
// launch floors and a second read; the grid-wide spin barrier of gn_fused_kernel cost more than it saved).
// CS = 1: the same kernel without any cross-CTA step - one CTA owns ALL pixels of (image n, G groups); used when
// that slice fits in shared memory (8x8 .. 32x32 levels), where the two-pass pair is launch-latency bound.
constexpr int GNC_CS = 8;         // portable maximum cluster size
constexpr int GNC_THREADS = 512;

__device__ __forceinline__ void cluster_arrive_rel() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acq() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_addr, uint32_t rank) {
  uint32_t ra;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}

template <int CS>
__global__ void __launch_bounds__(GNC_THREADS, CS == 1 ? 2 : 1)
gn_cluster_kernel(const __half* __restrict__ x1, int c1, const __half* __restrict__ x2, int c2, int HW, int groups,
                  int G, int npc, const __half* __restrict__ gamma, const __half* __restrict__ beta, float eps,
                  int silu, __half* __restrict__ out, float inv_cnt) {
  extern __shared__ uint4 gnc_smem[];
  const int C = c1 + c2;
  const int cpg = C / groups;
  const int Cs = G * cpg;               // channels of this cluster (multiple of 8)
  const int ncv = Cs / 8;
  const int nsub = groups / G;
  const uint32_t rank = CS > 1 ? cluster_rank() : 0u;
  const int cl = blockIdx.x / CS;
  const int n = cl / nsub, gs = cl % nsub;
  const int c0 = gs * Cs;
  const int p0 = (int)rank * npc;
  const int np = max(0, min(HW, p0 + npc) - p0);
  uint4* slice = gnc_smem;
  float* s_acc = reinterpret_cast<float*>(slice + (size_t)npc * ncv);   // [G] block accumulators
  float* s_part1 = s_acc + GN_MAX_GROUPS;                                 // [G] this CTA's sum        (read by peers)
  float* s_part2 = s_part1 + GN_MAX_GROUPS;                               // [G] this CTA's centred sq (read by peers)
  float* s_mean = s_part2 + GN_MAX_GROUPS;
  float* s_rstd = s_mean + GN_MAX_GROUPS;
  const int tid = threadIdx.x;
  if (tid < GN_MAX_GROUPS) s_acc[tid] = 0.f;
  const int lanes = GNC_THREADS / ncv;      // >= 1 (ncv <= 320)
  const int cv = tid % ncv;
  const int lp = tid / ncv;
  const bool active = lp < lanes;
  const int c = c0 + cv * 8;                // global channel of this thread's vector
  const int gl0 = (cv * 8) / cpg;           // first local group the vector touches
  const int split = min(8, (gl0 + 1) * cpg - cv * 8);   // elements [0, split) belong to gl0, the rest to gl0 + 1
  pdl_enter();
  __syncthreads();
  // ---- phase 1: load the slice, per-group sums
  float a0 = 0.f, a1 = 0.f;
  if (active) {
    float sm[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sm[i] = 0.f;
    for (int pix = lp; pix < np; pix += lanes) {
      const uint4 u = gn_load(x1, c1, x2, c2, (long long)n * HW + p0 + pix, c);
      slice[(size_t)pix * ncv + cv] = u;
      float f[8];
      unpack8(u, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) sm[i] += f[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < split) a0 += sm[i];
      else a1 += sm[i];
    }
    atomicAdd(&s_acc[gl0], a0);
    if (split < 8) atomicAdd(&s_acc[gl0 + 1], a1);
  }
  __syncthreads();
  if (tid < G) {
    s_part1[tid] = s_acc[tid];
    s_acc[tid] = 0.f;
  }
  if (CS > 1) {
    cluster_arrive_rel();
    cluster_wait_acq();
  } else {
    __syncthreads();
  }
  if (tid < G) {
    float t = 0.f;
    if (CS > 1) {
      const uint32_t la = static_cast<uint32_t>(__cvta_generic_to_shared(&s_part1[tid]));
#pragma unroll
      for (int r = 0; r < CS; ++r) t += ld_dsmem_f32(la, (uint32_t)r);
    } else {
      t = s_part1[tid];
    }
    s_mean[tid] = t * inv_cnt;
  }
  __syncthreads();
  // ---- phase 2: centred second moment from the cached slice
  if (active) {
    const float m0 = s_mean[gl0];
    const float m1 = split < 8 ? s_mean[gl0 + 1] : 0.f;
    float q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = 0.f;
    for (int pix = lp; pix < np; pix += lanes) {
      float f[8];
      unpack8(slice[(size_t)pix * ncv + cv], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float dlt = f[i] - (i < split ? m0 : m1);
        q[i] = fmaf(dlt, dlt, q[i]);
      }
    }
    a0 = a1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < split) a0 += q[i];
      else a1 += q[i];
    }
    atomicAdd(&s_acc[gl0], a0);
    if (split < 8) atomicAdd(&s_acc[gl0 + 1], a1);
  }
  __syncthreads();
  if (tid < G) s_part2[tid] = s_acc[tid];
  if (CS > 1) {
    cluster_arrive_rel();
    cluster_wait_acq();
  } else {
    __syncthreads();
  }
  if (tid < G) {
    float t = 0.f;
    if (CS > 1) {
      const uint32_t la = static_cast<uint32_t>(__cvta_generic_to_shared(&s_part2[tid]));
#pragma unroll
      for (int r = 0; r < CS; ++r) t += ld_dsmem_f32(la, (uint32_t)r);
    } else {
      t = s_part2[tid];
    }
    s_rstd[tid] = rsqrtf(t * inv_cnt + eps);
  }
  // peers may read this CTA's partials until they pass this point: arrive now, wait right before exit
  if (CS > 1) cluster_arrive_rel();
  __syncthreads();
  // ---- phase 3: normalise (+SiLU) straight out of shared memory
  if (active) {
    float gm[8], bt[8], mu[8], rs[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(gamma + c)), gm);
    unpack8(__ldg(reinterpret_cast<const uint4*>(beta + c)), bt);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int g = i < split ? gl0 : gl0 + 1;
      rs[i] = s_rstd[g] * gm[i];
      mu[i] = bt[i] - s_mean[g] * rs[i];          // y = x * rs + mu
    }
    for (int pix = lp; pix < np; pix += lanes) {
      float f[8];
      unpack8(slice[(size_t)pix * ncv + cv], f);
      uint4 o;
      __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float y0 = fmaf(f[2 * i], rs[2 * i], mu[2 * i]);
        float y1 = fmaf(f[2 * i + 1], rs[2 * i + 1], mu[2 * i + 1]);
        if (silu) {
          y0 = __fdividef(y0, 1.f + __expf(-y0));
          y1 = __fdividef(y1, 1.f + __expf(-y1));
        }
        oh[i] = __floats2half2_rn(y0, y1);
      }
      *reinterpret_cast<uint4*>(out + ((long long)n * HW + p0 + pix) * C + c) = o;
    }
  }
  if (CS > 1) cluster_wait_acq();
}

static inline bool gn_fused_enabled() {
  static int v = -1;
  if (v < 0) {
    // measured (r1 bench10 vs bench9): the activations are L2-resident between the two passes, so the
    // cooperative single-pass kernel's per-image barrier costs more than the second read saves
    // (1.31 vs 1.18 ms per UNet evaluation) -> opt-in only.
    const char* e = getenv("PFD_GN_FUSED");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

static inline int grid_for(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace pfd

using namespace pfd;

extern "C" PFD_API int pfd_groupnorm_f16(const void* x1, int32_t c1, const void* x2, int32_t c2, int32_t NB,
                                 int64_t HW, int32_t groups, const void* gamma, const void* beta,
                                 float eps, int32_t silu, void* out, float* ws, int32_t zero_ws, void* stream) {
  const int C = c1 + (x2 ? c2 : 0);
  if (!x2) c2 = 0;
  if (groups <= 0 || groups > GN_MAX_GROUPS || C % groups) return set_error("pfd_groupnorm_f16: C=%d groups=%d", C, groups);
  if (c1 % 8 || c2 % 8) return set_error("pfd_groupnorm_f16: channel counts must be multiples of 8 (%d,%d)", c1, c2);
  if (!ws) return set_error("pfd_groupnorm_f16: workspace required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* dws = reinterpret_cast<double*>(ws);
  const int vecs = C / 8;
  // block = largest multiple of vecs that fits 256 threads (or 320 for C = 2560); wide rows fall back to 256
  int threads = vecs <= 320 ? (vecs <= 256 ? (256 / vecs) * vecs : vecs) : 256;
  if (threads < 64) threads = vecs * ((64 + vecs - 1) / vecs);
  const double inv_cnt = 1.0 / ((double)HW * (C / groups));
  // ---- single-pass cluster path: one 8-CTA cluster per (image, G groups), slice cached in shared memory
  {
    static int cl_mode = -1;       // -1 unknown, 0 off, 1 on
    if (cl_mode < 0) {
      // opt-in: correct, but measured ~2x slower than the two-pass kernels on every UNet shape (r1 gn_perf.log:
      // 1.91 vs 1.04 ms per evaluation) - cluster launches of 8 x 512 threads cost more than the second read saves
      const char* e = getenv("PFD_GN_CLUSTER");
      cl_mode = (e && e[0] == '1') ? 1 : 0;
    }
    const int cpg = C / groups;
    static int solo_mode = -1;
    if (solo_mode < 0) {
      // opt-in: measured slower than the two-pass pair as well (r1 gn_perf2.log: 1.24 vs 1.04 ms per evaluation;
      // 64-256 CTAs walking a 40-100 KB slice three times lose to ~450 CTAs streaming at full bandwidth twice)
      const char* e = getenv("PFD_GN_SOLO");
      solo_mode = (e && e[0] == '1') ? 1 : 0;
    }
    if (solo_mode == 1 && cpg >= 8 && HW <= (1 << 20)) {
      // one CTA per (image, G groups): largest slice <= 100 KB (two CTAs per SM) that still gives >= 64 CTAs
      const size_t slice_max = 100 * 1024;
      int Gsel = 0;
      for (int G = groups; G >= 1; --G) {
        if (groups % G) continue;
        const int Cs = G * cpg;
        if (Cs % 8 || Cs / 8 > GNC_THREADS) continue;
        if ((size_t)HW * (Cs / 8) * 16 > slice_max) continue;
        Gsel = G;
        if ((long long)NB * (groups / G) >= 64) break;
      }
      if (Gsel > 0) {
        const int Cs = Gsel * cpg;
        const size_t smem = (size_t)HW * (Cs / 8) * 16 + 5 * GN_MAX_GROUPS * sizeof(float);
        static bool attr_set1 = false;
        if (!attr_set1) {
          cudaFuncSetAttribute(gn_cluster_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(slice_max + 5 * GN_MAX_GROUPS * sizeof(float)));
          attr_set1 = true;
        }
        launch_k(gn_cluster_kernel<1>, dim3((unsigned)(NB * (groups / Gsel))), dim3(GNC_THREADS), smem, st,
                 static_cast<const __half*>(x1), (int)c1, static_cast<const __half*>(x2), (int)c2, (int)HW, (int)groups,
                 Gsel, (int)HW, static_cast<const __half*>(gamma), static_cast<const __half*>(beta), eps, (int)silu,
                 static_cast<__half*>(out), (float)inv_cnt);
        return check_launch("gn_solo");
      }
    }
    if (cl_mode == 1 && cpg >= 8 && HW >= GNC_CS && HW <= (1 << 24)) {
      const int npc = (int)((HW + GNC_CS - 1) / GNC_CS);
      const size_t slice_max = 200 * 1024;
      int Gsel = 0;
      for (int G = groups; G >= 1; --G) {
        if (groups % G) continue;
        const int Cs = G * cpg;
        if (Cs % 8 || Cs / 8 > GNC_THREADS) continue;
        if ((size_t)npc * (Cs / 8) * 16 > slice_max) continue;
        Gsel = G;                                                   // valid; keep shrinking until the machine is filled
        if ((long long)NB * (groups / G) * GNC_CS >= num_sms()) break;
      }
      if (Gsel > 0) {
        const int Cs = Gsel * cpg;
        const size_t smem = (size_t)npc * (Cs / 8) * 16 + 5 * GN_MAX_GROUPS * sizeof(float);
        static bool attr_set = false;
        if (!attr_set) {
          cudaFuncSetAttribute(gn_cluster_kernel<GNC_CS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(slice_max + 5 * GN_MAX_GROUPS * sizeof(float)));
          attr_set = true;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)((long long)NB * (groups / Gsel) * GNC_CS));
        cfg.blockDim = dim3(GNC_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = GNC_CS;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = use_pdl() ? 2 : 1;
        cudaError_t le = cudaLaunchKernelEx(&cfg, gn_cluster_kernel<GNC_CS>, static_cast<const __half*>(x1), (int)c1,
                                            static_cast<const __half*>(x2), (int)c2, (int)HW, (int)groups, Gsel, npc,
                                            static_cast<const __half*>(gamma), static_cast<const __half*>(beta), eps,
                                            (int)silu, static_cast<__half*>(out), (float)inv_cnt);
        if (le == cudaSuccess) return check_launch("gn_cluster");
        (void)cudaGetLastError();     // cluster launch rejected on this device/config: use the two-pass kernels
        cl_mode = 0;
      }
    }
  }
  if (zero_ws) cudaMemsetAsync(dws, 0, sizeof(double) * 2 * NB * groups, st);
  // ---- single-pass path (chunk cached in shared memory, per-image arrival barrier) when the whole grid can be
  //      co-resident; counters live right behind the fp64 sums in the (pre-zeroed) scratch slot
  if (!zero_ws && vecs <= 320 && gn_fused_enabled()) {
    const int lanes = threads / vecs;
    static int max_smem_set = 0;
    const size_t smem_cap = 96 * 1024;
    long long ppc_f = (long long)(smem_cap / ((size_t)C * 2));
    ppc_f = (ppc_f / lanes) * lanes;
    if (ppc_f > HW) ppc_f = ((HW + lanes - 1) / lanes) * lanes;
    if (ppc_f >= 4 * lanes) {
      const size_t smem = (size_t)ppc_f * C * 2;
      if (!max_smem_set) {
        cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap);
        max_smem_set = 1;
      }
      int per_sm = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn_fused_kernel, threads, smem);
      // shrink the chunk (more CTAs) while everything still fits, to fill the machine
      long long chunks_f = (HW + ppc_f - 1) / ppc_f;
      const long long cap = (long long)per_sm * num_sms();
      if (per_sm > 0 && chunks_f * NB <= cap) {
        unsigned int* counter = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(ws) + (size_t)NB * groups * 16);
        dim3 gridf((unsigned)chunks_f, (unsigned)NB);
        launch_k(gn_fused_kernel, gridf, dim3(threads), smem, st, static_cast<const __half*>(x1), c1,
                 static_cast<const __half*>(x2), c2, (long long)HW, groups, static_cast<const __half*>(gamma),
                 static_cast<const __half*>(beta), eps, silu, dws, counter, static_cast<__half*>(out), ppc_f, inv_cnt);
        return check_launch("gn_fused");
      }
    }
  }
  // ~3 CTAs per SM, but at least 16 pixels per pixel-lane so the per-CTA setup is amortised
  long long chunks = (3LL * num_sms() + NB - 1) / NB;
  long long ppc = (HW + chunks - 1) / chunks;
  const long long min_ppc = 16LL * (threads / vecs > 0 ? threads / vecs : 1);
  if (ppc < min_ppc) ppc = min_ppc;
  chunks = (HW + ppc - 1) / ppc;
  dim3 grid((unsigned)chunks, (unsigned)NB);
  launch_k(gn_stats_kernel, dim3(grid), dim3(threads), (size_t)(0), st, static_cast<const __half*>(x1), c1, static_cast<const __half*>(x2), c2,
                                            HW, groups, ppc, dws);
  if (int rc = check_launch("gn_stats")) return rc;
  launch_k(gn_apply_kernel, dim3(grid), dim3(threads), (size_t)(0), st, static_cast<const __half*>(x1), c1, static_cast<const __half*>(x2), c2, HW,
                                            groups, static_cast<const __half*>(gamma), static_cast<const __half*>(beta),
                                            eps, silu, dws, static_cast<__half*>(out), ppc,
                                            inv_cnt);
  return check_launch("gn_apply");
}

extern "C" PFD_API int pfd_layernorm_f16(const void* x, const void* res, int64_t rows, int32_t C,
                                 const void* gamma, const void* beta, float eps, void* out,
                                 void* stream) {
  if (C % 8 || C > 8 * 32 * 16) return set_error("pfd_layernorm_f16: C=%d unsupported", C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int wpb = 8;
  const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
  const int vecs = C / 8;
  const __half* xp = static_cast<const __half*>(x);
  const __half* rp = static_cast<const __half*>(res);
  const __half* gp = static_cast<const __half*>(gamma);
  const __half* bp = static_cast<const __half*>(beta);
  __half* op = static_cast<__half*>(out);
  if (vecs <= 32) launch_k(layernorm_kernel<1>, dim3(grid), dim3(256), (size_t)(0), st, xp, rp, rows, C, gp, bp, eps, op);
  else if (vecs <= 64) launch_k(layernorm_kernel<2>, dim3(grid), dim3(256), (size_t)(0), st, xp, rp, rows, C, gp, bp, eps, op);
  else if (vecs <= 128) launch_k(layernorm_kernel<4>, dim3(grid), dim3(256), (size_t)(0), st, xp, rp, rows, C, gp, bp, eps, op);
  else if (vecs <= 256) launch_k(layernorm_kernel<8>, dim3(grid), dim3(256), (size_t)(0), st, xp, rp, rows, C, gp, bp, eps, op);
  else launch_k(layernorm_kernel<16>, dim3(grid), dim3(256), (size_t)(0), st, xp, rp, rows, C, gp, bp, eps, op);
  return check_launch("layernorm");
}

extern "C" PFD_API int pfd_softmax_f16(void* s, int64_t batch, int32_t rows, int32_t cols, int64_t ld,
                               float scale, const void* bias, int32_t nheads, const void* mask,
                               int32_t nwin, void* stream) {
  if (cols <= 0 || cols > 12288) return set_error("pfd_softmax_f16: cols=%d unsupported", cols);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long nrows = batch * rows;
  if (nheads <= 0) nheads = 1;
  if (nwin <= 0) nwin = 1;
  int threads = cols >= 1024 ? 256 : (cols >= 256 ? 128 : 64);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(softmax_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 12288 * 4);
    attr = true;
  }
  launch_k(softmax_rows_kernel, dim3((unsigned)nrows), dim3(threads), (size_t)(cols * sizeof(float)), st, 
      static_cast<__half*>(s), batch, rows, cols, ld, scale, static_cast<const __half*>(bias), nheads,
      static_cast<const __half*>(mask), nwin);
  return check_launch("softmax");
}

extern "C" PFD_API int pfd_timestep_embedding_f16(const int64_t* t, int32_t n, int32_t dim, float max_period,
                                          void* out, void* stream) {
  const int total = n * (dim / 2);
  launch_k(timestep_embedding_kernel, dim3((total + 127) / 128), dim3(128), (size_t)(0), static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const long long*>(t), n, dim, max_period, static_cast<__half*>(out));
  return check_launch("timestep_embedding");
}

extern "C" PFD_API int pfd_upsample2x_f16(const void* x, int32_t NB, int32_t H, int32_t W, int32_t C, void* out,
                                  void* stream) {
  if (C % 8) return set_error("pfd_upsample2x_f16: C=%d", C);
  const long long total = (long long)NB * 4 * H * W * (C / 8);
  launch_k(upsample2x_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), 
      static_cast<const uint4*>(x), NB, H, W, C / 8, static_cast<uint4*>(out));
  return check_launch("upsample2x");
}

extern "C" PFD_API int pfd_nchw_to_nhwc_f16(const void* x, int32_t src_is_f32, int32_t NB, int32_t C, int32_t H,
                                    int32_t W, int32_t Cpad, void* out, void* stream) {
  const long long total = (long long)NB * H * W * Cpad;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (src_is_f32)
    launch_k(nchw_to_nhwc_kernel<float>, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), st, static_cast<const float*>(x), NB, C, H, W, Cpad, static_cast<__half*>(out));
  else
    launch_k(nchw_to_nhwc_kernel<__half>, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), st, static_cast<const __half*>(x), NB, C, H, W, Cpad, static_cast<__half*>(out));
  return check_launch("nchw_to_nhwc");
}

extern "C" PFD_API int pfd_nhwc_to_nchw_f16(const void* x, int32_t NB, int32_t C, int32_t H, int32_t W, int32_t Cpad,
                                    float mul, float add, float lo, float hi, void* out, void* stream) {
  const long long total = (long long)NB * C * H * W;
  launch_k(nhwc_to_nchw_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), 
      static_cast<const __half*>(x), NB, C, H, W, Cpad, mul, add, lo, hi, static_cast<__half*>(out));
  return check_launch("nhwc_to_nchw");
}

extern "C" PFD_API int pfd_im2col3x3_f16(const void* x, int32_t NB, int32_t H, int32_t W, int32_t C, int32_t stride,
                                 int32_t Kpad, void* out, void* stream) {
  if (Kpad < 9 * C || Kpad % 8) return set_error("pfd_im2col3x3_f16: Kpad=%d for C=%d", Kpad, C);
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  const long long total = (long long)NB * Ho * Wo * Kpad;
  launch_k(im2col3x3_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), 
      static_cast<const __half*>(x), NB, H, W, C, stride, Ho, Wo, Kpad, static_cast<__half*>(out));
  return check_launch("im2col3x3");
}

extern "C" PFD_API int pfd_axpby_f16(const void* a, float sa, const void* b, float sb, int64_t n, void* out,
                             void* stream) {
  launch_k(axpby_kernel, dim3(grid_for(n, 256)), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), 
      static_cast<const __half*>(a), sa, static_cast<const __half*>(b), sb, n, static_cast<__half*>(out));
  return check_launch("axpby");
}

extern "C" PFD_API int pfd_add_rowvec_f16(const void* a, const void* row, int64_t rows, int32_t C, void* out,
                                  void* stream) {
  if (C % 8) return set_error("pfd_add_rowvec_f16: C=%d", C);
  launch_k(add_rowvec_kernel, dim3(grid_for(rows * (C / 8), 256)), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), 
      static_cast<const uint4*>(a), static_cast<const uint4*>(row), rows, C / 8, static_cast<uint4*>(out));
  return check_launch("add_rowvec");
}

extern "C" PFD_API int pfd_ddim_step_f16(const void* eps, const void* x, int64_t half_n, float guidance,
                                 const float* coef, const int32_t* step, void* x_prev, void* pred_x0,
                                 void* stream) {
  launch_k(ddim_step_kernel, dim3(grid_for(half_n, 256)), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), 
      static_cast<const __half*>(eps), static_cast<const __half*>(x), half_n, guidance, coef, step,
      static_cast<__half*>(x_prev), static_cast<__half*>(pred_x0));
  return check_launch("ddim_step");
}

extern "C" PFD_API int pfd_window_gather_f16(const void* x, int32_t B, int32_t H, int32_t W, int32_t C, int32_t ws,
                                     int32_t shift, void* out, void* stream) {
  if (C % 8) return set_error("pfd_window_gather_f16: C=%d", C);
  const int Hp = (H + ws - 1) / ws * ws, Wp = (W + ws - 1) / ws * ws;
  const long long total = (long long)B * Hp * Wp * (C / 8);
  launch_k(window_gather_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), 
      static_cast<const uint4*>(x), B, H, W, C / 8, ws, shift, Hp, Wp, static_cast<uint4*>(out));
  return check_launch("window_gather");
}

extern "C" PFD_API int pfd_window_scatter_f16(const void* win, int32_t B, int32_t H, int32_t W, int32_t C,
                                      int32_t ws, int32_t shift, const void* residual, void* out,
                                      void* stream) {
  if (C % 8) return set_error("pfd_window_scatter_f16: C=%d", C);
  const int Hp = (H + ws - 1) / ws * ws, Wp = (W + ws - 1) / ws * ws;
  const long long total = (long long)B * H * W * (C / 8);
  launch_k(window_scatter_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), 
      static_cast<const uint4*>(win), B, H, W, C / 8, ws, shift, Hp, Wp,
      static_cast<const uint4*>(residual), static_cast<uint4*>(out));
  return check_launch("window_scatter");
}

extern "C" PFD_API int pfd_patch_merge_gather_f16(const void* x, int32_t B, int32_t H, int32_t W, int32_t C,
                                          void* out, void* stream) {
  if (C % 8) return set_error("pfd_patch_merge_gather_f16: C=%d", C);
  const long long total = (long long)B * ((H + 1) / 2) * ((W + 1) / 2) * 4 * (C / 8);
  launch_k(patch_merge_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), 
      static_cast<const uint4*>(x), B, H, W, C / 8, static_cast<uint4*>(out));
  return check_launch("patch_merge");
}

extern "C" PFD_API int pfd_patchify_f16(const void* x, int32_t src_is_f32, int32_t B, int32_t C, int32_t H,
                                        int32_t W, int32_t P, int32_t Kpad, void* out, void* stream) {
  if (Kpad < C * P * P || Kpad % 8) return set_error("pfd_patchify_f16: Kpad=%d for C=%d P=%d", Kpad, C, P);
  const long long total = (long long)B * ((H + P - 1) / P) * ((W + P - 1) / P) * Kpad;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (src_is_f32)
    launch_k(patchify_kernel<float>, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), st, static_cast<const float*>(x), B, C, H, W, P, Kpad, static_cast<__half*>(out));
  else
    launch_k(patchify_kernel<__half>, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), st, static_cast<const __half*>(x), B, C, H, W, P, Kpad, static_cast<__half*>(out));
  return check_launch("patchify");
}
