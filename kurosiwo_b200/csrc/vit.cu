// Bandwidth-bound / small-matrix passes of the ViT encoder and the FloodViT segmentation head.
//
// Reference: models/vision_transformer.py (ViT.forward :139-156, Transformer.forward :84-89, Attention.forward :53-66,
// FeedForward :19-32) and models/model_utilities.py (FinetunerSegmentation.forward :80-94).
// The token matrix is [B*Tp, C] row-major (Tp = tokens per image padded to a multiple of 16, T valid), storage dtype
// fp32 (parity mode) or bf16 (perf mode); the GEMMs run on the conv engine as 1x1 convolutions over the same buffers.
//   ks_patchify_ln / _bwd   Rearrange 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)' + LayerNorm(patch_dim)   (:122-123)
//   ks_layernorm_fwd / _bwd nn.LayerNorm over the last dim (+ optional copy of x = the residual stream's next buffer)
//   ks_vit_assemble / _bwd  cat(cls_token, x) + pos_embedding                                          (:142-146)
//   ks_attention_fwd / _bwd softmax(q k^T * scale) v per (image, head), probabilities kept for the backward (:59-65)
//   ks_gelu_fwd / _bwd      exact-erf GELU                                                               (:25)
//   ks_bilinear_up_fwd/_bwd nn.Upsample(size, mode='bilinear') of the K-class token map (model_utilities.py:89-91; the
//                           1x1 head commutes with the interpolation, so K channels are upsampled instead of `dim`)
#include "common.cuh"
#include <mma.h>

namespace ks {

template <typename T> __device__ __forceinline__ void ldv8(const T *p, float (&f)[8]) { ld8(p, f); }
template <typename T> __device__ __forceinline__ void stv8(T *p, const float (&f)[8]) { st8(p, f); }

// ---------------------------------------------------------------------------------------------------------
// LayerNorm forward: one warp per row, the row lives in registers (C <= 2048, C % 8 == 0).
// ---------------------------------------------------------------------------------------------------------
constexpr int LN_KMAX = 8;   // 8 chunks x 32 lanes x 8 elements = 2048 columns

template <typename T>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(long long rows, int C, const T *__restrict__ x, long long ldx, const float *__restrict__ gamma,
                     const float *__restrict__ beta, float eps, T *__restrict__ y, long long ldy, float *__restrict__ mean,
                     float *__restrict__ rstd, T *__restrict__ copy_out, long long ldc) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (long long r = (long long)blockIdx.x * wpb + wib; r < rows; r += (long long)gridDim.x * wpb) {
    const T *xp = x + r * ldx;
    float v[LN_KMAX][8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LN_KMAX; ++k) {
      const int c = (k * 32 + lane) * 8;
      if (c < C) {
        ldv8(xp + c, v[k]);
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[k][i];
      }
    }
    const float mu = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < LN_KMAX; ++k) {
      const int c = (k * 32 + lane) * 8;
      if (c < C) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[k][i] - mu; q += d * d; }
      }
    }
    const float rs = rsqrtf(warp_sum(q) / (float)C + eps);
    if (lane == 0) { if (mean) mean[r] = mu; if (rstd) rstd[r] = rs; }
#pragma unroll
    for (int k = 0; k < LN_KMAX; ++k) {
      const int c = (k * 32 + lane) * 8;
      if (c < C) {
        if (copy_out) stv8(copy_out + r * ldc + c, v[k]);
        float g[8], b[8], o[8];
        ld8(gamma + c, g); ld8(beta + c, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (v[k][i] - mu) * rs * g[i] + b[i];
        stv8(y + r * ldy + c, o);
      }
    }
  }
}

// Narrow rows (C <= 256, e.g. the 64-channel ChangeFormer stage with 200 k rows): L = pow2 >= C/8 lanes per row, 32/L rows per warp -
// with a whole warp per row 24 of 32 lanes idled at C = 64.
template <typename T>
__global__ void __launch_bounds__(256)
layernorm_fwd_narrow_kernel(long long rows, int C, int L, const T *__restrict__ x, long long ldx, const float *__restrict__ gamma,
                            const float *__restrict__ beta, float eps, T *__restrict__ y, long long ldy, float *__restrict__ mean,
                            float *__restrict__ rstd, T *__restrict__ copy_out, long long ldc) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int rpw = 32 / L, sub = lane / L, c = (lane % L) * 8;
  const bool cl = c < C;
  float g[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { g[i] = 0.f; b[i] = 0.f; }
  if (cl) { ld8(gamma + c, g); ld8(beta + c, b); }
  const float invC = 1.f / (float)C;
  for (long long r0 = ((long long)blockIdx.x * wpb + wib) * rpw; r0 < rows; r0 += (long long)gridDim.x * wpb * rpw) {
    const long long r = r0 + sub;
    const bool live = cl && r < rows;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (live) ldv8(x + r * ldx + c, v);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    for (int o = L >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu = s * invC;
    float q = 0.f;
    if (live) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = v[i] - mu; q += d * d; }
    }
    for (int o = L >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rs = rsqrtf(q * invC + eps);
    if (live) {
      if (c == 0) { if (mean) mean[r] = mu; if (rstd) rstd[r] = rs; }
      if (copy_out) stv8(copy_out + r * ldc + c, v);
      float o8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o8[i] = (v[i] - mu) * rs * g[i] + b[i];
      stv8(y + r * ldy + c, o8);
    }
  }
}

// LayerNorm backward (C <= 1024), two kernels:
//   rows: dx (+)= rstd*(g - mean(g) - xhat*mean(g*xhat)), g = dy*gamma.  L = min(32, pow2 >= C/8) lanes own one row (32/L rows per warp,
//         so narrow rows - the 64-channel ChangeFormer stage - keep every lane busy), K = ceil(C / (8 L)) vectors per lane.
//   cols: dgamma += sum_rows dy*xhat, dbeta += sum_rows dy, thread = 8 channels x a row lane (coalesced over C, no cross-lane traffic
//         per row); dy and x are re-read, for the ViT / ChangeFormer sizes out of L2.
// (One fused kernel with register column accumulators needed 218 registers: 8 warps per SM and 1.1-1.5 TB/s.)
constexpr int LNB_KMAX = 4;

template <typename T, int K>
__global__ void __launch_bounds__(256, 2)
layernorm_bwd_rows_kernel(long long rows, int C, int L, const T *__restrict__ dy, long long lddy, const T *__restrict__ x, long long ldx,
                          const float *__restrict__ mean, const float *__restrict__ rstd, const float *__restrict__ gamma,
                          T *__restrict__ dx, long long lddx, int acc_dx) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int rpw = 32 / L, sub = lane / L, sl = lane % L;
  float gm[K][8];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int c = (k * L + sl) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) gm[k][i] = 0.f;
    if (c < C) ld8(gamma + c, gm[k]);
  }
  const float invC = 1.f / (float)C;
  if constexpr (sizeof(T) == 2) {
    // bf16 storage: two-deep software pipeline.  The raw 16-byte vectors of the NEXT row are in flight while this one is reduced
    // (one row per warp iteration left every lane with 2 K loads outstanding: the pass ran at ~1.6 TB/s out of L2); the row is
    // converted twice (sums, then output) so that only the raw registers live across the shuffle reduction.
    const long long stride = (long long)gridDim.x * wpb * rpw;
    uint4 nd[K], nx[K];
    float nmu = 0.f, nrs = 0.f;
    auto fetch = [&](long long r0_) {
      const long long r = r0_ + sub;
      const bool live = r < rows;
      nmu = live ? mean[r] : 0.f; nrs = live ? rstd[r] : 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int c = (k * L + sl) * 8;
        nd[k] = make_uint4(0u, 0u, 0u, 0u); nx[k] = make_uint4(0u, 0u, 0u, 0u);
        if (live && c < C) {
          nd[k] = *reinterpret_cast<const uint4 *>(dy + r * lddy + c);
          nx[k] = *reinterpret_cast<const uint4 *>(x + r * ldx + c);
        }
      }
    };
    auto cvt = [](const uint4 &u, float (&f)[8]) {
      const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    };
    long long r0 = ((long long)blockIdx.x * wpb + wib) * rpw;
    if (r0 < rows) fetch(r0);
    for (; r0 < rows; r0 += stride) {
      uint4 cd[K], cx[K];
#pragma unroll
      for (int k = 0; k < K; ++k) { cd[k] = nd[k]; cx[k] = nx[k]; }
      const float mu = nmu, rs = nrs;
      const long long r = r0 + sub;
      const bool live = r < rows;
      if (r0 + stride < rows) fetch(r0 + stride);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float d[8], xv[8];
        cvt(cd[k], d); cvt(cx[k], xv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = (xv[i] - mu) * rs, g = d[i] * gm[k][i];
          s1 += g; s2 = fmaf(g, xh, s2);
        }
      }
      for (int o = L >> 1; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
      const float m1 = s1 * invC, m2 = s2 * invC;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int c = (k * L + sl) * 8;
        if (live && c < C) {
          float d[8], xv[8], o[8];
          cvt(cd[k], d); cvt(cx[k], xv);
          if (acc_dx) ldv8(dx + r * lddx + c, o);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float xh = (xv[i] - mu) * rs, g = d[i] * gm[k][i];
            const float t = rs * (g - m1 - xh * m2);
            o[i] = acc_dx ? o[i] + t : t;
          }
          stv8(dx + r * lddx + c, o);
        }
      }
    }
    return;
  }
  for (long long r0 = ((long long)blockIdx.x * wpb + wib) * rpw; r0 < rows; r0 += (long long)gridDim.x * wpb * rpw) {
    const long long r = r0 + sub;
    const bool live = r < rows;
    const float mu = live ? mean[r] : 0.f, rs = live ? rstd[r] : 0.f;
    float g[K][8], xh[K][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int c = (k * L + sl) * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) { g[k][i] = 0.f; xh[k][i] = 0.f; }
      if (live && c < C) {
        float d[8], xv[8];
        ldv8(dy + r * lddy + c, d); ldv8(x + r * ldx + c, xv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xh[k][i] = (xv[i] - mu) * rs;
          g[k][i] = d[i] * gm[k][i];
          s1 += g[k][i]; s2 += g[k][i] * xh[k][i];
        }
      }
    }
    for (int o = L >> 1; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    const float m1 = s1 * invC, m2 = s2 * invC;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int c = (k * L + sl) * 8;
      if (live && c < C) {
        float o[8];
        if (acc_dx) ldv8(dx + r * lddx + c, o);
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float t = rs * (g[k][i] - m1 - xh[k][i] * m2); o[i] = acc_dx ? o[i] + t : t; }
        stv8(dx + r * lddx + c, o);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
layernorm_bwd_cols_kernel(long long rows, int C, const T *__restrict__ dy, long long lddy, const T *__restrict__ x, long long ldx,
                          const float *__restrict__ mean, const float *__restrict__ rstd, float *__restrict__ dgamma, float *__restrict__ dbeta) {
  extern __shared__ float sred[];     // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  const int CV = C / 8, cvb = min(CV, 256), rl = 256 / cvb;           // channel vectors per pass, row lanes
  const int tx = threadIdx.x % cvb, ty = threadIdx.x / cvb;
  if (ty < rl) {
    for (int cv = tx; cv < CV; cv += cvb) {
      const int c = cv * 8;
      float ag[8], ab[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { ag[i] = 0.f; ab[i] = 0.f; }
      const long long stride = (long long)gridDim.x * rl;
      for (long long r0 = (long long)blockIdx.x * rl + ty; r0 < rows; r0 += 4 * stride) {
        float d[4][8], xv[4][8], mu[4], rs[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long long r = r0 + u * stride;
          if (r < rows) { ldv8(dy + r * lddy + c, d[u]); ldv8(x + r * ldx + c, xv[u]); mu[u] = mean[r]; rs[u] = rstd[r]; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (r0 + u * stride < rows) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { ag[i] = fmaf(d[u][i], (xv[u][i] - mu[u]) * rs[u], ag[i]); ab[i] += d[u][i]; }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) { atomicAdd(&sred[c + i], ag[i]); atomicAdd(&sred[C + c + i], ab[i]); }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + i, sred[i]);
    if (dbeta) atomicAdd(dbeta + i, sred[C + i]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Patchify + LayerNorm(patch_dim).  img NCHW fp32 [B][Cc][Hi][Wi]; patch vector order (p1 p2 c), channel fastest
// (vision_transformer.py:122).  One warp per patch; lane l owns pixel row p1 = l/2 and the 8 pixels p2 = 8*(l%2)..+7,
// all Cc channels: 8*Cc CONTIGUOUS output elements.  Patch p of image b goes to row b*Tp + 1 + p (row 0 is the cls slot).
// ---------------------------------------------------------------------------------------------------------
constexpr int PATCH = 16, PCMAX = 8;

template <typename T>
__global__ void __launch_bounds__(256)
patchify_ln_kernel(int B, int Cc, int Hi, int Wi, int Tp, const float *__restrict__ img, const float *__restrict__ gamma,
                   const float *__restrict__ beta, float eps, T *__restrict__ out, float *__restrict__ mean, float *__restrict__ rstd) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int gw = Wi / PATCH, gh = Hi / PATCH, np = gw * gh, PD = PATCH * PATCH * Cc;
  const int p1 = lane >> 1, p20 = (lane & 1) * 8;
  for (long long pi = (long long)blockIdx.x * wpb + wib; pi < (long long)B * np; pi += (long long)gridDim.x * wpb) {
    const int b = (int)(pi / np), pp = (int)(pi % np), gy = pp / gw, gx = pp % gw;
    float v[PCMAX][8];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < PCMAX; ++c)
      if (c < Cc) {
        const float *ip = img + (((long long)b * Cc + c) * Hi + (gy * PATCH + p1)) * Wi + gx * PATCH + p20;
        const float4 a = *reinterpret_cast<const float4 *>(ip), bq = *reinterpret_cast<const float4 *>(ip + 4);
        v[c][0] = a.x; v[c][1] = a.y; v[c][2] = a.z; v[c][3] = a.w; v[c][4] = bq.x; v[c][5] = bq.y; v[c][6] = bq.z; v[c][7] = bq.w;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[c][i];
      }
    const float mu = warp_sum(s) / (float)PD;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < PCMAX; ++c)
      if (c < Cc) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[c][i] - mu; q += d * d; }
      }
    const float rs = rsqrtf(warp_sum(q) / (float)PD + eps);
    const long long row = (long long)b * Tp + 1 + pp;
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
    T *op = out + row * PD + (p1 * PATCH + p20) * Cc;
    const float *gp = gamma + (p1 * PATCH + p20) * Cc, *bp = beta + (p1 * PATCH + p20) * Cc;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < PCMAX; ++c)
        if (c < Cc) Cvt<T>::st(op + i * Cc + c, (v[c][i] - mu) * rs * __ldg(gp + i * Cc + c) + __ldg(bp + i * Cc + c));
  }
}

// dgamma[e] = sum_patches dy[row][e]*xhat[row][e], dbeta[e] = sum dy[row][e]  (the image needs no gradient)
template <typename T>
__global__ void __launch_bounds__(256)
patchify_ln_bwd_kernel(int B, int Cc, int Hi, int Wi, int Tp, const float *__restrict__ img, const float *__restrict__ mean,
                       const float *__restrict__ rstd, const T *__restrict__ dy, float *__restrict__ dgamma, float *__restrict__ dbeta) {
  extern __shared__ float sred[];    // [2][PD]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int gw = Wi / PATCH, gh = Hi / PATCH, np = gw * gh, PD = PATCH * PATCH * Cc;
  for (int i = threadIdx.x; i < 2 * PD; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  const int p1 = lane >> 1, p20 = (lane & 1) * 8;
  float ag[PCMAX][8], ab[PCMAX][8];
#pragma unroll
  for (int c = 0; c < PCMAX; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) { ag[c][i] = 0.f; ab[c][i] = 0.f; }
  for (long long pi = (long long)blockIdx.x * wpb + wib; pi < (long long)B * np; pi += (long long)gridDim.x * wpb) {
    const int b = (int)(pi / np), pp = (int)(pi % np), gy = pp / gw, gx = pp % gw;
    const long long row = (long long)b * Tp + 1 + pp;
    const float mu = mean[row], rs = rstd[row];
    const T *dp = dy + row * PD + (p1 * PATCH + p20) * Cc;
#pragma unroll
    for (int c = 0; c < PCMAX; ++c)
      if (c < Cc) {
        const float *ip = img + (((long long)b * Cc + c) * Hi + (gy * PATCH + p1)) * Wi + gx * PATCH + p20;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = Cvt<T>::ld(dp + i * Cc + c);
          ag[c][i] += d * (ip[i] - mu) * rs; ab[c][i] += d;
        }
      }
  }
#pragma unroll
  for (int c = 0; c < PCMAX; ++c)
    if (c < Cc) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int e = (p1 * PATCH + p20 + i) * Cc + c;
        atomicAdd(&sred[e], ag[c][i]); atomicAdd(&sred[PD + e], ab[c][i]);
      }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < PD; i += blockDim.x) { atomicAdd(dgamma + i, sred[i]); atomicAdd(dbeta + i, sred[PD + i]); }
}

// ---------------------------------------------------------------------------------------------------------
// x0[b,0] = cls + pos[0]; x0[b,t] = e[b,t] + pos[t] (1 <= t < T); padding rows (t >= T) = 0.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
vit_assemble_kernel(int B, int Tt, int Tp, int D, const T *__restrict__ e, const float *__restrict__ cls, const float *__restrict__ pos,
                    T *__restrict__ x0) {
  const int DV = D / 8;
  const long long total = (long long)B * Tp * DV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % DV) * 8; const long long row = i / DV; const int t = (int)(row % Tp);
    float o[8];
    if (t >= Tt) {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = 0.f;
    } else {
      float a[8], pz[8];
      if (t == 0) ld8(cls + c, a); else ldv8(e + row * D + c, a);
      ld8(pos + (long long)t * D + c, pz);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = a[k] + pz[k];
    }
    stv8(x0 + row * D + c, o);
  }
}

// dpos[t] = sum_b dx0[b,t]; dcls = sum_b dx0[b,0]; de[b,t] = dx0[b,t] for 1 <= t < T, 0 elsewhere.   grid: (Tp, D/8 chunks)
template <typename T>
__global__ void __launch_bounds__(128)
vit_assemble_bwd_kernel(int B, int Tt, int Tp, int D, const T *__restrict__ dx0, T *__restrict__ de, float *__restrict__ dcls,
                        float *__restrict__ dpos) {
  const int t = blockIdx.x;
  for (int c = (blockIdx.y * blockDim.x + threadIdx.x) * 8; c < D; c += gridDim.y * blockDim.x * 8) {
    float s[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = 0.f;
    for (int b = 0; b < B; ++b) {
      const long long row = (long long)b * Tp + t;
      float g[8], z[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) z[k] = 0.f;
      if (t < Tt) {
        ldv8(dx0 + row * D + c, g);
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] += g[k];
      }
      if (t >= 1 && t < Tt) stv8(de + row * D + c, g); else stv8(de + row * D + c, z);
    }
    if (t < Tt) {
      st8(dpos + (long long)t * D + c, s);
      if (t == 0) st8(dcls + c, s);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Attention, exact fp32 arithmetic on CUDA cores (parity-mode engine and first bf16 path).  One CTA per (b, head, row block).
// qkv row layout: [q (heads*dh) | k (heads*dh) | v (heads*dh)], head-major inside each third ('b n (h d) -> b h n d').
// K and V of the head sit in shared memory as fp32 with row pitch dh+1 (conflict-free column walks).
// ---------------------------------------------------------------------------------------------------------
constexpr int ATT_DH = 64, ATT_PITCH = ATT_DH + 1, ATT_WARPS = 8, ATT_JMAX = 8;   // Tp <= 256

template <typename T>
__device__ __forceinline__ void att_load_tile(float *dst, const T *src, long long ld, int rows_valid, int rows_total) {
  for (int i = threadIdx.x; i < rows_total * (ATT_DH / 8); i += blockDim.x) {
    const int r = i / (ATT_DH / 8), c = (i % (ATT_DH / 8)) * 8;
    float f[8];
    if (r < rows_valid) ldv8(src + (long long)r * ld + c, f);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) dst[r * ATT_PITCH + c + k] = f[k];
  }
}

template <typename T>
__global__ void __launch_bounds__(ATT_WARPS * 32)
attention_fwd_kernel(int Tt, int Tp, int heads, const T *__restrict__ qkv, float scale, T *__restrict__ out, T *__restrict__ probs,
                     int rows_per_cta) {
  extern __shared__ float sm[];
  float *Ks = sm, *Vs = Ks + Tp * ATT_PITCH, *Pw = Vs + Tp * ATT_PITCH;       // Pw: [ATT_WARPS][Tp] probability rows, [ATT_WARPS][64] q rows
  float *Qw = Pw + ATT_WARPS * Tp;
  const int b = blockIdx.z, h = blockIdx.y, inner = heads * ATT_DH;
  const long long ld = 3LL * inner;
  const T *base = qkv + (long long)b * Tp * ld + h * ATT_DH;
  att_load_tile(Ks, base + inner, ld, Tt, Tp);
  att_load_tile(Vs, base + 2 * inner, ld, Tt, Tp);
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float *pw = Pw + w * Tp, *qw = Qw + w * ATT_DH;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(Tp, r0 + rows_per_cta);
  for (int i = r0 + w; i < r1; i += ATT_WARPS) {
    T *orow = out + ((long long)b * Tp + i) * inner + h * ATT_DH;
    T *prow = probs + (((long long)b * heads + h) * Tp + i) * Tp;
    if (i >= Tt) {        // padding query row: defined zeros (its gradient is zero too)
      Cvt<T>::st(orow + lane, 0.f); Cvt<T>::st(orow + 32 + lane, 0.f);
      for (int j = lane; j < Tp; j += 32) Cvt<T>::st(prow + j, 0.f);
      continue;
    }
    const T *qrow = base + (long long)i * ld;
    qw[lane] = Cvt<T>::ld(qrow + lane); qw[32 + lane] = Cvt<T>::ld(qrow + 32 + lane);
    __syncwarp();
    float s[ATT_JMAX], mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < ATT_JMAX; ++jj) {
      const int j = jj * 32 + lane;
      s[jj] = -INFINITY;
      if (j < Tt) {
        float a = 0.f;
        const float *kp = Ks + j * ATT_PITCH;
#pragma unroll 16
        for (int d = 0; d < ATT_DH; ++d) a = fmaf(qw[d], kp[d], a);
        s[jj] = a * scale;
        mx = fmaxf(mx, s[jj]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < ATT_JMAX; ++jj) {
      const int j = jj * 32 + lane;
      s[jj] = (j < Tt) ? expf(s[jj] - mx) : 0.f;
      sum += s[jj];
    }
    const float inv = 1.f / warp_sum(sum);
#pragma unroll
    for (int jj = 0; jj < ATT_JMAX; ++jj) {
      const int j = jj * 32 + lane;
      if (j < Tp) {
        const float pq = round_as<T>(s[jj] * inv);       // the value the backward (and P.V below) sees is the stored one
        pw[j] = pq;
        Cvt<T>::st(prow + j, pq);
      }
    }
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < Tt; ++j) {
      const float pj = pw[j];
      o0 = fmaf(pj, Vs[j * ATT_PITCH + lane], o0);
      o1 = fmaf(pj, Vs[j * ATT_PITCH + 32 + lane], o1);
    }
    Cvt<T>::st(orow + lane, o0); Cvt<T>::st(orow + 32 + lane, o1);
    __syncwarp();
  }
}

// Backward, phase A (row-parallel): dP = dO V^T, dS = P*(dP - sum_j dP*P), dQ = scale * dS K; dS is written over `ds`.
template <typename T>
__global__ void __launch_bounds__(ATT_WARPS * 32)
attention_bwd_rows_kernel(int Tt, int Tp, int heads, const T *__restrict__ qkv, const T *__restrict__ probs, const T *__restrict__ dout,
                          float scale, T *__restrict__ dqkv, T *__restrict__ ds, int rows_per_cta) {
  extern __shared__ float sm[];
  float *Ks = sm, *Vs = Ks + Tp * ATT_PITCH, *Pw = Vs + Tp * ATT_PITCH;
  float *Qw = Pw + ATT_WARPS * Tp;
  const int b = blockIdx.z, h = blockIdx.y, inner = heads * ATT_DH;
  const long long ld = 3LL * inner;
  const T *base = qkv + (long long)b * Tp * ld + h * ATT_DH;
  att_load_tile(Ks, base + inner, ld, Tt, Tp);
  att_load_tile(Vs, base + 2 * inner, ld, Tt, Tp);
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float *pw = Pw + w * Tp, *qw = Qw + w * ATT_DH;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(Tp, r0 + rows_per_cta);
  for (int i = r0 + w; i < r1; i += ATT_WARPS) {
    T *dqrow = dqkv + ((long long)b * Tp + i) * ld + h * ATT_DH;
    T *dsrow = ds + (((long long)b * heads + h) * Tp + i) * Tp;
    if (i >= Tt) {
      Cvt<T>::st(dqrow + lane, 0.f); Cvt<T>::st(dqrow + 32 + lane, 0.f);
      for (int j = lane; j < Tp; j += 32) Cvt<T>::st(dsrow + j, 0.f);
      continue;
    }
    const T *dorow = dout + ((long long)b * Tp + i) * inner + h * ATT_DH;
    const T *prow = probs + (((long long)b * heads + h) * Tp + i) * Tp;
    qw[lane] = Cvt<T>::ld(dorow + lane); qw[32 + lane] = Cvt<T>::ld(dorow + 32 + lane);
    __syncwarp();
    float dp[ATT_JMAX], pv[ATT_JMAX], dot = 0.f;
#pragma unroll
    for (int jj = 0; jj < ATT_JMAX; ++jj) {
      const int j = jj * 32 + lane;
      dp[jj] = 0.f; pv[jj] = 0.f;
      if (j < Tt) {
        float a = 0.f;
        const float *vp = Vs + j * ATT_PITCH;
#pragma unroll 16
        for (int d = 0; d < ATT_DH; ++d) a = fmaf(qw[d], vp[d], a);
        dp[jj] = a; pv[jj] = Cvt<T>::ld(prow + j);
        dot += a * pv[jj];
      }
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int jj = 0; jj < ATT_JMAX; ++jj) {
      const int j = jj * 32 + lane;
      if (j < Tp) {
        const float v = round_as<T>((j < Tt) ? pv[jj] * (dp[jj] - dot) : 0.f);
        pw[j] = v;
        Cvt<T>::st(dsrow + j, v);
      }
    }
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < Tt; ++j) {
      const float sj = pw[j];
      o0 = fmaf(sj, Ks[j * ATT_PITCH + lane], o0);
      o1 = fmaf(sj, Ks[j * ATT_PITCH + 32 + lane], o1);
    }
    Cvt<T>::st(dqrow + lane, o0 * scale); Cvt<T>::st(dqrow + 32 + lane, o1 * scale);
    __syncwarp();
  }
}

// Backward, phase B (key-parallel): dK[j] = scale * sum_i dS[i][j] Q[i], dV[j] = sum_i P[i][j] dO[i].
template <typename T>
__global__ void __launch_bounds__(ATT_WARPS * 32)
attention_bwd_cols_kernel(int Tt, int Tp, int heads, const T *__restrict__ qkv, const T *__restrict__ probs, const T *__restrict__ ds,
                          const T *__restrict__ dout, float scale, T *__restrict__ dqkv, int rows_per_cta) {
  extern __shared__ float sm[];
  float *Qs = sm, *Os = Qs + Tp * ATT_PITCH, *Cw = Os + Tp * ATT_PITCH;       // Cw: [ATT_WARPS][2][Tp] columns of dS and P
  const int b = blockIdx.z, h = blockIdx.y, inner = heads * ATT_DH;
  const long long ld = 3LL * inner;
  att_load_tile(Qs, qkv + (long long)b * Tp * ld + h * ATT_DH, ld, Tt, Tp);
  att_load_tile(Os, dout + (long long)b * Tp * inner + h * ATT_DH, (long long)inner, Tt, Tp);
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float *cs = Cw + w * 2 * Tp, *cp = cs + Tp;
  const T *pb = probs + ((long long)b * heads + h) * Tp * Tp, *sb = ds + ((long long)b * heads + h) * Tp * Tp;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(Tp, r0 + rows_per_cta);
  for (int j = r0 + w; j < r1; j += ATT_WARPS) {
    T *dkrow = dqkv + ((long long)b * Tp + j) * ld + inner + h * ATT_DH;
    T *dvrow = dkrow + inner;
    if (j >= Tt) {
      Cvt<T>::st(dkrow + lane, 0.f); Cvt<T>::st(dkrow + 32 + lane, 0.f);
      Cvt<T>::st(dvrow + lane, 0.f); Cvt<T>::st(dvrow + 32 + lane, 0.f);
      continue;
    }
    for (int i = lane; i < Tt; i += 32) { cs[i] = Cvt<T>::ld(sb + (long long)i * Tp + j); cp[i] = Cvt<T>::ld(pb + (long long)i * Tp + j); }
    __syncwarp();
    float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
    for (int i = 0; i < Tt; ++i) {
      const float si = cs[i], pi = cp[i];
      k0 = fmaf(si, Qs[i * ATT_PITCH + lane], k0); k1 = fmaf(si, Qs[i * ATT_PITCH + 32 + lane], k1);
      v0 = fmaf(pi, Os[i * ATT_PITCH + lane], v0); v1 = fmaf(pi, Os[i * ATT_PITCH + 32 + lane], v1);
    }
    Cvt<T>::st(dkrow + lane, k0 * scale); Cvt<T>::st(dkrow + 32 + lane, k1 * scale);
    Cvt<T>::st(dvrow + lane, v0); Cvt<T>::st(dvrow + 32 + lane, v1);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------
// GELU (exact-erf form, nn.GELU default).  fp32 storage (parity mode): erff / expf.  bf16 storage: erf by Abramowitz-Stegun 7.1.26
// (|error| < 2e-7 + the MUFU approximations, three orders below bf16 resolution) - with erff the kernels were instruction-bound
// (~40 instructions per element) at half the HBM rate; the backward shares the one exponential between erf and the Gaussian.
// ---------------------------------------------------------------------------------------------------------
template <typename T> struct GeluMath {
  static __device__ __forceinline__ void eval(float x, float &cdf, float &pdf_x) {          // Phi(x), x * phi(x)
    cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
    pdf_x = x * 0.3989422804014327f * expf(-0.5f * x * x);
  }
};
template <> struct GeluMath<__nv_bfloat16> {
  static __device__ __forceinline__ void eval(float x, float &cdf, float &pdf_x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    float e, t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));     // exp(-x^2/2)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f); p = fmaf(p, t, -0.284496736f); p = fmaf(p, t, 0.254829592f);
    const float erfc_z = p * t * e;                                                          // 1 - erf(z), z >= 0
    cdf = (x >= 0.f) ? 1.f - 0.5f * erfc_z : 0.5f * erfc_z;
    pdf_x = x * 0.3989422804014327f * e;
  }
};
template <typename T>
__global__ void __launch_bounds__(256) gelu_fwd_kernel(long long n8, const T *__restrict__ u, T *__restrict__ hout) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    ldv8(u + i * 8, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { float c, px; GeluMath<T>::eval(f[k], c, px); f[k] *= c; }
    stv8(hout + i * 8, f);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) gelu_bwd_kernel(long long n8, const T *__restrict__ u, const T *__restrict__ dh, T *__restrict__ du) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float f[8], g[8];
    ldv8(u + i * 8, f); ldv8(dh + i * 8, g);
#pragma unroll
    for (int k = 0; k < 8; ++k) { float c, px; GeluMath<T>::eval(f[k], c, px); g[k] *= c + px; }
    stv8(du + i * 8, g);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Bilinear upsample (align_corners=False, aten upsample_bilinear2d) of a K-channel token map to NCHW fp32.
// src row of grid cell (gy,gx) of image b: b*Tp + row0 + gy*G + gx, channels 0..K-1 of a Cs-wide row.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bil_src(int d, float ratio, int in, int &i0, int &i1, float &l1) {
  float s = ratio * ((float)d + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s; if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + ((i0 < in - 1) ? 1 : 0);
  l1 = s - (float)i0;
}

template <typename T>
__global__ void __launch_bounds__(256)
bilinear_up_fwd_kernel(int B, int G, int Tp, int row0, int Cs, int K, int Ho, int Wo, const T *__restrict__ src, float *__restrict__ dst) {
  const long long total = (long long)B * Ho * Wo;
  const float ry = (float)G / (float)Ho, rx = (float)G / (float)Wo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wo), y = (int)((i / Wo) % Ho), b = (int)(i / ((long long)Wo * Ho));
    int y0, y1, x0, x1; float ly, lx;
    bil_src(y, ry, G, y0, y1, ly); bil_src(x, rx, G, x0, x1, lx);
    const T *base = src + ((long long)b * Tp + row0) * Cs;
    for (int k = 0; k < K; ++k) {
      const float a = Cvt<T>::ld(base + (long long)(y0 * G + x0) * Cs + k), bq = Cvt<T>::ld(base + (long long)(y0 * G + x1) * Cs + k);
      const float c = Cvt<T>::ld(base + (long long)(y1 * G + x0) * Cs + k), d = Cvt<T>::ld(base + (long long)(y1 * G + x1) * Cs + k);
      dst[(((long long)b * K + k) * Ho + y) * Wo + x] = (1.f - ly) * ((1.f - lx) * a + lx * bq) + ly * ((1.f - lx) * c + lx * d);
    }
  }
}

// Adjoint, gather form (deterministic): one WARP per (b, source row t in [0,Tp)); the lanes split the columns of the cell's
// footprint in the output image (coalesced row reads of the K gradient planes), the Cs outputs of the row are written by lanes 0..Cs-1.
// (One thread per (row, channel) left 3 of every Cs threads looping over a ~36 x 36 footprint: 0.3-0.5 ms per step.)
template <typename T>
__global__ void __launch_bounds__(256)
bilinear_up_bwd_kernel(int B, int G, int Tp, int row0, int Cs, int K, int Ho, int Wo, const float *__restrict__ ddst, T *__restrict__ dsrc) {
  const int lane = threadIdx.x & 31;
  const long long nwarp = ((long long)gridDim.x * blockDim.x) >> 5, total = (long long)B * Tp;
  const float ry = (float)G / (float)Ho, rx = (float)G / (float)Wo;
  for (long long item = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < total; item += nwarp) {
    const int t = (int)(item % Tp), b = (int)(item / Tp);
    const int cell = t - row0;
    float mine = 0.f;                                   // lane c keeps channel c's result
    if (cell >= 0 && cell < G * G) {
      const int gy = cell / G, gx = cell % G;
      int ylo = (int)floorf(((float)gy - 0.5f) / ry - 0.5f) - 1, yhi = (int)ceilf(((float)gy + 1.5f) / ry - 0.5f) + 1;
      int xlo = (int)floorf(((float)gx - 0.5f) / rx - 0.5f) - 1, xhi = (int)ceilf(((float)gx + 1.5f) / rx - 0.5f) + 1;
      if (gy == 0) ylo = 0; if (gy == G - 1) yhi = Ho - 1;
      if (gx == 0) xlo = 0; if (gx == G - 1) xhi = Wo - 1;
      ylo = max(ylo, 0); yhi = min(yhi, Ho - 1); xlo = max(xlo, 0); xhi = min(xhi, Wo - 1);
      for (int c = 0; c < K; ++c) {
        const float *gp = ddst + ((long long)b * K + c) * Ho * Wo;
        float acc = 0.f;
        for (int x = xlo + lane; x <= xhi; x += 32) {
          int x0, x1; float lx;
          bil_src(x, rx, G, x0, x1, lx);
          const float wx = ((x0 == gx) ? (1.f - lx) : 0.f) + ((x1 == gx) ? lx : 0.f);
          if (wx == 0.f) continue;
          float col = 0.f;
          for (int y = ylo; y <= yhi; ++y) {
            int y0, y1; float ly;
            bil_src(y, ry, G, y0, y1, ly);
            const float wy = ((y0 == gy) ? (1.f - ly) : 0.f) + ((y1 == gy) ? ly : 0.f);
            if (wy != 0.f) col = fmaf(wy, gp[(long long)y * Wo + x], col);
          }
          acc = fmaf(wx, col, acc);
        }
        acc = warp_sum(acc);
        if (lane == c) mine = acc;
      }
    }
    for (int c = lane; c < Cs; c += 32) Cvt<T>::st(dsrc + item * Cs + c, (c < K && c < 32) ? mine : 0.f);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Attention on the tensor cores (bf16 storage, fp32 accumulate; warp-level mma through nvcuda::wmma, 16x16x16 tiles).
// ~4 % of the ViT FLOPs; one CTA per (image, head), each warp owns 16-query (or 16-key) strips:
//   fwd  : S = Q K^T (13x4 mma per strip), row softmax in shared memory, P -> global (kept for the backward), O = P V
//   bwd A: dP = dO V^T, dS = P*(dP - rowsum(dP*P)) -> global scratch, dQ = scale * dS K
//   bwd B: dK = scale * dS^T Q, dV = P^T dO   (A operands read column-major straight from the L2-resident P / dS)
// ---------------------------------------------------------------------------------------------------------
namespace wm = nvcuda::wmma;
constexpr int ATC_WARPS = 7, ATC_SLAB = 24, ATC_KP = ATT_DH + 8;      // 72-element bf16 row pitch: 16-byte aligned rows, conflict-light ldmatrix

__device__ __forceinline__ void atc_load_tile(__nv_bfloat16 *dst, const __nv_bfloat16 *src, long long ld, int rows_valid, int rows_total) {
  for (int i = threadIdx.x; i < rows_total * (ATT_DH / 8); i += blockDim.x) {
    const int r = i / (ATT_DH / 8), c = (i % (ATT_DH / 8)) * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < rows_valid) v = *reinterpret_cast<const uint4 *>(src + (long long)r * ld + c);
    *reinterpret_cast<uint4 *>(dst + r * ATC_KP + c) = v;
  }
}

// write a 16 x 64 fp32 tile staged in shared memory (pitch sp) as bf16 rows to global (row pitch ld); lane pair per row
__device__ __forceinline__ void atc_store_rows(const float *sw, int sp, __nv_bfloat16 *dst, long long ld, int lane, int row0, int rows_valid,
                                               float mul) {
  const int r = lane >> 1, c0 = (lane & 1) * 32;
  const bool ok = (row0 + r) < rows_valid;
#pragma unroll
  for (int c = 0; c < 32; c += 8) {
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = ok ? sw[r * sp + c0 + c + k] * mul : 0.f;
    st8(dst + (long long)r * ld + c0 + c, f);
  }
}

// ---- register-resident ("flash"-style) strips on mma.sync.m16n8k16 ------------------------------------------------------
// One warp owns a 16-query strip; the whole score row block S[16 x Tp] lives in the accumulator fragments (Tp <= 256: 128
// registers), so the softmax is a per-lane pass plus two quad shuffles, and the bf16 probabilities are re-used IN REGISTERS as the
// A operand of P.V (accumulator layout of two adjacent n8 tiles == A layout of one k16 step).  K and V sit in shared memory
// (72-element pitch) and are read with ldmatrix (.trans for the [key][d] -> (k = key, n = d) operands).
constexpr int AM_WARPS = 8, AM_T8 = 32;       // up to 32 n8 tiles = 256 keys

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void *p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void *p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t *>(&h2);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u));
}
// A fragments (4 k16 steps over d = 64) of a 16-row strip read straight from global rows (pitch ld elements)
__device__ __forceinline__ void am_load_a(uint32_t (&a)[4][4], const __nv_bfloat16 *rows, long long ld, int g, int t) {
  const __nv_bfloat16 *r0 = rows + (long long)g * ld + 2 * t, *r1 = r0 + 8 * ld;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    a[kk][0] = *reinterpret_cast<const uint32_t *>(r0 + kk * 16);
    a[kk][1] = *reinterpret_cast<const uint32_t *>(r1 + kk * 16);
    a[kk][2] = *reinterpret_cast<const uint32_t *>(r0 + kk * 16 + 8);
    a[kk][3] = *reinterpret_cast<const uint32_t *>(r1 + kk * 16 + 8);
  }
}
// acc[jt] (+)= A[16 x 64] . M^T where M = [Tp][64] in shared memory (scores against K, or dO against V)
__device__ __forceinline__ void am_scores(float (&acc)[AM_T8][4], const uint32_t (&a)[4][4], const __nv_bfloat16 *Ms, int NT8, int lane) {
  // k-outer / tile-inner, two key tiles per step: consecutive MMAs hit different accumulators (the 4 k-steps of one tile are a
  // dependent chain; issued back to back they left the warp waiting on the MMA latency - ncu: IPC 0.19, "wait" the top stall)
#pragma unroll
  for (int jt = 0; jt < AM_T8; ++jt) acc[jt][0] = acc[jt][1] = acc[jt][2] = acc[jt][3] = 0.f;
  const __nv_bfloat16 *row = Ms + (lane & 7) * ATC_KP + (lane >> 3) * 8;
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2) {
#pragma unroll
    for (int jt = 0; jt < AM_T8; jt += 2) {
      if (jt < NT8) {                                            // NT8 is even (Tp % 16 == 0)
        uint32_t r0[4], r1[4];
        ldsm_x4(r0, row + jt * 8 * ATC_KP + h2 * 32);
        ldsm_x4(r1, row + (jt + 1) * 8 * ATC_KP + h2 * 32);
        mma16816(acc[jt], a[2 * h2], r0[0], r0[1]);
        mma16816(acc[jt + 1], a[2 * h2], r1[0], r1[1]);
        mma16816(acc[jt], a[2 * h2 + 1], r0[2], r0[3]);
        mma16816(acc[jt + 1], a[2 * h2 + 1], r1[2], r1[3]);
      }
    }
  }
}
// o[dt] += P[16 x Tp] . M where the bf16 P comes packed per n8 tile: pk[jt][0] = row g, pk[jt][1] = row g+8
__device__ __forceinline__ void am_apply(float (&o)[8][4], const uint32_t (&pk)[AM_T8][2], const __nv_bfloat16 *Ms, int NT16, int lane) {
#pragma unroll
  for (int kk = 0; kk < AM_T8 / 2; ++kk) {
    if (kk < NT16) {
      const uint32_t a[4] = {pk[2 * kk][0], pk[2 * kk][1], pk[2 * kk + 1][0], pk[2 * kk + 1][1]};
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t r[4];
        ldsm_x4_t(r, Ms + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ATC_KP + (dp * 2 + (lane >> 4)) * 8);
        mma16816(o[2 * dp], a, r[0], r[1]);
        mma16816(o[2 * dp + 1], a, r[2], r[3]);
      }
    }
  }
}
__device__ __forceinline__ void am_store_o(const float (&o)[8][4], __nv_bfloat16 *rows, long long ld, int g, int t, bool ok0, bool ok1, float mul) {
  __nv_bfloat16 *r0 = rows + (long long)g * ld + 2 * t, *r1 = r0 + 8 * ld;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    *reinterpret_cast<uint32_t *>(r0 + dt * 8) = ok0 ? pack_bf16(o[dt][0] * mul, o[dt][1] * mul) : 0u;
    *reinterpret_cast<uint32_t *>(r1 + dt * 8) = ok1 ? pack_bf16(o[dt][2] * mul, o[dt][3] * mul) : 0u;
  }
}

__global__ void __launch_bounds__(AM_WARPS * 32, 1)
attention_fwd_tc_kernel(int Tt, int Tp, int heads, int hpc, const __nv_bfloat16 *__restrict__ qkv, float scale, __nv_bfloat16 *__restrict__ out,
                        __nv_bfloat16 *__restrict__ probs) {
  extern __shared__ __align__(128) unsigned char smraw[];
  // hpc heads of one image per CTA: 13 strips over 8 warps leave 3 of 16 warp slots idle, 3 heads = 39 strips leave 1 of 40
  const int b = blockIdx.y, inner = heads * ATT_DH, NT8 = Tp / 8, NT16 = Tp / 16;
  const long long ld = 3LL * inner;
  for (int hl = 0; hl < hpc; ++hl) {
    const __nv_bfloat16 *hb = qkv + (long long)b * Tp * ld + (blockIdx.x * hpc + hl) * ATT_DH;
    atc_load_tile(reinterpret_cast<__nv_bfloat16 *>(smraw) + (size_t)hl * 2 * Tp * ATC_KP, hb + inner, ld, Tt, Tp);
    atc_load_tile(reinterpret_cast<__nv_bfloat16 *>(smraw) + (size_t)(hl * 2 + 1) * Tp * ATC_KP, hb + 2 * inner, ld, Tt, Tp);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  for (int item = w; item < hpc * NT16; item += AM_WARPS) {
    const int hl = item / NT16, strip = item - hl * NT16, h = blockIdx.x * hpc + hl;
    const __nv_bfloat16 *Ks = reinterpret_cast<__nv_bfloat16 *>(smraw) + (size_t)hl * 2 * Tp * ATC_KP, *Vs = Ks + Tp * ATC_KP;
    const __nv_bfloat16 *base = qkv + (long long)b * Tp * ld + h * ATT_DH;
    const int i0 = strip * 16;
    const bool ok0 = (i0 + g) < Tt, ok1 = (i0 + g + 8) < Tt;
    uint32_t qa[4][4];
    am_load_a(qa, base + (long long)i0 * ld, ld, g, t);
    float s[AM_T8][4];
    am_scores(s, qa, Ks, NT8, lane);
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int jt = 0; jt < AM_T8; ++jt) {
      if (jt < NT8) {
        const int c = jt * 8 + 2 * t;
        if (c < Tt) { m0 = fmaxf(m0, s[jt][0]); m1 = fmaxf(m1, s[jt][2]); }
        if (c + 1 < Tt) { m0 = fmaxf(m0, s[jt][1]); m1 = fmaxf(m1, s[jt][3]); }
      }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
    const float sl2 = scale * 1.4426950408889634f, m0s = m0 * sl2, m1s = m1 * sl2;
#pragma unroll
    for (int jt = 0; jt < AM_T8; ++jt) {
      if (jt < NT8) {
        const int c = jt * 8 + 2 * t;
        s[jt][0] = (c < Tt) ? ex2_approx(fmaf(s[jt][0], sl2, -m0s)) : 0.f;      // exp((s - m) * scale) as one FFMA + MUFU.EX2
        s[jt][1] = (c + 1 < Tt) ? ex2_approx(fmaf(s[jt][1], sl2, -m0s)) : 0.f;
        s[jt][2] = (c < Tt) ? ex2_approx(fmaf(s[jt][2], sl2, -m1s)) : 0.f;
        s[jt][3] = (c + 1 < Tt) ? ex2_approx(fmaf(s[jt][3], sl2, -m1s)) : 0.f;
        l0 += s[jt][0] + s[jt][1]; l1 += s[jt][2] + s[jt][3];
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = ok0 ? 1.f / l0 : 0.f, inv1 = ok1 ? 1.f / l1 : 0.f;
    uint32_t pk[AM_T8][2];
    __nv_bfloat16 *pr0 = probs + (((long long)b * heads + h) * Tp + i0 + g) * Tp + 2 * t, *pr1 = pr0 + 8LL * Tp;
#pragma unroll
    for (int jt = 0; jt < AM_T8; ++jt) {
      if (jt < NT8) {
        pk[jt][0] = pack_bf16(s[jt][0] * inv0, s[jt][1] * inv0);
        pk[jt][1] = pack_bf16(s[jt][2] * inv1, s[jt][3] * inv1);
        *reinterpret_cast<uint32_t *>(pr0 + jt * 8) = pk[jt][0];
        *reinterpret_cast<uint32_t *>(pr1 + jt * 8) = pk[jt][1];
      }
    }
    float o[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
    am_apply(o, pk, Vs, NT16, lane);
    am_store_o(o, out + ((long long)b * Tp + i0) * inner + h * ATT_DH, inner, g, t, ok0, ok1, 1.f);
  }
}

__global__ void __launch_bounds__(AM_WARPS * 32, 1)
attention_bwd_rows_tc_kernel(int Tt, int Tp, int heads, int hpc, const __nv_bfloat16 *__restrict__ qkv, const __nv_bfloat16 *__restrict__ probs,
                             const __nv_bfloat16 *__restrict__ dout, float scale, __nv_bfloat16 *__restrict__ dqkv,
                             __nv_bfloat16 *__restrict__ ds) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int b = blockIdx.y, inner = heads * ATT_DH, NT8 = Tp / 8, NT16 = Tp / 16;
  const long long ld = 3LL * inner;
  for (int hl = 0; hl < hpc; ++hl) {
    const __nv_bfloat16 *hb = qkv + (long long)b * Tp * ld + (blockIdx.x * hpc + hl) * ATT_DH;
    atc_load_tile(reinterpret_cast<__nv_bfloat16 *>(smraw) + (size_t)hl * 2 * Tp * ATC_KP, hb + inner, ld, Tt, Tp);
    atc_load_tile(reinterpret_cast<__nv_bfloat16 *>(smraw) + (size_t)(hl * 2 + 1) * Tp * ATC_KP, hb + 2 * inner, ld, Tt, Tp);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  for (int item = w; item < hpc * NT16; item += AM_WARPS) {
    const int hl = item / NT16, strip = item - hl * NT16, h = blockIdx.x * hpc + hl;
    const __nv_bfloat16 *Ks = reinterpret_cast<__nv_bfloat16 *>(smraw) + (size_t)hl * 2 * Tp * ATC_KP, *Vs = Ks + Tp * ATC_KP;
    const __nv_bfloat16 *dob = dout + (long long)b * Tp * inner + h * ATT_DH;
    const int i0 = strip * 16;
    const bool ok0 = (i0 + g) < Tt, ok1 = (i0 + g + 8) < Tt;
    uint32_t da[4][4];
    am_load_a(da, dob + (long long)i0 * inner, inner, g, t);
    float dp[AM_T8][4];
    am_scores(dp, da, Vs, NT8, lane);                          // dP = dO V^T
    const __nv_bfloat16 *pr0 = probs + (((long long)b * heads + h) * Tp + i0 + g) * Tp + 2 * t, *pr1 = pr0 + 8LL * Tp;
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int jt = 0; jt < AM_T8; ++jt) {
      if (jt < NT8) {
        const float2 p0 = unpack_bf16(*reinterpret_cast<const uint32_t *>(pr0 + jt * 8));
        const float2 p1 = unpack_bf16(*reinterpret_cast<const uint32_t *>(pr1 + jt * 8));
        d0 += dp[jt][0] * p0.x + dp[jt][1] * p0.y;
        d1 += dp[jt][2] * p1.x + dp[jt][3] * p1.y;
      }
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    uint32_t pk[AM_T8][2];
    __nv_bfloat16 *sr0 = ds + (((long long)b * heads + h) * Tp + i0 + g) * Tp + 2 * t, *sr1 = sr0 + 8LL * Tp;
#pragma unroll
    for (int jt = 0; jt < AM_T8; ++jt) {
      if (jt < NT8) {
        const float2 p0 = unpack_bf16(*reinterpret_cast<const uint32_t *>(pr0 + jt * 8));
        const float2 p1 = unpack_bf16(*reinterpret_cast<const uint32_t *>(pr1 + jt * 8));
        pk[jt][0] = ok0 ? pack_bf16(p0.x * (dp[jt][0] - d0), p0.y * (dp[jt][1] - d0)) : 0u;     // P is 0 for keys >= T
        pk[jt][1] = ok1 ? pack_bf16(p1.x * (dp[jt][2] - d1), p1.y * (dp[jt][3] - d1)) : 0u;
        *reinterpret_cast<uint32_t *>(sr0 + jt * 8) = pk[jt][0];
        *reinterpret_cast<uint32_t *>(sr1 + jt * 8) = pk[jt][1];
      }
    }
    float o[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
    am_apply(o, pk, Ks, NT16, lane);                           // dQ = dS K
    am_store_o(o, dqkv + ((long long)b * Tp + i0) * ld + h * ATT_DH, ld, g, t, ok0, ok1, scale);
  }
}

// dK = scale * dS^T Q and dV = P^T dO for one (batch, head): one warp owns a 16-KEY strip.  Its 16 columns of P and dS for all
// queries are staged once in a per-warp shared-memory slab ([Tp][ATC_SLAB], every load in flight together); ldmatrix.trans then hands
// out the transposed A fragments (A[key][query] = S[query][key]) and the [query][d] B fragments of Q / dO, and the strip's two
// 16 x 64 results stay in mma.sync accumulators until they are stored.
__global__ void __launch_bounds__(AM_WARPS * 32, 2)
attention_bwd_cols_tc_kernel(int Tt, int Tp, int heads, const __nv_bfloat16 *__restrict__ qkv, const __nv_bfloat16 *__restrict__ probs,
                             const __nv_bfloat16 *__restrict__ ds, const __nv_bfloat16 *__restrict__ dout, float scale,
                             __nv_bfloat16 *__restrict__ dqkv) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int NT = Tp / 16;
  __nv_bfloat16 *Qs = reinterpret_cast<__nv_bfloat16 *>(smraw), *Os = Qs + Tp * ATC_KP, *slabs = Os + Tp * ATC_KP;
  const int b = blockIdx.y, h = blockIdx.x, inner = heads * ATT_DH;
  const long long ld = 3LL * inner;
  atc_load_tile(Qs, qkv + (long long)b * Tp * ld + h * ATT_DH, ld, Tt, Tp);
  atc_load_tile(Os, dout + (long long)b * Tp * inner + h * ATT_DH, (long long)inner, Tt, Tp);
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  // per-warp double-buffered slab: [2 buffers][P, dS][32 queries][ATC_SLAB]
  __nv_bfloat16 *Ww = slabs + (size_t)w * 2 * 2 * 32 * ATC_SLAB;
  const __nv_bfloat16 *pb = probs + ((long long)b * heads + h) * Tp * Tp, *sb = ds + ((long long)b * heads + h) * Tp * Tp;
  // ldmatrix.x4.trans source rows: matrices 0..3 = (queries 0-7, keys 0-7), (q 0-7, k 8-15), (q 8-15, k 0-7), (q 8-15, k 8-15)
  const int a_off = ((lane & 7) + (lane >> 4) * 8) * ATC_SLAB + ((lane >> 3) & 1) * 8;
  const int lr = lane >> 1, lh = (lane & 1) * 8;               // this lane's slab rows lr, lr + 16 and 8-key half
  const int NC = (Tp + 31) / 32;                               // 32-query chunks (the last one may hold 16)
  for (int strip = w; strip < NT; strip += (int)(blockDim.x >> 5)) {
    const int j0 = strip * 16;
    float ok[8][4], ov[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) { ok[dt][0] = ok[dt][1] = ok[dt][2] = ok[dt][3] = 0.f; ov[dt][0] = ov[dt][1] = ov[dt][2] = ov[dt][3] = 0.f; }
    uint4 vp[2], vs[2];
    auto fetch = [&](int c) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int r = c * 32 + q * 16 + lr;
        if (r < Tp) {
          vp[q] = *reinterpret_cast<const uint4 *>(pb + (long long)r * Tp + j0 + lh);
          vs[q] = *reinterpret_cast<const uint4 *>(sb + (long long)r * Tp + j0 + lh);
        }
      }
    };
    fetch(0);
    for (int c = 0; c < NC; ++c) {
      __nv_bfloat16 *Pw = Ww + (c & 1) * 2 * 32 * ATC_SLAB, *Dw = Pw + 32 * ATC_SLAB;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        *reinterpret_cast<uint4 *>(Pw + (q * 16 + lr) * ATC_SLAB + lh) = vp[q];
        *reinterpret_cast<uint4 *>(Dw + (q * 16 + lr) * ATC_SLAB + lh) = vs[q];
      }
      __syncwarp();
      if (c + 1 < NC) fetch(c + 1);                            // the next chunk's loads fly while this chunk's MMAs run
#pragma unroll
      for (int k2 = 0; k2 < 2; ++k2) {
        const int kk = c * 2 + k2;                             // 16 queries per step
        if (kk < NT) {
          uint32_t as[4], ap[4];
          ldsm_x4_t(as, Dw + k2 * 16 * ATC_SLAB + a_off);
          ldsm_x4_t(ap, Pw + k2 * 16 * ATC_SLAB + a_off);
          const int brow = (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ATC_KP + (lane >> 4) * 8;
#pragma unroll
          for (int dp = 0; dp < 4; ++dp) {
            uint32_t rq[4], ro[4];
            ldsm_x4_t(rq, Qs + brow + dp * 16);
            ldsm_x4_t(ro, Os + brow + dp * 16);
            mma16816(ok[2 * dp], as, rq[0], rq[1]);
            mma16816(ov[2 * dp], ap, ro[0], ro[1]);
            mma16816(ok[2 * dp + 1], as, rq[2], rq[3]);
            mma16816(ov[2 * dp + 1], ap, ro[2], ro[3]);
          }
        }
      }
      // buffer (c & 1) is rewritten at chunk c + 2; the __syncwarp() of chunk c + 1 orders these ldmatrix reads before that store
    }
    __syncwarp();
    const bool ok0 = (j0 + g) < Tt, ok1 = (j0 + g + 8) < Tt;
    __nv_bfloat16 *dk = dqkv + ((long long)b * Tp + j0) * ld + inner + h * ATT_DH;
    am_store_o(ok, dk, ld, g, t, ok0, ok1, scale);
    am_store_o(ov, dk + inner, ld, g, t, ok0, ok1, 1.f);
  }
}

static inline size_t atc_smem_rows(int Tp) { return (size_t)2 * Tp * ATC_KP * 2 + 128; }
// heads of one image per CTA for the row kernels: the count (dividing `heads`, fitting shared memory) whose strips fill the 8 warps' rounds best
static inline int atc_heads_per_cta(int Tp, int heads) {
  const int nt = Tp / 16;
  int best = 1; double beff = -1.0;
  for (int c = 1; c <= heads && (size_t)c * atc_smem_rows(Tp) <= 220 * 1024; ++c) {
    if (heads % c) continue;
    const int rounds = (c * nt + AM_WARPS - 1) / AM_WARPS;
    const double eff = (double)(c * nt) / (double)(rounds * AM_WARPS);
    if (eff > beff + 1e-9) { beff = eff; best = c; }
  }
  return best;
}
static inline size_t atc_smem_cols(int Tp, int warps) { return (size_t)2 * Tp * ATC_KP * 2 + (size_t)warps * 2 * 2 * 32 * ATC_SLAB * 2 + 128; }   // Q, dO + per-warp slabs
static inline int atc_cols_warps(int Tp) {
  int w = AM_WARPS;
  while (w > 1 && atc_smem_cols(Tp, w) > 220 * 1024) --w;
  return w;
}
static inline bool atc_ok(int dtype, int T, int Tp, int dh, int heads) {
  return dtype == KS_BF16 && !g_opt.att_simt && dh == ATT_DH && Tp % 16 == 0 && T <= Tp && atc_smem_rows(Tp) <= 220 * 1024 &&
         atc_smem_cols(Tp, atc_cols_warps(Tp)) <= 220 * 1024 &&
         (Tp * 2) % 16 == 0 && ((3LL * heads * dh) % 8) == 0;
}

// ---------------------------------------------------------------------------------------------------------
// nn.AdaptiveAvgPool2d(S) on NHWC views (UPerNet pyramid pooling, HF modeling_upernet.py UperNetPyramidPoolingBlock):
// bin (i, j) averages rows [floor(i*H/S), ceil((i+1)*H/S)) x cols [floor(j*W/S), ceil((j+1)*W/S)).  _bwd is the adjoint.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
adaptive_pool_fwd_kernel(View src, int N, int H, int W, int S, T *__restrict__ dst) {
  const int C = src.C;
  const long long total = (long long)N * S * S * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C); long long r = i / C;
    const int j = (int)(r % S); r /= S; const int bi = (int)(r % S); const long long n = r / S;
    const int h0 = (bi * H) / S, h1 = ((bi + 1) * H + S - 1) / S, w0 = (j * W) / S, w1 = ((j + 1) * W + S - 1) / S;
    float a = 0.f;
    for (int h = h0; h < h1; ++h)
      for (int w = w0; w < w1; ++w)
        a += Cvt<T>::ld(reinterpret_cast<const T *>(src.ptr) + (n * src.sn + (long long)h * src.sh + (long long)w * src.sw + c));
    Cvt<T>::st(dst + i, a / (float)((h1 - h0) * (w1 - w0)));
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
adaptive_pool_bwd_kernel(View dsrc, int N, int H, int W, int S, const T *__restrict__ ddst, int accumulate) {
  const int C = dsrc.C;
  const long long total = (long long)N * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C); long long r = i / C;
    const int w = (int)(r % W); r /= W; const int h = (int)(r % H); const long long n = r / H;
    float a = 0.f;
    for (int bi = 0; bi < S; ++bi) {
      const int h0 = (bi * H) / S, h1 = ((bi + 1) * H + S - 1) / S;
      if (h < h0 || h >= h1) continue;
      for (int j = 0; j < S; ++j) {
        const int w0 = (j * W) / S, w1 = ((j + 1) * W + S - 1) / S;
        if (w < w0 || w >= w1) continue;
        a += Cvt<T>::ld(ddst + ((n * S + bi) * S + j) * C + c) / (float)((h1 - h0) * (w1 - w0));
      }
    }
    T *p = reinterpret_cast<T *>(dsrc.ptr) + (n * dsrc.sn + (long long)h * dsrc.sh + (long long)w * dsrc.sw + c);
    Cvt<T>::st(p, accumulate ? Cvt<T>::ld(p) + a : a);
  }
}

// the same, 8 channels per thread (16-byte accesses)
template <typename T>
__global__ void __launch_bounds__(256)
adaptive_pool_bwd_vec_kernel(View dsrc, int N, int H, int W, int S, const T *__restrict__ ddst, int accumulate) {
  const int C = dsrc.C, CV = C / 8;
  const long long total = (long long)N * H * W * CV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % CV) * 8; long long r = i / CV;
    const int w = (int)(r % W); r /= W; const int h = (int)(r % H); const long long n = r / H;
    T *p = reinterpret_cast<T *>(dsrc.ptr) + (n * dsrc.sn + (long long)h * dsrc.sh + (long long)w * dsrc.sw + c);
    float a[8];
    if (accumulate) ldv8(p, a); else {
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = 0.f;
    }
    for (int bi = 0; bi < S; ++bi) {
      const int h0 = (bi * H) / S, h1 = ((bi + 1) * H + S - 1) / S;
      if (h < h0 || h >= h1) continue;
      for (int j = 0; j < S; ++j) {
        const int w0 = (j * W) / S, w1 = ((j + 1) * W + S - 1) / S;
        if (w < w0 || w >= w1) continue;
        float g[8];
        ldv8(ddst + ((n * S + bi) * S + j) * C + c, g);
        const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = fmaf(g[k], inv, a[k]);
      }
    }
    stv8(p, a);
  }
}

static inline int grid_for(long long work, int per_block, int cap_mult = 8) {
  long long g = (work + per_block - 1) / per_block;
  const long long cap = (long long)kNumSMs * cap_mult;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace ks

using namespace ks;

#define KS_DISPATCH_T(dtype, CALL)                                  \
  do {                                                              \
    if ((dtype) == KS_F32) { CALL(float); }                         \
    else if ((dtype) == KS_BF16) { CALL(__nv_bfloat16); }           \
    else return KS_EINVAL;                                          \
  } while (0)

static inline bool al16(const void *p) { return ((uintptr_t)p % 16) == 0; }

extern "C" int ks_layernorm_fwd(int dtype, int64_t rows, int C, const void *x, int64_t ldx, const float *gamma, const float *beta,
                                float eps, void *y, int64_t ldy, float *mean, float *rstd, void *copy_out, int64_t ldc, void *stream) {
  KS_CHECK_ARG(rows > 0 && C > 0 && x && y && gamma && beta);
  if (C % 8 || C > LN_KMAX * 256 || ldx % 8 || ldy % 8 || (copy_out && ldc % 8) || !al16(x) || !al16(y) || !al16(gamma) || !al16(beta) ||
      (copy_out && !al16(copy_out))) return KS_EUNSUPPORTED;
  if (C <= 256) {
    int L = 1;
    while (L * 8 < C) L <<= 1;
    const int grid = grid_for(rows, 8 * (32 / L) * 2, 8);
#define CALL(T) layernorm_fwd_narrow_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(rows, C, L, (const T *)x, ldx, gamma, beta, eps, (T *)y, ldy, \
                                                                                         mean, rstd, (T *)copy_out, ldc)
    KS_DISPATCH_T(dtype, CALL);
#undef CALL
    KS_LAUNCH_RET();
  }
  const int grid = grid_for(rows, 8);
#define CALL(T) layernorm_fwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(rows, C, (const T *)x, ldx, gamma, beta, eps, (T *)y, ldy, \
                                                                                  mean, rstd, (T *)copy_out, ldc)
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_layernorm_bwd(int dtype, int64_t rows, int C, const void *dy, int64_t lddy, const void *x, int64_t ldx,
                                const float *mean, const float *rstd, const float *gamma, void *dx, int64_t lddx, int accumulate_dx,
                                float *dgamma, float *dbeta, void *stream) {
  KS_CHECK_ARG(rows > 0 && C > 0 && dy && x && mean && rstd && gamma);
  if (C % 8 || C > LNB_KMAX * 256 || lddy % 8 || ldx % 8 || (dx && lddx % 8) || !al16(dy) || !al16(x) || !al16(gamma) || (dx && !al16(dx)))
    return KS_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  int L = 1;
  while (L < 32 && L * 8 < C) L <<= 1;
  const int K = (C + 8 * L - 1) / (8 * L);
  if (dx) {
    const int grid = grid_for(rows, 8 * (32 / L) * 2, 4);
#define CALL(T) { \
    if (K == 1) layernorm_bwd_rows_kernel<T, 1><<<grid, 256, 0, st>>>(rows, C, L, (const T *)dy, lddy, (const T *)x, ldx, mean, rstd, gamma, (T *)dx, lddx, accumulate_dx); \
    else if (K == 2) layernorm_bwd_rows_kernel<T, 2><<<grid, 256, 0, st>>>(rows, C, L, (const T *)dy, lddy, (const T *)x, ldx, mean, rstd, gamma, (T *)dx, lddx, accumulate_dx); \
    else if (K == 3) layernorm_bwd_rows_kernel<T, 3><<<grid, 256, 0, st>>>(rows, C, L, (const T *)dy, lddy, (const T *)x, ldx, mean, rstd, gamma, (T *)dx, lddx, accumulate_dx); \
    else layernorm_bwd_rows_kernel<T, 4><<<grid, 256, 0, st>>>(rows, C, L, (const T *)dy, lddy, (const T *)x, ldx, mean, rstd, gamma, (T *)dx, lddx, accumulate_dx); }
    KS_DISPATCH_T(dtype, CALL);
#undef CALL
  }
  if (dgamma || dbeta) {
    const int cvb = (C / 8) < 256 ? (C / 8) : 256, rl = 256 / cvb;
    // >= 16 rows per thread so that the per-CTA column atomics amortise; measured (scripts/bench_lnbwd.py): C = 768 18.3 -> 14.3 us with 32, C = 128 better with 16
    const int grid = grid_for(rows, rl * (g_opt.ln_rows > 0 ? g_opt.ln_rows : (C >= 512 ? 32 : 16)), 4);
    const size_t smem = (size_t)2 * C * sizeof(float);
#define CALL(T) layernorm_bwd_cols_kernel<T><<<grid, 256, smem, st>>>(rows, C, (const T *)dy, lddy, (const T *)x, ldx, mean, rstd, dgamma, dbeta)
    KS_DISPATCH_T(dtype, CALL);
#undef CALL
  }
  KS_LAUNCH_RET();
}

extern "C" int ks_patchify_ln(int dtype, int B, int Cc, int Hi, int Wi, int Tp, const float *img, const float *gamma, const float *beta,
                              float eps, void *out, float *mean, float *rstd, void *stream) {
  KS_CHECK_ARG(B > 0 && img && gamma && beta && out && mean && rstd);
  if (Cc < 1 || Cc > PCMAX || Hi % PATCH || Wi % PATCH || (Hi / PATCH) * (Wi / PATCH) + 1 > Tp || !al16(img)) return KS_EUNSUPPORTED;
  const int grid = grid_for((long long)B * (Hi / PATCH) * (Wi / PATCH), 8);
#define CALL(T) patchify_ln_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(B, Cc, Hi, Wi, Tp, img, gamma, beta, eps, (T *)out, mean, rstd)
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_patchify_ln_bwd(int dtype, int B, int Cc, int Hi, int Wi, int Tp, const float *img, const float *mean, const float *rstd,
                                  const void *dy, float *dgamma, float *dbeta, void *stream) {
  KS_CHECK_ARG(B > 0 && img && mean && rstd && dy && dgamma && dbeta);
  if (Cc < 1 || Cc > PCMAX || Hi % PATCH || Wi % PATCH) return KS_EUNSUPPORTED;
  const int grid = grid_for((long long)B * (Hi / PATCH) * (Wi / PATCH), 8 * 8, 2);
  const size_t smem = (size_t)2 * PATCH * PATCH * Cc * sizeof(float);
#define CALL(T) patchify_ln_bwd_kernel<T><<<grid, 256, smem, (cudaStream_t)stream>>>(B, Cc, Hi, Wi, Tp, img, mean, rstd, (const T *)dy, dgamma, dbeta)
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_vit_assemble(int dtype, int B, int T, int Tp, int D, const void *e, const float *cls, const float *pos, void *x0,
                               void *stream) {
  KS_CHECK_ARG(B > 0 && T > 0 && Tp >= T && e && cls && pos && x0);
  if (D % 8 || !al16(e) || !al16(x0) || !al16(cls) || !al16(pos)) return KS_EUNSUPPORTED;
  const int grid = grid_for((long long)B * Tp * (D / 8), 256);
#define CALL(Ty) vit_assemble_kernel<Ty><<<grid, 256, 0, (cudaStream_t)stream>>>(B, T, Tp, D, (const Ty *)e, cls, pos, (Ty *)x0)
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_vit_assemble_bwd(int dtype, int B, int T, int Tp, int D, const void *dx0, void *de, float *dcls, float *dpos, void *stream) {
  KS_CHECK_ARG(B > 0 && T > 0 && Tp >= T && dx0 && de && dcls && dpos);
  if (D % 8 || !al16(dx0) || !al16(de) || !al16(dcls) || !al16(dpos)) return KS_EUNSUPPORTED;
  dim3 grid((unsigned)Tp, (unsigned)((D / 8 + 127) / 128));
#define CALL(Ty) vit_assemble_bwd_kernel<Ty><<<grid, 128, 0, (cudaStream_t)stream>>>(B, T, Tp, D, (const Ty *)dx0, (Ty *)de, dcls, dpos)
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

static int att_cfg(int T, int Tp, int dh, int extra_rows, size_t &smem, int &rows_per_cta, int &nblk, int B, int heads) {
  if (dh != ATT_DH || Tp > ATT_JMAX * 32 || T > Tp || Tp % 8) return KS_EUNSUPPORTED;
  smem = ((size_t)2 * Tp * ATT_PITCH + (size_t)ATT_WARPS * extra_rows * Tp + (size_t)ATT_WARPS * ATT_DH) * sizeof(float);
  if (smem > 220 * 1024) return KS_EUNSUPPORTED;
  // enough CTAs to fill the machine: split the rows of one (b, head) when B*heads is small
  nblk = 1;
  while ((long long)B * heads * nblk < 2 * kNumSMs && nblk < 8) nblk <<= 1;
  rows_per_cta = ((Tp + nblk - 1) / nblk + ATT_WARPS - 1) / ATT_WARPS * ATT_WARPS;
  nblk = (Tp + rows_per_cta - 1) / rows_per_cta;
  return KS_OK;
}

namespace ks { int attention_bwd_umma(int B, int T, int Tp, int heads, const void *qkv, const void *probs, const void *dout, float scale, void *dqkv,
                                      void *ds, cudaStream_t st); }
namespace ks { int attention_fwd_umma(int B, int T, int Tp, int heads, const void *qkv, float scale, void *out, void *probs, cudaStream_t st); }

extern "C" int ks_attention_fwd(int dtype, int B, int T, int Tp, int heads, int dh, const void *qkv, float scale, void *out, void *probs,
                                void *stream) {
  KS_CHECK_ARG(B > 0 && T > 0 && heads > 0 && qkv && out && probs);
  if (!al16(qkv) || !al16(out) || !al16(probs)) return KS_EUNSUPPORTED;
  if (atc_ok(dtype, T, Tp, dh, heads) && !g_opt.att_no_umma) {
    // tcgen05 path (csrc/attention_tc.cu): TMA -> smem -> UMMA -> TMEM, softmax out of TMEM; falls through when the shape does not fit
    const int rc = attention_fwd_umma(B, T, Tp, heads, qkv, scale, out, probs, (cudaStream_t)stream);
    if (rc != KS_EUNSUPPORTED) return rc;
  }
  if (atc_ok(dtype, T, Tp, dh, heads)) {
    static bool attr = false;
    if (!attr) { cudaError_t e = cudaFuncSetAttribute(attention_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
                 if (e != cudaSuccess) return (int)e; attr = true; }
    const int hpc = atc_heads_per_cta(Tp, heads);
    attention_fwd_tc_kernel<<<dim3((unsigned)(heads / hpc), (unsigned)B), AM_WARPS * 32, hpc * atc_smem_rows(Tp), (cudaStream_t)stream>>>(
        T, Tp, heads, hpc, (const __nv_bfloat16 *)qkv, scale, (__nv_bfloat16 *)out, (__nv_bfloat16 *)probs);
    KS_LAUNCH_RET();
  }
  size_t smem; int rpc, nblk;
  int rc = att_cfg(T, Tp, dh, 1, smem, rpc, nblk, B, heads); if (rc) return rc;
  dim3 grid((unsigned)nblk, (unsigned)heads, (unsigned)B);
#define CALL(Ty) { cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel<Ty>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); \
    if (e != cudaSuccess) return (int)e; \
    attention_fwd_kernel<Ty><<<grid, ATT_WARPS * 32, smem, (cudaStream_t)stream>>>(T, Tp, heads, (const Ty *)qkv, scale, (Ty *)out, (Ty *)probs, rpc); }
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_attention_bwd(int dtype, int B, int T, int Tp, int heads, int dh, const void *qkv, const void *probs, const void *dout,
                                float scale, void *dqkv, void *ds_scratch, void *stream) {
  KS_CHECK_ARG(B > 0 && T > 0 && heads > 0 && qkv && probs && dout && dqkv && ds_scratch);
  if (!al16(qkv) || !al16(dout) || !al16(probs) || !al16(dqkv) || !al16(ds_scratch)) return KS_EUNSUPPORTED;
  if (atc_ok(dtype, T, Tp, dh, heads) && !g_opt.att_no_umma) {
    const int rc = attention_bwd_umma(B, T, Tp, heads, qkv, probs, dout, scale, dqkv, ds_scratch, (cudaStream_t)stream);
    if (rc != KS_EUNSUPPORTED) return rc;
  }
  if (atc_ok(dtype, T, Tp, dh, heads)) {
    static bool attr = false;
    if (!attr) {
      cudaError_t e = cudaFuncSetAttribute(attention_bwd_rows_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      if (e != cudaSuccess) return (int)e;
      e = cudaFuncSetAttribute(attention_bwd_cols_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      if (e != cudaSuccess) return (int)e;
      attr = true;
    }
    const dim3 g((unsigned)heads, (unsigned)B);
    const int hpc = 1;        // several heads per CTA measured slower here (3.33 -> 3.69 ms): the 180 KB of shared memory leave too little L1 for the re-read P rows
    attention_bwd_rows_tc_kernel<<<dim3((unsigned)(heads / hpc), (unsigned)B), AM_WARPS * 32, hpc * atc_smem_rows(Tp), (cudaStream_t)stream>>>(
        T, Tp, heads, hpc, (const __nv_bfloat16 *)qkv, (const __nv_bfloat16 *)probs, (const __nv_bfloat16 *)dout, scale, (__nv_bfloat16 *)dqkv,
        (__nv_bfloat16 *)ds_scratch);
    attention_bwd_cols_tc_kernel<<<g, atc_cols_warps(Tp) * 32, atc_smem_cols(Tp, atc_cols_warps(Tp)), (cudaStream_t)stream>>>(
        T, Tp, heads, (const __nv_bfloat16 *)qkv, (const __nv_bfloat16 *)probs, (const __nv_bfloat16 *)ds_scratch, (const __nv_bfloat16 *)dout,
        scale, (__nv_bfloat16 *)dqkv);
    KS_LAUNCH_RET();
  }
  size_t smem; int rpc, nblk;
  int rc = att_cfg(T, Tp, dh, 2, smem, rpc, nblk, B, heads); if (rc) return rc;
  dim3 grid((unsigned)nblk, (unsigned)heads, (unsigned)B);
#define CALL(Ty) { cudaError_t e = cudaFuncSetAttribute(attention_bwd_rows_kernel<Ty>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); \
    if (e != cudaSuccess) return (int)e; \
    e = cudaFuncSetAttribute(attention_bwd_cols_kernel<Ty>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); \
    if (e != cudaSuccess) return (int)e; \
    attention_bwd_rows_kernel<Ty><<<grid, ATT_WARPS * 32, smem, (cudaStream_t)stream>>>(T, Tp, heads, (const Ty *)qkv, (const Ty *)probs, \
        (const Ty *)dout, scale, (Ty *)dqkv, (Ty *)ds_scratch, rpc); \
    attention_bwd_cols_kernel<Ty><<<grid, ATT_WARPS * 32, smem, (cudaStream_t)stream>>>(T, Tp, heads, (const Ty *)qkv, (const Ty *)probs, \
        (const Ty *)ds_scratch, (const Ty *)dout, scale, (Ty *)dqkv, rpc); }
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_gelu_fwd(int dtype, int64_t n, const void *u, void *h, void *stream) {
  KS_CHECK_ARG(n > 0 && u && h);
  if (n % 8 || !al16(u) || !al16(h)) return KS_EUNSUPPORTED;
  const int grid = grid_for(n / 8, 256 * 4);
#define CALL(Ty) gelu_fwd_kernel<Ty><<<grid, 256, 0, (cudaStream_t)stream>>>(n / 8, (const Ty *)u, (Ty *)h)
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_gelu_bwd(int dtype, int64_t n, const void *u, const void *dh, void *du, void *stream) {
  KS_CHECK_ARG(n > 0 && u && dh && du);
  if (n % 8 || !al16(u) || !al16(dh) || !al16(du)) return KS_EUNSUPPORTED;
  const int grid = grid_for(n / 8, 256 * 4);
#define CALL(Ty) gelu_bwd_kernel<Ty><<<grid, 256, 0, (cudaStream_t)stream>>>(n / 8, (const Ty *)u, (const Ty *)dh, (Ty *)du)
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_bilinear_up_fwd(int dtype, int B, int G, int Tp, int row0, int Cs, int K, int Ho, int Wo, const void *src, float *dst,
                                  void *stream) {
  KS_CHECK_ARG(B > 0 && G > 0 && K > 0 && K <= Cs && row0 >= 0 && row0 + G * G <= Tp && Ho > 0 && Wo > 0 && src && dst);
  const int grid = grid_for((long long)B * Ho * Wo, 256);
#define CALL(Ty) bilinear_up_fwd_kernel<Ty><<<grid, 256, 0, (cudaStream_t)stream>>>(B, G, Tp, row0, Cs, K, Ho, Wo, (const Ty *)src, dst)
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_bilinear_up_bwd(int dtype, int B, int G, int Tp, int row0, int Cs, int K, int Ho, int Wo, const float *ddst, void *dsrc,
                                  void *stream) {
  KS_CHECK_ARG(B > 0 && G > 0 && K > 0 && K <= Cs && row0 >= 0 && row0 + G * G <= Tp && Ho > 0 && Wo > 0 && ddst && dsrc);
  if (K > 32) return KS_EUNSUPPORTED;
  const int grid = grid_for((long long)B * Tp, 8, 16);
#define CALL(Ty) bilinear_up_bwd_kernel<Ty><<<grid, 256, 0, (cudaStream_t)stream>>>(B, G, Tp, row0, Cs, K, Ho, Wo, ddst, (Ty *)dsrc)
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_adaptive_avgpool_fwd(int dtype, int N, int H, int W, int S, const ks_view_t *src, void *dst, void *stream) {
  KS_CHECK_ARG(N > 0 && H > 0 && W > 0 && S > 0 && S <= H && S <= W && src && src->ptr && dst);
  const int grid = grid_for((long long)N * S * S * src->C, 256);
#define CALL(Ty) adaptive_pool_fwd_kernel<Ty><<<grid, 256, 0, (cudaStream_t)stream>>>(to_view(*src), N, H, W, S, (Ty *)dst)
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_adaptive_avgpool_bwd(int dtype, int N, int H, int W, int S, const void *ddst, const ks_view_t *dsrc, int accumulate, void *stream) {
  KS_CHECK_ARG(N > 0 && H > 0 && W > 0 && S > 0 && S <= H && S <= W && dsrc && dsrc->ptr && ddst);
  const int es = (dtype == KS_F32) ? 4 : 2;
  const bool vec = dsrc->C % 8 == 0 && al16(dsrc->ptr) && al16(ddst) && (dsrc->sn * es) % 16 == 0 && (dsrc->sh * es) % 16 == 0 && (dsrc->sw * es) % 16 == 0;
  const int grid = grid_for((long long)N * H * W * (vec ? dsrc->C / 8 : dsrc->C), 256);
#define CALL(Ty) { if (vec) adaptive_pool_bwd_vec_kernel<Ty><<<grid, 256, 0, (cudaStream_t)stream>>>(to_view(*dsrc), N, H, W, S, (const Ty *)ddst, accumulate); \
    else adaptive_pool_bwd_kernel<Ty><<<grid, 256, 0, (cudaStream_t)stream>>>(to_view(*dsrc), N, H, W, S, (const Ty *)ddst, accumulate); }
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}
