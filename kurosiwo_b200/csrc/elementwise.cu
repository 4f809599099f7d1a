// Bandwidth-bound NHWC passes: layout/precision plumbing, train-mode BatchNorm
// (+ReLU, +residual, +2x2 max-pool) forward/backward, pooling backward, bias
// gradients, and the flat Adam step.  All kernels work on strided NHWC views
// (channel stride 1) so that concat slots and ConvTranspose phases need no copies.
//
// Reference call sites: models/snunet.py:20-29 (conv_block_nested.forward),
// :73 (MaxPool2d), training/change_detection_trainer.py:52-54,174-180 (Adam step).
#include "common.cuh"

namespace ks {

// ------------------------------------------------------------------------------------------
// permute + cast
// ------------------------------------------------------------------------------------------
template <typename TS, typename TD>
__global__ void permute_cast_kernel(const TS *__restrict__ src, TD *__restrict__ dst,
                                    int d1, int d2, int d3, long long total,
                                    long long s0, long long s1, long long s2, long long s3, int accumulate) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int i3 = (int)(r % d3); r /= d3;
    const int i2 = (int)(r % d2); r /= d2;
    const int i1 = (int)(r % d1); r /= d1;
    const long long i0 = r;
    float v = Cvt<TS>::ld(src + i0 * s0 + i1 * s1 + i2 * s2 + i3 * s3);
    if (accumulate) v += Cvt<TD>::ld(dst + i);
    Cvt<TD>::st(dst + i, v);
  }
}

// Batched variant: one launch for a whole table of permute/cast jobs (all weight packs of a step, or all gradient
// unpacks).  `chunks[i] = {job, first element}`; every block processes up to 4096 consecutive dst elements of one job.
__global__ void __launch_bounds__(256)
permute_cast_batched_kernel(const ks_permute_job_t *__restrict__ jobs, const int2 *__restrict__ chunks) {
  const int2 ch = chunks[blockIdx.x];
  const ks_permute_job_t j = jobs[ch.x];
  if (j.dst_strided == 2) {
    // 2-D transpose job (dst[a][b] = src[a + b*s1], contiguous dst [d0][d1]): chunk = one 64 x 64 tile through shared memory, so
    // that BOTH the reads (along a) and the writes (along b) are coalesced - the packed data-gradient copies of every Linear weight
    __shared__ float tile[64][65];
    const int A = (int)(j.total / j.d1), B = j.d1;
    const int tiles_b = (B + 63) / 64;
    const int a0 = (ch.y / tiles_b) * 64, b0 = (ch.y % tiles_b) * 64;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    for (int k = ty; k < 64; k += 4) {
      const int a = a0 + tx, b = b0 + k;
      if (a < A && b < B) {
        const long long so = (long long)a * j.s0 + (long long)b * j.s1;
        tile[k][tx] = (j.src_dtype == KS_F32) ? reinterpret_cast<const float *>(j.src)[so]
                                              : __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(j.src)[so]);
      }
    }
    __syncthreads();
    for (int k = ty; k < 64; k += 4) {
      const int a = a0 + k, b = b0 + tx;
      if (a < A && b < B) {
        float v = tile[tx][k];
        if (j.scale != 0.f) v *= j.scale;
        const long long di = (long long)a * B + b;
        if (j.dst_dtype == KS_F32) reinterpret_cast<float *>(j.dst)[di] = v;
        else reinterpret_cast<__nv_bfloat16 *>(j.dst)[di] = __float2bfloat16_rn(v);
      }
    }
    return;
  }
  const long long end = min((long long)j.total, (long long)ch.y + 4096);
  if (j.d1 == 1 && j.d2 == 1 && j.d3 == 1 && j.s0 == 1 && !j.dst_strided && !j.accumulate && j.src_dtype == KS_F32 && j.dst_dtype == KS_BF16 &&
      (ch.y & 7) == 0 && (((uintptr_t)j.src) & 31) == 0 && (((uintptr_t)j.dst) & 15) == 0) {
    // contiguous fp32 -> bf16 cast (the forward copies of the Linear weights): 8 elements per thread and pass
    const float *sp = reinterpret_cast<const float *>(j.src);
    __nv_bfloat16 *dp = reinterpret_cast<__nv_bfloat16 *>(j.dst);
    for (long long i = ch.y + 8LL * threadIdx.x; i < end; i += 8LL * blockDim.x) {
      if (i + 8 <= end) {
        float f[8];
        ld8(sp + i, f);
        if (j.scale != 0.f) {
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] *= j.scale;
        }
        st8(dp + i, f);
      } else {
        for (long long q = i; q < end; ++q) dp[q] = __float2bfloat16_rn(j.scale != 0.f ? sp[q] * j.scale : sp[q]);
      }
    }
    return;
  }
  if (j.total < (1ll << 31) && !j.dst_strided && !j.accumulate && j.src_dtype == KS_F32 && j.d3 == 1) {
    // the conv-weight packs / gradient unpacks of a step ([t][o][i] <-> OIHW): 32-bit index arithmetic (the 64-bit divisions of the
    // generic loop below were ~100 instructions per element: 0.21 ms per SNUNet step for 36 M elements)
    const float *sp = reinterpret_cast<const float *>(j.src);
    const unsigned int d1 = (unsigned int)j.d1, d2 = (unsigned int)j.d2;
    const int s0 = (int)j.s0, s1 = (int)j.s1, s2 = (int)j.s2;
    for (unsigned int i = (unsigned int)ch.y + threadIdx.x; i < (unsigned int)end; i += blockDim.x) {
      unsigned int r = i;
      const unsigned int i2 = r % d2; r /= d2;
      const unsigned int i1 = r % d1; r /= d1;
      float v = sp[(long long)((int)r * s0 + (int)i1 * s1 + (int)i2 * s2)];
      if (j.scale != 0.f) v *= j.scale;
      if (j.dst_dtype == KS_F32) reinterpret_cast<float *>(j.dst)[i] = v;
      else reinterpret_cast<__nv_bfloat16 *>(j.dst)[i] = __float2bfloat16_rn(v);
    }
    return;
  }
  for (long long i = ch.y + threadIdx.x; i < end; i += blockDim.x) {
    long long r = i;
    const int i3 = (int)(r % j.d3); r /= j.d3;
    const int i2 = (int)(r % j.d2); r /= j.d2;
    const int i1 = (int)(r % j.d1); r /= j.d1;
    const long long so = r * j.s0 + i1 * j.s1 + i2 * j.s2 + i3 * j.s3;
    float v = (j.src_dtype == KS_F32) ? reinterpret_cast<const float *>(j.src)[so]
                                      : __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(j.src)[so]);
    if (j.scale != 0.f) v *= j.scale;
    const long long di = j.dst_strided ? (r * j.t0 + i1 * j.t1 + i2 * j.t2 + i3 * j.t3) : i;
    if (j.dst_dtype == KS_F32) {
      float *d = reinterpret_cast<float *>(j.dst) + di;
      *d = j.accumulate ? (*d + v) : v;
    } else {
      __nv_bfloat16 *d = reinterpret_cast<__nv_bfloat16 *>(j.dst) + di;
      *d = __float2bfloat16_rn(j.accumulate ? (__bfloat162float(*d) + v) : v);
    }
  }
}

// ------------------------------------------------------------------------------------------
// generic helpers for NHWC view kernels
// ------------------------------------------------------------------------------------------
template <typename T, int V> struct VecIO {
  static __device__ __forceinline__ void ld(const T *p, float (&f)[V]) {
    if constexpr (V == 8) ld8(p, f); else {
#pragma unroll
      for (int i = 0; i < V; ++i) f[i] = Cvt<T>::ld(p + i);
    }
  }
  static __device__ __forceinline__ void st(T *p, const float (&f)[V]) {
    if constexpr (V == 8) st8(p, f); else {
#pragma unroll
      for (int i = 0; i < V; ++i) Cvt<T>::st(p + i, f[i]);
    }
  }
};

template <typename T>
__device__ __forceinline__ T *vaddr(const View &v, int n, int h, int w, int c) {
  return reinterpret_cast<T *>(v.ptr) + ((long long)n * v.sn + (long long)h * v.sh + (long long)w * v.sw + c);
}

static inline bool view_vec8_ok(const ks_view_t &v, int esize) {
  return (v.C % 8 == 0) && (((uintptr_t)v.ptr % 16) == 0) && ((v.sn * esize) % 16 == 0) &&
         ((v.sh * esize) % 16 == 0) && ((v.sw * esize) % 16 == 0);
}

// Block-level per-channel reduction of R running sums held as acc[R][V] per thread.
// Thread layout: tx = threadIdx.x % CV (channel vector), ty = threadIdx.x / CV (pixel lane).
template <int R, int V>
__device__ __forceinline__ void block_channel_reduce(float (&acc)[R][V], int CV, int rows, int C,
                                                     double *out /* [R][C] */, float *smem) {
  const int tx = threadIdx.x % CV, ty = threadIdx.x / CV;
  const int CW = CV * V;  // == C
  if (ty < rows) {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int i = 0; i < V; ++i) smem[(r * rows + ty) * CW + tx * V + i] = acc[r][i];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < R * CW; j += blockDim.x) {
    const int r = j / CW, c = j % CW;
    double s = 0.0;
    for (int y = 0; y < rows; ++y) s += (double)smem[(r * rows + y) * CW + c];
    atomicAdd(out + (size_t)r * C + c, s);
  }
}

__device__ __forceinline__ long long pix_off(const View &v, long long p, int H, int W, bool flat) {
  if (flat) return p * v.sw;
  const int w = (int)(p % W); const long long r = p / W; const int h = (int)(r % H); const long long n = r / H;
  return n * v.sn + (long long)h * v.sh + (long long)w * v.sw;
}
static inline bool view_flat(const ks_view_t &v, int H, int W) { return v.sh == (int64_t)W * v.sw && v.sn == (int64_t)H * v.sh; }

// Raw register image of V elements: loads are issued for several pixels BEFORE any use so that each thread keeps
// 4 (bf16) / 2 (fp32) x #tensors 16-byte requests in flight (ncu: 1-2 requests per thread gave ~50 % of HBM peak).
template <typename T, int V> struct Raw {
  T x[V];
  __device__ __forceinline__ void ld(const T *p) {
#pragma unroll
    for (int i = 0; i < V; ++i) x[i] = p[i];
  }
  __device__ __forceinline__ void get(float (&f)[V]) const {
#pragma unroll
    for (int i = 0; i < V; ++i) { T t = x[i]; f[i] = Cvt<T>::ld(&t); }
  }
};
template <> struct Raw<__nv_bfloat16, 8> {
  uint4 u;
  __device__ __forceinline__ void ld(const __nv_bfloat16 *p) { u = *reinterpret_cast<const uint4 *>(p); }
  __device__ __forceinline__ void get(float (&f)[8]) const {
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
};
template <> struct Raw<float, 8> {
  float4 a, b;
  __device__ __forceinline__ void ld(const float *p) { a = *reinterpret_cast<const float4 *>(p); b = *reinterpret_cast<const float4 *>(p + 4); }
  __device__ __forceinline__ void get(float (&f)[8]) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};
template <typename T> struct Unr { static constexpr int value = (sizeof(T) == 2) ? 4 : 2; };

// ------------------------------------------------------------------------------------------
// BatchNorm statistics: sums[0][c] += sum x, sums[1][c] += sum x^2
// ------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(256)
bn_stats_kernel(View x, int N, int H, int W, double *sums) {
  extern __shared__ float smem[];
  const int C = x.C, CV = C / V, rows = blockDim.x / CV;
  const int tx = threadIdx.x % CV, ty = threadIdx.x / CV;
  float acc[2][V];
#pragma unroll
  for (int i = 0; i < V; ++i) { acc[0][i] = 0.f; acc[1][i] = 0.f; }
  const long long npix = (long long)N * H * W;
  if (ty < rows) {
    for (long long p = (long long)blockIdx.x * rows + ty; p < npix; p += (long long)gridDim.x * rows) {
      const int w = (int)(p % W); long long r = p / W; const int h = (int)(r % H); const int n = (int)(r / H);
      float f[V]; VecIO<T, V>::ld(vaddr<T>(x, n, h, w, tx * V), f);
#pragma unroll
      for (int i = 0; i < V; ++i) { acc[0][i] += f[i]; acc[1][i] += f[i] * f[i]; }
    }
  }
  block_channel_reduce<2, V>(acc, CV, rows, C, sums, smem);
}

__global__ void bn_finalize_kernel(int C, double count, const double *__restrict__ sums,
                                   const float *__restrict__ gamma, const float *__restrict__ beta,
                                   float eps, float momentum, float *running_mean, float *running_var,
                                   float *scale, float *shift, float *mean_out, float *rstd_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = sums[c] / count;
  double var = sums[C + c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float sc = g * rstd;
  scale[c] = sc; shift[c] = b - (float)mean * sc;
  mean_out[c] = (float)mean; rstd_out[c] = rstd;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// ------------------------------------------------------------------------------------------
// Pixel-lane layout shared by the BN passes: tx = channel vector (fixed per thread, so per-channel
// coefficients live in registers), ty = pixel lane; a warp touches (32/CV) whole pixels x C channels.
// `flat` views (pixel stride uniform: sh == W*sw, sn == H*sh) skip the n/h/w decomposition.
// ------------------------------------------------------------------------------------------

// out = relu(y*scale+shift (+res)), optional fused 2x2 max-pool output
// Bytes in flight decide these passes (B200: ~32 KB per SM in flight measured 4.4 TB/s, ~96 KB 6.2 TB/s): the pixel unroll is chosen
// per instantiation so that every thread keeps 8-12 independent 16-byte loads outstanding whatever the number of input tensors.
template <typename T, int V, bool POOL, bool HAS_RES>
__global__ void __launch_bounds__(256, 2)
bn_act_kernel(View y, View res, View out, View pool, int N, int H, int W,
              const float *__restrict__ scale, const float *__restrict__ shift, int relu, bool flat) {
  constexpr bool has_res = HAS_RES;
  const int CV = y.C / V, rows = blockDim.x / CV;
  const int tx = threadIdx.x % CV, ty = threadIdx.x / CV;
  if (ty >= rows) return;
  const int c = tx * V;
  float sc[V], sh[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { sc[k] = __ldg(scale + c + k); sh[k] = __ldg(shift + c + k); }
  const T *yp = reinterpret_cast<const T *>(y.ptr) + c;
  const T *rp = reinterpret_cast<const T *>(res.ptr) + c;
  T *op = reinterpret_cast<T *>(out.ptr) + c;
  auto finish = [&](const Raw<T, V> &ry, const Raw<T, V> &rr, long long p, float (&o)[V]) {
    float f[V], r2[V];
    ry.get(f);
    if (has_res) rr.get(r2);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float v = fmaf(f[k], sc[k], sh[k]);
      if (has_res) v += r2[k];
      if (relu) v = fmaxf(v, 0.f);
      o[k] = v;
    }
    VecIO<T, V>::st(op + pix_off(out, p, H, W, flat), o);
  };
  if constexpr (POOL) {
    const int HP = H / 2, WP = W / 2;
    const long long npool = (long long)N * HP * WP;
    T *pp = reinterpret_cast<T *>(pool.ptr) + c;
    for (long long q = (long long)blockIdx.x * rows + ty; q < npool; q += (long long)gridDim.x * rows) {
      const int wp = (int)(q % WP); const long long r = q / WP; const int hp = (int)(r % HP); const long long n = r / HP;
      const long long p00 = (n * H + 2 * hp) * W + 2 * wp;
      Raw<T, V> ry[4], rr[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long p = p00 + (i >> 1) * W + (i & 1);
        ry[i].ld(yp + pix_off(y, p, H, W, flat));
        if (has_res) rr[i].ld(rp + pix_off(res, p, H, W, flat));
      }
      float mx[V], o[V];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        finish(ry[i], rr[i], p00 + (i >> 1) * W + (i & 1), o);
#pragma unroll
        for (int k = 0; k < V; ++k) { const float os = round_as<T>(o[k]); mx[k] = (i == 0) ? os : fmaxf(mx[k], os); }
      }
      VecIO<T, V>::st(pp + (n * pool.sn + (long long)hp * pool.sh + (long long)wp * pool.sw), mx);
    }
  } else {
    constexpr int U = HAS_RES ? (Unr<T>::value * 3) / 2 : Unr<T>::value * 2;
    const long long npix = (long long)N * H * W, stride = (long long)gridDim.x * rows;
    for (long long p0 = (long long)blockIdx.x * rows + ty; p0 < npix; p0 += stride * U) {
      Raw<T, V> ry[U], rr[HAS_RES ? U : 1];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long p = p0 + u * stride;
        if (p < npix) {
          ry[u].ld(yp + pix_off(y, p, H, W, flat));
          if (has_res) rr[HAS_RES ? u : 0].ld(rp + pix_off(res, p, H, W, flat));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long p = p0 + u * stride;
        float o[V];
        if (p < npix) finish(ry[u], rr[HAS_RES ? u : 0], p, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// BN backward, pass 1: g = dout * mask;  sums[0][c] += sum g,  sums[1][c] += sum g*xhat.
//   MASK_OUT : mask = (out > 0); the masked gradient is written back over dout (later passes re-use it)
//   !MASK_OUT: mask = (y*scale+shift > 0)  (== (relu output > 0), recomputed instead of re-read)
// ------------------------------------------------------------------------------------------
template <typename T, int V, bool MASK_OUT>
__global__ void __launch_bounds__(256, 2)
bn_bwd_reduce_kernel(View dout, View out, View y, int N, int H, int W, const float *__restrict__ scale,
                     const float *__restrict__ shift, const float *__restrict__ mean, const float *__restrict__ rstd,
                     double *sums, bool flat) {
  extern __shared__ float smem[];
  const int C = y.C, CV = C / V, rows = blockDim.x / CV;
  const int tx = threadIdx.x % CV, ty = threadIdx.x / CV;
  float acc[2][V];
#pragma unroll
  for (int i = 0; i < V; ++i) { acc[0][i] = 0.f; acc[1][i] = 0.f; }
  const long long npix = (long long)N * H * W;
  if (ty < rows) {
    const int c = tx * V;
    float mu[V], rs[V], sc[V], sh[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      mu[k] = mean[c + k]; rs[k] = rstd[c + k];
      sc[k] = MASK_OUT ? 0.f : scale[c + k]; sh[k] = MASK_OUT ? 0.f : shift[c + k];
    }
    T *gp = reinterpret_cast<T *>(dout.ptr) + c;
    const T *op = reinterpret_cast<const T *>(out.ptr) + c;
    const T *yp = reinterpret_cast<const T *>(y.ptr) + c;
    constexpr int U = MASK_OUT ? Unr<T>::value : (Unr<T>::value * 3) / 2;     // 12 loads in flight per thread either way
    const long long stride = (long long)gridDim.x * rows;
    for (long long p0 = (long long)blockIdx.x * rows + ty; p0 < npix; p0 += stride * U) {
      Raw<T, V> rg[U], ro[MASK_OUT ? U : 1], rf[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long p = p0 + u * stride;
        if (p < npix) {
          rg[u].ld(gp + pix_off(dout, p, H, W, flat));
          rf[u].ld(yp + pix_off(y, p, H, W, flat));
          if (MASK_OUT) ro[MASK_OUT ? u : 0].ld(op + pix_off(out, p, H, W, flat));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long p = p0 + u * stride;
        if (p < npix) {
          float g[V], o[V], f[V];
          rg[u].get(g); rf[u].get(f);
          if (MASK_OUT) ro[MASK_OUT ? u : 0].get(o);
#pragma unroll
          for (int k = 0; k < V; ++k) {
            const bool on = MASK_OUT ? (o[k] > 0.f) : (fmaf(f[k], sc[k], sh[k]) > 0.f);
            const float gk = on ? g[k] : 0.f;
            g[k] = gk;
            acc[0][k] += gk;
            acc[1][k] += gk * ((f[k] - mu[k]) * rs[k]);
          }
          if (MASK_OUT) VecIO<T, V>::st(gp + pix_off(dout, p, H, W, flat), g);
        }
      }
    }
  }
  block_channel_reduce<2, V>(acc, CV, rows, C, sums, smem);
}

// pass 1 for encoder outputs: as MASK_OUT above, with the 2x2 max-pool backward folded in.  Each thread owns whole
// pooling windows: g = (dout + [pixel is the window's first maximum] * dpool) * (out > 0), written back over dout.
template <typename T, int V>
__global__ void __launch_bounds__(256, 2)
bn_bwd_reduce_pool_kernel(View dout, View out, View y, View dpool, int N, int H, int W,
                          const float *__restrict__ mean, const float *__restrict__ rstd, double *sums, bool flat) {
  extern __shared__ float smem[];
  const int C = y.C, CV = C / V, rows = blockDim.x / CV;
  const int tx = threadIdx.x % CV, ty = threadIdx.x / CV;
  float acc[2][V];
#pragma unroll
  for (int i = 0; i < V; ++i) { acc[0][i] = 0.f; acc[1][i] = 0.f; }
  if (ty < rows) {
    const int c = tx * V;
    float mu[V], rs[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { mu[k] = mean[c + k]; rs[k] = rstd[c + k]; }
    T *gp = reinterpret_cast<T *>(dout.ptr) + c;
    const T *op = reinterpret_cast<const T *>(out.ptr) + c;
    const T *yp = reinterpret_cast<const T *>(y.ptr) + c;
    const T *pp = reinterpret_cast<const T *>(dpool.ptr) + c;
    const int HP = H / 2, WP = W / 2;
    const long long npool = (long long)N * HP * WP;
    for (long long q = (long long)blockIdx.x * rows + ty; q < npool; q += (long long)gridDim.x * rows) {
      const int wp = (int)(q % WP); const long long r = q / WP; const int hp = (int)(r % HP); const long long n = r / HP;
      const long long p00 = (n * H + 2 * hp) * W + 2 * wp;
      Raw<T, V> rg[4], ro[4], rf[4], rp;
      rp.ld(pp + (n * dpool.sn + (long long)hp * dpool.sh + (long long)wp * dpool.sw));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long p = p00 + (i >> 1) * W + (i & 1);
        rg[i].ld(gp + pix_off(dout, p, H, W, flat));
        ro[i].ld(op + pix_off(out, p, H, W, flat));
        rf[i].ld(yp + pix_off(y, p, H, W, flat));
      }
      float o[4][V], dp[V];
      rp.get(dp);
#pragma unroll
      for (int i = 0; i < 4; ++i) ro[i].get(o[i]);
      int am[V];
#pragma unroll
      for (int k = 0; k < V; ++k) {
        int a = 0; float best = o[0][k];
#pragma unroll
        for (int i = 1; i < 4; ++i) if (o[i][k] > best) { best = o[i][k]; a = i; }
        am[k] = a;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float g[V], f[V];
        rg[i].get(g); rf[i].get(f);
#pragma unroll
        for (int k = 0; k < V; ++k) {
          float gk = g[k] + ((am[k] == i) ? dp[k] : 0.f);
          gk = (o[i][k] > 0.f) ? gk : 0.f;
          g[k] = gk;
          acc[0][k] += gk;
          acc[1][k] += gk * ((f[k] - mu[k]) * rs[k]);
        }
        VecIO<T, V>::st(gp + pix_off(dout, p00 + (i >> 1) * W + (i & 1), H, W, flat), g);
      }
    }
  }
  block_channel_reduce<2, V>(acc, CV, rows, C, sums, smem);
}

// pass 2: dy = gamma*rstd*(g - sum_g/M - xhat*sum_gx/M) (+ add) = a*g + (k1*y + k0) (+ add);  `premasked`: g already masked by pass 1
template <typename T, int V, bool HAS_ADD>
__global__ void __launch_bounds__(256, 2)
bn_bwd_apply_kernel(View gin, View y, View add, View dy, int N, int H, int W, int premasked,
                    const float *__restrict__ scale, const float *__restrict__ shift,
                    const float *__restrict__ mean, const float *__restrict__ rstd, const float *__restrict__ gamma,
                    const double *__restrict__ sums, double count, float *dgamma, float *dbeta, float *dsum_out,
                    int accumulate, bool flat) {
  const int C = y.C, CV = C / V, rows = blockDim.x / CV;
  const int tx = threadIdx.x % CV, ty = threadIdx.x / CV;
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)sums[C + c];
      if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)sums[c];
      if (dsum_out) dsum_out[c] = (accumulate ? dsum_out[c] : 0.f) + (float)sums[c];
    }
  }
  if (ty >= rows) return;
  const int c = tx * V;
  const float invM = (float)(1.0 / count);
  float a[V], k1[V], k0[V], sc[V], sh[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const float mu = mean[c + k], rs = rstd[c + k];
    a[k] = (gamma ? gamma[c + k] : 1.f) * rs;
    const float b = (float)sums[c + k] * invM, c2 = (float)sums[C + c + k] * invM;
    k1[k] = -a[k] * c2 * rs;
    k0[k] = -a[k] * b - k1[k] * mu;
    sc[k] = premasked ? 0.f : scale[c + k]; sh[k] = premasked ? 0.f : shift[c + k];
  }
  const T *gp = reinterpret_cast<const T *>(gin.ptr) + c;
  const T *yp = reinterpret_cast<const T *>(y.ptr) + c;
  const T *ap = reinterpret_cast<const T *>(add.ptr) + c;
  T *dp = reinterpret_cast<T *>(dy.ptr) + c;
  const long long npix = (long long)N * H * W, stride = (long long)gridDim.x * rows;
  constexpr bool has_add = HAS_ADD;
  constexpr int U = HAS_ADD ? Unr<T>::value : (Unr<T>::value * 3) / 2;
  for (long long p0 = (long long)blockIdx.x * rows + ty; p0 < npix; p0 += stride * U) {
    Raw<T, V> rg[U], rf[U], ra[HAS_ADD ? U : 1];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = p0 + u * stride;
      if (p < npix) {
        rg[u].ld(gp + pix_off(gin, p, H, W, flat));
        rf[u].ld(yp + pix_off(y, p, H, W, flat));
        if (has_add) ra[HAS_ADD ? u : 0].ld(ap + pix_off(add, p, H, W, flat));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = p0 + u * stride;
      if (p < npix) {
        float g[V], f[V], ad[V], d[V];
        rg[u].get(g); rf[u].get(f);
        if (has_add) ra[HAS_ADD ? u : 0].get(ad);
#pragma unroll
        for (int k = 0; k < V; ++k) {
          const bool on = premasked ? true : (fmaf(f[k], sc[k], sh[k]) > 0.f);
          float v = fmaf(a[k], on ? g[k] : 0.f, fmaf(k1[k], f[k], k0[k]));
          if (has_add) v += ad[k];
          d[k] = v;
        }
        VecIO<T, V>::st(dp + pix_off(dy, p, H, W, flat), d);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// 2x2/s2 max-pool backward (first maximum wins, as aten max_pool2d_with_indices)
// ------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(View x, View dpool, View dx, int N, int H, int W, int accumulate) {
  const int CV = x.C / V;
  const long long total = (long long)N * H * W * CV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV); long long r = i / CV;
    const int w = (int)(r % W); r /= W; const int h = (int)(r % H); const int n = (int)(r / H);
    const int c = cv * V;
    float xv[4][V], g[V];
#pragma unroll
    for (int q = 0; q < 4; ++q) VecIO<T, V>::ld(vaddr<T>(x, n, 2 * h + (q >> 1), 2 * w + (q & 1), c), xv[q]);
    VecIO<T, V>::ld(vaddr<T>(dpool, n, h, w, c), g);
    int am[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      int a = 0; float best = xv[0][k];
#pragma unroll
      for (int q = 1; q < 4; ++q) if (xv[q][k] > best) { best = xv[q][k]; a = q; }
      am[k] = a;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      T *p = vaddr<T>(dx, n, 2 * h + (q >> 1), 2 * w + (q & 1), c);
      float d[V];
      if (accumulate) VecIO<T, V>::ld(p, d); else {
#pragma unroll
        for (int k = 0; k < V; ++k) d[k] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < V; ++k) if (am[k] == q) d[k] += g[k];
      VecIO<T, V>::st(p, d);
    }
  }
}

// ------------------------------------------------------------------------------------------
// channel sum (bias gradients)
// ------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(256)
channel_sum_kernel(View x, int N, int H, int W, float *out, bool flat) {
  extern __shared__ float smem[];
  const int C = x.C, CV = C / V, rows = blockDim.x / CV;
  const int tx = threadIdx.x % CV, ty = threadIdx.x / CV;
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  const long long npix = (long long)N * H * W;
  if (ty < rows) {
    const T *xp = reinterpret_cast<const T *>(x.ptr) + tx * V;
    constexpr int U = 2 * Unr<T>::value;
    const long long stride = (long long)gridDim.x * rows;
    for (long long p0 = (long long)blockIdx.x * rows + ty; p0 < npix; p0 += stride * U) {
      Raw<T, V> r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) if (p0 + u * stride < npix) r[u].ld(xp + pix_off(x, p0 + u * stride, H, W, flat));
#pragma unroll
      for (int u = 0; u < U; ++u) if (p0 + u * stride < npix) {
        float f[V]; r[u].get(f);
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] += f[i];
      }
    }
    for (int i = 0; i < V; ++i) smem[ty * C + tx * V + i] = acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int yy = 0; yy < rows; ++yy) s += smem[yy * C + c];
    atomicAdd(out + c, s);
  }
}

// ------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam semantics, L2 weight decay folded into the gradient) and, with DECOUPLED, AdamW
// (torch.optim.AdamW: p *= 1 - lr*wd before the Adam update; the gradient carries no decay term)
// ------------------------------------------------------------------------------------------
template <bool DECOUPLED>
__global__ void __launch_bounds__(256)
adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
            long long n, float lr, float b1, float b2, float eps, float wd, float gscale, const int *step_ptr) {
  const int t = *step_ptr + 1;
  const double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  const float decay = 1.f - lr * wd;
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4 *>(p)[i], gg = reinterpret_cast<const float4 *>(g)[i];
    float4 mm = reinterpret_cast<float4 *>(m)[i], vv = reinterpret_cast<float4 *>(v)[i];
    float *P = &pp.x, *G = &gg.x, *M = &mm.x, *Vv = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gk = G[k] * gscale;
      if (DECOUPLED) P[k] *= decay; else gk += wd * P[k];
      M[k] = b1 * M[k] + (1.f - b1) * gk;
      Vv[k] = b2 * Vv[k] + (1.f - b2) * gk * gk;
      P[k] -= step_size * (M[k] / (sqrtf(Vv[k]) * inv_bc2_sqrt + eps));
    }
    reinterpret_cast<float4 *>(p)[i] = pp; reinterpret_cast<float4 *>(m)[i] = mm; reinterpret_cast<float4 *>(v)[i] = vv;
  }
  if (blockIdx.x == 0) {
    for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
      float gk = g[i] * gscale;
      float pi = p[i];
      if (DECOUPLED) pi *= decay; else gk += wd * pi;
      m[i] = b1 * m[i] + (1.f - b1) * gk;
      v[i] = b2 * v[i] + (1.f - b2) * gk * gk;
      p[i] = pi - step_size * (m[i] / (sqrtf(v[i]) * inv_bc2_sqrt + eps));
    }
  }
}
__global__ void incr_kernel(int *p) { *p += 1; }

static inline int ew_grid(long long total, int block) {
  long long b = (total + block - 1) / block;
  const long long cap = (long long)kNumSMs * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <typename T, int V>
static int launch_reduce_cfg(int C, long long npix, int &grid, size_t &smem, int R) {
  const int CV = C / V;
  if (CV < 1 || CV > 256) return KS_EUNSUPPORTED;
  const int rows = 256 / CV;
  long long g = (npix + (long long)rows * 8 - 1) / ((long long)rows * 8);
  const long long cap = (long long)kNumSMs * 8;
  grid = (int)(g < 1 ? 1 : (g > cap ? cap : g));
  smem = (size_t)R * rows * C * sizeof(float);
  return KS_OK;
}

}  // namespace ks

using namespace ks;

#define KS_DISPATCH_TV(dtype, vec_ok, CALL)                                                  \
  do {                                                                                       \
    if ((dtype) == KS_F32) { if (vec_ok) { CALL(float, 8); } else { CALL(float, 1); } }      \
    else if ((dtype) == KS_BF16) { if (vec_ok) { CALL(__nv_bfloat16, 8); } else { CALL(__nv_bfloat16, 1); } } \
    else return KS_EINVAL;                                                                   \
  } while (0)

static inline int esize_of(int dtype) { return dtype == KS_F32 ? 4 : 2; }

extern "C" int ks_version(void) { return 100; }

extern "C" const char *ks_error_string(int code) {
  switch (code) {
    case KS_OK: return "ok";
    case KS_EINVAL: return "invalid argument";
    case KS_EUNSUPPORTED: return "unsupported shape for the requested implementation";
    case KS_EDRIVER: return "driver entry point unavailable";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown";
  }
}

extern "C" int ks_permute_cast(int src_dtype, const void *src, int dst_dtype, void *dst,
                               int d0, int d1, int d2, int d3,
                               int64_t s0, int64_t s1, int64_t s2, int64_t s3, int accumulate, void *stream) {
  KS_CHECK_ARG(src && dst && d0 > 0 && d1 > 0 && d2 > 0 && d3 > 0);
  const long long total = (long long)d0 * d1 * d2 * d3;
  const int grid = ew_grid(total, 256);
  cudaStream_t st = (cudaStream_t)stream;
#define PC(TS, TD) permute_cast_kernel<TS, TD><<<grid, 256, 0, st>>>((const TS *)src, (TD *)dst, d1, d2, d3, total, s0, s1, s2, s3, accumulate)
  if (src_dtype == KS_F32 && dst_dtype == KS_F32) PC(float, float);
  else if (src_dtype == KS_F32 && dst_dtype == KS_BF16) PC(float, __nv_bfloat16);
  else if (src_dtype == KS_BF16 && dst_dtype == KS_F32) PC(__nv_bfloat16, float);
  else if (src_dtype == KS_BF16 && dst_dtype == KS_BF16) PC(__nv_bfloat16, __nv_bfloat16);
  else return KS_EINVAL;
#undef PC
  KS_LAUNCH_RET();
}

extern "C" int ks_permute_cast_batched(const ks_permute_job_t *jobs_dev, const int32_t *chunks_dev, int n_chunks, void *stream) {
  KS_CHECK_ARG(jobs_dev && chunks_dev && n_chunks > 0);
  permute_cast_batched_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(jobs_dev, reinterpret_cast<const int2 *>(chunks_dev));
  KS_LAUNCH_RET();
}

extern "C" int ks_bn_stats(int dtype, int N, int H, int W, const ks_view_t *x, double *sums, void *stream) {
  KS_CHECK_ARG(x && x->ptr && sums && N > 0 && H > 0 && W > 0);
  const bool vec = view_vec8_ok(*x, esize_of(dtype));
  const long long npix = (long long)N * H * W;
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(T, V) { int grid; size_t smem; int rc = launch_reduce_cfg<T, V>(x->C, npix, grid, smem, 2); if (rc) return rc; \
    bn_stats_kernel<T, V><<<grid, 256, smem, st>>>(to_view(*x), N, H, W, sums); }
  KS_DISPATCH_TV(dtype, vec, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_bn_finalize(int C, double count, const double *sums, const float *gamma, const float *beta,
                              float eps, float momentum, float *running_mean, float *running_var,
                              float *scale, float *shift, float *mean, float *rstd, void *stream) {
  KS_CHECK_ARG(C > 0 && count > 0 && sums && scale && shift && mean && rstd);
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(C, count, sums, gamma, beta, eps, momentum,
                                                                    running_mean, running_var, scale, shift, mean, rstd);
  KS_LAUNCH_RET();
}

static inline int pl_grid(int C, int V, long long npix, int per_thread) {
  const int CV = C / V, rows = 256 / CV;
  long long g = (npix + (long long)rows * per_thread - 1) / ((long long)rows * per_thread);
  // the vector kernels hold 2 CTAs per SM (98-128 registers): a grid of exactly the resident CTAs, each looping over more pixels,
  // measured 4-6 % faster than 8 CTAs per SM in 4 waves (scripts/bench_bn.py)
  const long long cap = (long long)kNumSMs * (g_opt.ew_cap > 0 ? g_opt.ew_cap : 2);
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

extern "C" int ks_bn_act(int dtype, int N, int H, int W, const ks_view_t *y, const float *scale, const float *shift,
                         const ks_view_t *res, int relu, const ks_view_t *out, const ks_view_t *pool, void *stream) {
  KS_CHECK_ARG(y && y->ptr && out && out->ptr && scale && shift && N > 0 && H > 0 && W > 0);
  KS_CHECK_ARG(out->C == y->C && (!res || res->C == y->C) && (!pool || pool->C == y->C));
  if (pool) KS_CHECK_ARG(H % 2 == 0 && W % 2 == 0);
  const int es = esize_of(dtype);
  const bool vec = view_vec8_ok(*y, es) && view_vec8_ok(*out, es) && (!res || view_vec8_ok(*res, es)) && (!pool || view_vec8_ok(*pool, es));
  if ((vec ? y->C / 8 : y->C) > 256) return KS_EUNSUPPORTED;
  const bool flat = view_flat(*y, H, W) && view_flat(*out, H, W) && (!res || view_flat(*res, H, W));
  cudaStream_t st = (cudaStream_t)stream;
  const View vy = to_view(*y), vo = to_view(*out), vr = res ? to_view(*res) : vy, vp = pool ? to_view(*pool) : vo;
#define CALL(T, V) { const long long np = (long long)N * (pool ? H / 2 : H) * (pool ? W / 2 : W); \
    const int grid = pl_grid(y->C, V, np, pool ? 2 : 8); \
    if (pool && res) bn_act_kernel<T, V, true, true><<<grid, 256, 0, st>>>(vy, vr, vo, vp, N, H, W, scale, shift, relu, flat); \
    else if (pool) bn_act_kernel<T, V, true, false><<<grid, 256, 0, st>>>(vy, vr, vo, vp, N, H, W, scale, shift, relu, flat); \
    else if (res) bn_act_kernel<T, V, false, true><<<grid, 256, 0, st>>>(vy, vr, vo, vp, N, H, W, scale, shift, relu, flat); \
    else bn_act_kernel<T, V, false, false><<<grid, 256, 0, st>>>(vy, vr, vo, vp, N, H, W, scale, shift, relu, flat); }
  KS_DISPATCH_TV(dtype, vec, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_bn_bwd_reduce(int dtype, int N, int H, int W, const ks_view_t *dout, const ks_view_t *out,
                                const ks_view_t *y, const ks_view_t *dpool, const float *scale, const float *shift,
                                const float *mean, const float *rstd, double *sums, void *stream) {
  KS_CHECK_ARG(dout && y && mean && rstd && sums && N > 0 && H > 0 && W > 0);
  KS_CHECK_ARG(out != nullptr || (scale != nullptr && shift != nullptr));
  KS_CHECK_ARG(dout->C == y->C && (!out || out->C == y->C));
  KS_CHECK_ARG(!dpool || (out && dpool->C == y->C && H % 2 == 0 && W % 2 == 0));
  const int es = esize_of(dtype);
  const bool vec = view_vec8_ok(*y, es) && (!out || view_vec8_ok(*out, es)) && view_vec8_ok(*dout, es) && (!dpool || view_vec8_ok(*dpool, es));
  const bool flat = view_flat(*y, H, W) && view_flat(*dout, H, W) && (!out || view_flat(*out, H, W));
  const long long npix = (long long)N * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  const View vd = to_view(*dout), vy = to_view(*y), vo = out ? to_view(*out) : vy, vp = dpool ? to_view(*dpool) : vy;
#define CALL(T, V) { int grid; size_t smem; int rc = launch_reduce_cfg<T, V>(y->C, dpool ? npix / 2 : npix, grid, smem, 2); if (rc) return rc; \
    if (dpool) bn_bwd_reduce_pool_kernel<T, V><<<grid, 256, smem, st>>>(vd, vo, vy, vp, N, H, W, mean, rstd, sums, flat); \
    else if (out) bn_bwd_reduce_kernel<T, V, true><<<grid, 256, smem, st>>>(vd, vo, vy, N, H, W, scale, shift, mean, rstd, sums, flat); \
    else bn_bwd_reduce_kernel<T, V, false><<<grid, 256, smem, st>>>(vd, vo, vy, N, H, W, scale, shift, mean, rstd, sums, flat); }
  KS_DISPATCH_TV(dtype, vec, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_bn_bwd_apply(int dtype, int N, int H, int W, const ks_view_t *g, int premasked, const ks_view_t *y,
                               const float *scale, const float *shift, const float *mean, const float *rstd, const float *gamma,
                               const double *sums, double count, const ks_view_t *add, const ks_view_t *dy,
                               float *dgamma, float *dbeta, float *dsum_out, int accumulate_param_grads, void *stream) {
  KS_CHECK_ARG(g && y && dy && mean && rstd && sums && count > 0 && N > 0 && H > 0 && W > 0);
  KS_CHECK_ARG(premasked || (scale && shift));
  KS_CHECK_ARG(g->C == y->C && dy->C == y->C && (!add || add->C == y->C));
  const int es = esize_of(dtype);
  const bool has_add = add != nullptr;
  const bool vec = view_vec8_ok(*y, es) && view_vec8_ok(*g, es) && view_vec8_ok(*dy, es) && (!has_add || view_vec8_ok(*add, es));
  if ((vec ? y->C / 8 : y->C) > 256) return KS_EUNSUPPORTED;
  const bool flat = view_flat(*y, H, W) && view_flat(*g, H, W) && view_flat(*dy, H, W) && (!has_add || view_flat(*add, H, W));
  cudaStream_t st = (cudaStream_t)stream;
  const View vg = to_view(*g), vy = to_view(*y), vdy = to_view(*dy), va = has_add ? to_view(*add) : vg;
#define CALL(T, V) { const int grid = pl_grid(y->C, V, (long long)N * H * W, 8); \
    if (has_add) bn_bwd_apply_kernel<T, V, true><<<grid, 256, 0, st>>>(vg, vy, va, vdy, N, H, W, premasked, scale, shift, mean, rstd, gamma, sums, count, \
                                                 dgamma, dbeta, dsum_out, accumulate_param_grads, flat); \
    else bn_bwd_apply_kernel<T, V, false><<<grid, 256, 0, st>>>(vg, vy, va, vdy, N, H, W, premasked, scale, shift, mean, rstd, gamma, sums, count, \
                                                 dgamma, dbeta, dsum_out, accumulate_param_grads, flat); }
  KS_DISPATCH_TV(dtype, vec, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_maxpool2x2_bwd(int dtype, int N, int H, int W, const ks_view_t *x, const ks_view_t *dpool,
                                 const ks_view_t *dx, int accumulate, void *stream) {
  KS_CHECK_ARG(x && dpool && dx && N > 0 && H > 0 && W > 0 && dpool->C == x->C && dx->C == x->C);
  const int es = esize_of(dtype);
  const bool vec = view_vec8_ok(*x, es) && view_vec8_ok(*dpool, es) && view_vec8_ok(*dx, es);
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(T, V) { const long long total = (long long)N * H * W * (x->C / V); const int grid = ew_grid(total, 256); \
    maxpool_bwd_kernel<T, V><<<grid, 256, 0, st>>>(to_view(*x), to_view(*dpool), to_view(*dx), N, H, W, accumulate); }
  KS_DISPATCH_TV(dtype, vec, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_channel_sum(int dtype, int N, int H, int W, const ks_view_t *x, float *out, int accumulate, void *stream) {
  KS_CHECK_ARG(x && x->ptr && out && N > 0 && H > 0 && W > 0);
  const bool vec = view_vec8_ok(*x, esize_of(dtype));
  const long long npix = (long long)N * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) { cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * x->C, st); if (e != cudaSuccess) return (int)e; }
  // every CTA ends with C atomic adds on the same C addresses: few, long-running CTAs (>= 32 rows per row lane) instead of many short
  // ones - on the [13312 x 768] ViT bias gradients 832 CTAs x 768 atomics were the cost, not the 20 MB read
  const int per_lane = g_opt.cs_rows > 0 ? g_opt.cs_rows : 32;
#define CALL(T, V) { int grid; size_t smem; int rc = launch_reduce_cfg<T, V>(x->C, npix, grid, smem, 1); if (rc) return rc; \
    { const int CVv = x->C / V, rowsv = 256 / CVv; long long g2 = (npix + (long long)rowsv * per_lane - 1) / ((long long)rowsv * per_lane); \
      if (g2 < 1) g2 = 1; if (g2 < grid) grid = (int)g2; } \
    channel_sum_kernel<T, V><<<grid, 256, smem, st>>>(to_view(*x), N, H, W, out, view_flat(*x, H, W)); }
  KS_DISPATCH_TV(dtype, vec, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

namespace ks {
// torch.optim.SGD (dampening 0, no nesterov): g += wd*p; buf = momentum*buf + g (buf starts at 0 == torch's first-step clone); p -= lr*buf
__global__ void __launch_bounds__(256)
sgd_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ buf, long long n, float lr, float momentum, float wd, float gscale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pv = p[i];
    const float gk = g[i] * gscale + wd * pv;
    const float b = momentum * buf[i] + gk;
    buf[i] = b;
    p[i] = pv - lr * b;
  }
}
}  // namespace ks

extern "C" int ks_sgd_step(float *p, const float *g, float *buf, int64_t n, float lr, float momentum, float weight_decay, float grad_scale,
                           void *stream) {
  KS_CHECK_ARG(p && g && buf && n > 0);
  long long grid = (n + 256 * 4 - 1) / (256 * 4);
  if (grid > ks::kNumSMs * 8) grid = ks::kNumSMs * 8;
  ks::sgd_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(p, g, buf, n, lr, momentum, weight_decay, grad_scale);
  KS_LAUNCH_RET();
}

extern "C" int ks_adam_step(float *p, const float *g, float *m, float *v, int64_t n, float lr, float beta1, float beta2,
                            float eps, float weight_decay, float grad_scale, int *step_ptr, void *stream) {
  KS_CHECK_ARG(p && g && m && v && step_ptr && n > 0);
  KS_CHECK_ARG(((uintptr_t)p % 16) == 0 && ((uintptr_t)g % 16) == 0 && ((uintptr_t)m % 16) == 0 && ((uintptr_t)v % 16) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ew_grid(n / 4 + 1, 256);
  adam_kernel<false><<<grid, 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, grad_scale, step_ptr);
  incr_kernel<<<1, 1, 0, st>>>(step_ptr);
  KS_LAUNCH_RET();
}

extern "C" int ks_adamw_step(float *p, const float *g, float *m, float *v, int64_t n, float lr, float beta1, float beta2,
                             float eps, float weight_decay, float grad_scale, int *step_ptr, void *stream) {
  KS_CHECK_ARG(p && g && m && v && step_ptr && n > 0);
  KS_CHECK_ARG(((uintptr_t)p % 16) == 0 && ((uintptr_t)g % 16) == 0 && ((uintptr_t)m % 16) == 0 && ((uintptr_t)v % 16) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ks::ew_grid(n / 4 + 1, 256);
  ks::adam_kernel<true><<<grid, 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, grad_scale, step_ptr);
  ks::incr_kernel<<<1, 1, 0, st>>>(step_ptr);
  KS_LAUNCH_RET();
}

// ---- input pipeline on the GPU ------------------------------------------------------------
// dataset/Dataset.py:162-168 (clamp to [0, clamp_input] and nan_to_num(nan = clamp_input); without clamp_input nan_to_num(nan = 200))
// followed by :192-198 (torchvision Normalize(data_mean, data_std): (x - mean_c) / std_c), on the raw float32 SAR planes
// [B][C][HW] AFTER the host->device copy - so the host loop only reads files, and the pinned copy carries raw tiles.
namespace ks {
__global__ void __launch_bounds__(256)
sar_preprocess_kernel(const float *__restrict__ raw, float *__restrict__ out, long long HW, int C, long long total4, long long total,
                      const float *__restrict__ mean, const float *__restrict__ stdv, float clamp_max) {
  auto tr = [&](float v, float m, float s) {
    if (clamp_max > 0.f) v = (v != v) ? clamp_max : fminf(fmaxf(v, 0.f), clamp_max);       // torch.clamp propagates NaN, nan_to_num maps it
    else v = (v != v) ? 200.f : (isinf(v) ? (v > 0.f ? 3.4028234663852886e38f : -3.4028234663852886e38f) : v);
    return __fdiv_rn(__fsub_rn(v, m), s);                                                  // Normalize: sub then div, both IEEE (bit-exact)
  };
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(((i * 4) / HW) % C);                                               // HW % 4 == 0 on this path: one channel per float4
    const float m = __ldg(mean + c), s = __ldg(stdv + c);
    float4 v = __ldcs(reinterpret_cast<const float4 *>(raw) + i);
    v.x = tr(v.x, m, s); v.y = tr(v.y, m, s); v.z = tr(v.z, m, s); v.w = tr(v.w, m, s);
    reinterpret_cast<float4 *>(out)[i] = v;
  }
  if (total4 == 0)
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const int c = (int)((i / HW) % C);
      out[i] = tr(raw[i], __ldg(mean + c), __ldg(stdv + c));
    }
}
}  // namespace ks

extern "C" int ks_sar_preprocess(const float *raw, float *out, int B, int C, int64_t HW, const float *mean, const float *stdv,
                                 float clamp_max, void *stream) {
  KS_CHECK_ARG(raw && out && mean && stdv && B > 0 && C > 0 && HW > 0);
  const long long total = (long long)B * C * HW;
  const bool vec = (HW % 4 == 0) && (((uintptr_t)raw | (uintptr_t)out) % 16 == 0);
  const long long total4 = vec ? total / 4 : 0;
  const int grid = ks::ew_grid(vec ? total4 : total, 256);
  ks::sar_preprocess_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(raw, out, (long long)HW, C, total4, total, mean, stdv, clamp_max);
  KS_LAUNCH_RET();
}
