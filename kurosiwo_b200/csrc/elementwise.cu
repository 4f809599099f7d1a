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
// out = relu(y*scale+shift (+res)), optional fused 2x2 max-pool output
// ------------------------------------------------------------------------------------------
template <typename T, int V, bool POOL>
__global__ void __launch_bounds__(256)
bn_act_kernel(View y, View res, bool has_res, View out, View pool, int N, int H, int W,
              const float *__restrict__ scale, const float *__restrict__ shift, int relu) {
  const int CV = y.C / V;
  const int HH = POOL ? H / 2 : H, WW = POOL ? W / 2 : W;
  const long long total = (long long)N * HH * WW * CV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV); long long r = i / CV;
    const int w = (int)(r % WW); r /= WW; const int h = (int)(r % HH); const int n = (int)(r / HH);
    const int c = cv * V;
    float sc[V], sh[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { sc[k] = __ldg(scale + c + k); sh[k] = __ldg(shift + c + k); }
    if constexpr (POOL) {
      float mx[V];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int hh = 2 * h + (q >> 1), ww = 2 * w + (q & 1);
        float f[V]; VecIO<T, V>::ld(vaddr<T>(y, n, hh, ww, c), f);
        float rr[V];
        if (has_res) VecIO<T, V>::ld(vaddr<T>(res, n, hh, ww, c), rr);
        float o[V];
#pragma unroll
        for (int k = 0; k < V; ++k) {
          float v = fmaf(f[k], sc[k], sh[k]);
          if (has_res) v += rr[k];
          if (relu) v = fmaxf(v, 0.f);
          o[k] = v;
        }
        VecIO<T, V>::st(vaddr<T>(out, n, hh, ww, c), o);
        // pool over the values AS STORED (bf16-rounded in bf16 mode)
#pragma unroll
        for (int k = 0; k < V; ++k) { const float os = round_as<T>(o[k]); mx[k] = (q == 0) ? os : fmaxf(mx[k], os); }
      }
      VecIO<T, V>::st(vaddr<T>(pool, n, h, w, c), mx);
    } else {
      float f[V]; VecIO<T, V>::ld(vaddr<T>(y, n, h, w, c), f);
      float rr[V];
      if (has_res) VecIO<T, V>::ld(vaddr<T>(res, n, h, w, c), rr);
      float o[V];
#pragma unroll
      for (int k = 0; k < V; ++k) {
        float v = fmaf(f[k], sc[k], sh[k]);
        if (has_res) v += rr[k];
        if (relu) v = fmaxf(v, 0.f);
        o[k] = v;
      }
      VecIO<T, V>::st(vaddr<T>(out, n, h, w, c), o);
    }
  }
}

// ------------------------------------------------------------------------------------------
// BN backward: reductions, then apply
// ------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(View dout, View out, View y, int N, int H, int W,
                     const float *__restrict__ mean, const float *__restrict__ rstd, double *sums) {
  extern __shared__ float smem[];
  const int C = y.C, CV = C / V, rows = blockDim.x / CV;
  const int tx = threadIdx.x % CV, ty = threadIdx.x / CV;
  float acc[2][V];
#pragma unroll
  for (int i = 0; i < V; ++i) { acc[0][i] = 0.f; acc[1][i] = 0.f; }
  const long long npix = (long long)N * H * W;
  if (ty < rows) {
    float mu[V], rs[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { mu[k] = mean[tx * V + k]; rs[k] = rstd[tx * V + k]; }
    for (long long p = (long long)blockIdx.x * rows + ty; p < npix; p += (long long)gridDim.x * rows) {
      const int w = (int)(p % W); long long r = p / W; const int h = (int)(r % H); const int n = (int)(r / H);
      float g[V], o[V], f[V];
      VecIO<T, V>::ld(vaddr<T>(dout, n, h, w, tx * V), g);
      VecIO<T, V>::ld(vaddr<T>(out, n, h, w, tx * V), o);
      VecIO<T, V>::ld(vaddr<T>(y, n, h, w, tx * V), f);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float gk = (o[k] > 0.f) ? g[k] : 0.f;
        acc[0][k] += gk;
        acc[1][k] += gk * ((f[k] - mu[k]) * rs[k]);
      }
    }
  }
  block_channel_reduce<2, V>(acc, CV, rows, C, sums, smem);
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(View dout, View out, View y, View add_dout, View add_out, bool has_add, View dy,
                    int N, int H, int W, const float *__restrict__ mean, const float *__restrict__ rstd,
                    const float *__restrict__ gamma, const double *__restrict__ sums, double count,
                    float *dgamma, float *dbeta, int accumulate) {
  const int C = y.C, CV = C / V;
  const long long total = (long long)N * H * W * CV;
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)sums[C + c];
      if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)sums[c];
    }
  }
  const float invM = (float)(1.0 / count);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV); long long r = i / CV;
    const int w = (int)(r % W); r /= W; const int h = (int)(r % H); const int n = (int)(r / H);
    const int c = cv * V;
    float g[V], o[V], f[V], ag[V], ao[V], d[V];
    VecIO<T, V>::ld(vaddr<T>(dout, n, h, w, c), g);
    VecIO<T, V>::ld(vaddr<T>(out, n, h, w, c), o);
    VecIO<T, V>::ld(vaddr<T>(y, n, h, w, c), f);
    if (has_add) { VecIO<T, V>::ld(vaddr<T>(add_dout, n, h, w, c), ag); VecIO<T, V>::ld(vaddr<T>(add_out, n, h, w, c), ao); }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float mu = __ldg(mean + c + k), rs = __ldg(rstd + c + k), ga = gamma ? __ldg(gamma + c + k) : 1.f;
      const float sg = (float)sums[c + k] * invM, sgx = (float)sums[C + c + k] * invM;
      const float gk = (o[k] > 0.f) ? g[k] : 0.f;
      const float xh = (f[k] - mu) * rs;
      float v = ga * rs * (gk - sg - xh * sgx);
      if (has_add) v += (ao[k] > 0.f) ? ag[k] : 0.f;
      d[k] = v;
    }
    VecIO<T, V>::st(vaddr<T>(dy, n, h, w, c), d);
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
channel_sum_kernel(View x, int N, int H, int W, float *out) {
  extern __shared__ float smem[];
  const int C = x.C, CV = C / V, rows = blockDim.x / CV;
  const int tx = threadIdx.x % CV, ty = threadIdx.x / CV;
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  const long long npix = (long long)N * H * W;
  if (ty < rows) {
    for (long long p = (long long)blockIdx.x * rows + ty; p < npix; p += (long long)gridDim.x * rows) {
      const int w = (int)(p % W); long long r = p / W; const int h = (int)(r % H); const int n = (int)(r / H);
      float f[V]; VecIO<T, V>::ld(vaddr<T>(x, n, h, w, tx * V), f);
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] += f[i];
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
// Adam (torch.optim.Adam semantics, L2 weight decay folded into the gradient)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
            long long n, float lr, float b1, float b2, float eps, float wd, float gscale, const int *step_ptr) {
  const int t = *step_ptr + 1;
  const double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4 *>(p)[i], gg = reinterpret_cast<const float4 *>(g)[i];
    float4 mm = reinterpret_cast<float4 *>(m)[i], vv = reinterpret_cast<float4 *>(v)[i];
    float *P = &pp.x, *G = &gg.x, *M = &mm.x, *Vv = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gk = G[k] * gscale + wd * P[k];
      M[k] = b1 * M[k] + (1.f - b1) * gk;
      Vv[k] = b2 * Vv[k] + (1.f - b2) * gk * gk;
      P[k] -= step_size * (M[k] / (sqrtf(Vv[k]) * inv_bc2_sqrt + eps));
    }
    reinterpret_cast<float4 *>(p)[i] = pp; reinterpret_cast<float4 *>(m)[i] = mm; reinterpret_cast<float4 *>(v)[i] = vv;
  }
  if (blockIdx.x == 0) {
    for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
      float gk = g[i] * gscale + wd * p[i];
      m[i] = b1 * m[i] + (1.f - b1) * gk;
      v[i] = b2 * v[i] + (1.f - b2) * gk * gk;
      p[i] -= step_size * (m[i] / (sqrtf(v[i]) * inv_bc2_sqrt + eps));
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

extern "C" int ks_bn_act(int dtype, int N, int H, int W, const ks_view_t *y, const float *scale, const float *shift,
                         const ks_view_t *res, int relu, const ks_view_t *out, const ks_view_t *pool, void *stream) {
  KS_CHECK_ARG(y && y->ptr && out && out->ptr && scale && shift && N > 0 && H > 0 && W > 0);
  KS_CHECK_ARG(out->C == y->C && (!res || res->C == y->C) && (!pool || pool->C == y->C));
  if (pool) KS_CHECK_ARG(H % 2 == 0 && W % 2 == 0);
  const int es = esize_of(dtype);
  const bool vec = view_vec8_ok(*y, es) && view_vec8_ok(*out, es) && (!res || view_vec8_ok(*res, es)) && (!pool || view_vec8_ok(*pool, es));
  cudaStream_t st = (cudaStream_t)stream;
  const View vy = to_view(*y), vo = to_view(*out), vr = res ? to_view(*res) : vy, vp = pool ? to_view(*pool) : vo;
#define CALL(T, V) { const long long total = (long long)N * (pool ? H / 2 : H) * (pool ? W / 2 : W) * (y->C / V); \
    const int grid = ew_grid(total, 256); \
    if (pool) bn_act_kernel<T, V, true><<<grid, 256, 0, st>>>(vy, vr, res != nullptr, vo, vp, N, H, W, scale, shift, relu); \
    else bn_act_kernel<T, V, false><<<grid, 256, 0, st>>>(vy, vr, res != nullptr, vo, vp, N, H, W, scale, shift, relu); }
  KS_DISPATCH_TV(dtype, vec, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_bn_bwd_reduce(int dtype, int N, int H, int W, const ks_view_t *dout, const ks_view_t *out,
                                const ks_view_t *y, const float *mean, const float *rstd, double *sums, void *stream) {
  KS_CHECK_ARG(dout && out && y && mean && rstd && sums && N > 0 && H > 0 && W > 0);
  KS_CHECK_ARG(dout->C == y->C && out->C == y->C);
  const int es = esize_of(dtype);
  const bool vec = view_vec8_ok(*y, es) && view_vec8_ok(*out, es) && view_vec8_ok(*dout, es);
  const long long npix = (long long)N * H * W;
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(T, V) { int grid; size_t smem; int rc = launch_reduce_cfg<T, V>(y->C, npix, grid, smem, 2); if (rc) return rc; \
    bn_bwd_reduce_kernel<T, V><<<grid, 256, smem, st>>>(to_view(*dout), to_view(*out), to_view(*y), N, H, W, mean, rstd, sums); }
  KS_DISPATCH_TV(dtype, vec, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_bn_bwd_apply(int dtype, int N, int H, int W, const ks_view_t *dout, const ks_view_t *out,
                               const ks_view_t *y, const float *mean, const float *rstd, const float *gamma,
                               const double *sums, double count, const ks_view_t *add_dout, const ks_view_t *add_out,
                               const ks_view_t *dy, float *dgamma, float *dbeta, int accumulate_param_grads, void *stream) {
  KS_CHECK_ARG(dout && out && y && dy && mean && rstd && sums && count > 0 && N > 0 && H > 0 && W > 0);
  KS_CHECK_ARG((add_dout == nullptr) == (add_out == nullptr));
  KS_CHECK_ARG(dout->C == y->C && out->C == y->C && dy->C == y->C);
  const int es = esize_of(dtype);
  const bool has_add = add_dout != nullptr;
  const bool vec = view_vec8_ok(*y, es) && view_vec8_ok(*out, es) && view_vec8_ok(*dout, es) && view_vec8_ok(*dy, es) &&
                   (!has_add || (view_vec8_ok(*add_dout, es) && view_vec8_ok(*add_out, es)));
  cudaStream_t st = (cudaStream_t)stream;
  const View vd = to_view(*dout), vo = to_view(*out), vy = to_view(*y), vdy = to_view(*dy);
  const View vad = has_add ? to_view(*add_dout) : vd, vao = has_add ? to_view(*add_out) : vo;
#define CALL(T, V) { const long long total = (long long)N * H * W * (y->C / V); const int grid = ew_grid(total, 256); \
    bn_bwd_apply_kernel<T, V><<<grid, 256, 0, st>>>(vd, vo, vy, vad, vao, has_add, vdy, N, H, W, mean, rstd, gamma, sums, count, dgamma, dbeta, accumulate_param_grads); }
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
#define CALL(T, V) { int grid; size_t smem; int rc = launch_reduce_cfg<T, V>(x->C, npix, grid, smem, 1); if (rc) return rc; \
    channel_sum_kernel<T, V><<<grid, 256, smem, st>>>(to_view(*x), N, H, W, out); }
  KS_DISPATCH_TV(dtype, vec, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_adam_step(float *p, const float *g, float *m, float *v, int64_t n, float lr, float beta1, float beta2,
                            float eps, float weight_decay, float grad_scale, int *step_ptr, void *stream) {
  KS_CHECK_ARG(p && g && m && v && step_ptr && n > 0);
  KS_CHECK_ARG(((uintptr_t)p % 16) == 0 && ((uintptr_t)g % 16) == 0 && ((uintptr_t)m % 16) == 0 && ((uintptr_t)v % 16) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ew_grid(n / 4 + 1, 256);
  adam_kernel<<<grid, 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, grad_scale, step_ptr);
  incr_kernel<<<1, 1, 0, st>>>(step_ptr);
  KS_LAUNCH_RET();
}
