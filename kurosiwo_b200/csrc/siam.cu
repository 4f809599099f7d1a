// Bandwidth-bound passes of the Siamese U-Nets (FC-Siam-conc / FC-Siam-diff):
//   Softmax / LogSoftmax head          models/siam_conc.py:93,177, models/siam_diff.py:93,173
//   Dropout2d (per-sample channel mask) models/siam_conc.py:21 ... (p = 0.2 after every conv+BN+ReLU)
//   |x1 - x2| skip connections          models/siam_diff.py:141,150,158,165
// All kernels work on strided NHWC views (channel stride 1), 8 channels (16 bytes of bf16) per thread.
#include "common.cuh"

namespace ks {

template <typename T>
__device__ __forceinline__ T *vp(const View &v, long long p, int H, int W, int c) {
  const int w = (int)(p % W); const long long r = p / W; const int h = (int)(r % H); const long long n = r / H;
  return reinterpret_cast<T *>(v.ptr) + (n * v.sn + (long long)h * v.sh + (long long)w * v.sw + c);
}

// ---- softmax head ----------------------------------------------------------------------------------------
// z: NHWC view, logits in channels 0..K-1 (the conv engine pads Cout to 16); out: NCHW fp32 (what the loss reads).
template <typename T, int K>
__global__ void __launch_bounds__(256)
softmax_head_fwd_kernel(View z, int H, int W, long long NP, int log_mode, float *__restrict__ out) {
  const long long HW = (long long)H * W;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < NP; p += (long long)gridDim.x * blockDim.x) {
    const T *zp = vp<T>(z, p, H, W, 0);
    float v[K], m = -INFINITY;
#pragma unroll
    for (int k = 0; k < K; ++k) { v[k] = Cvt<T>::ld(zp + k); m = fmaxf(m, v[k]); }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) { v[k] -= m; s += expf(v[k]); }
    const long long n = p / HW, q = p % HW;
    const float ls = logf(s), inv = 1.f / s;
#pragma unroll
    for (int k = 0; k < K; ++k) out[(n * K + k) * HW + q] = log_mode ? (v[k] - ls) : expf(v[k]) * inv;
  }
}

// dz_k = y_k (g_k - sum_j g_j y_j)            (softmax)
// dz_k = g_k - exp(y_k) sum_j g_j             (log-softmax)      channels K..C-1 of dz are written as 0.
template <typename T, int K>
__global__ void __launch_bounds__(256)
softmax_head_bwd_kernel(View dz, int H, int W, long long NP, int log_mode, const float *__restrict__ out,
                        const float *__restrict__ dout) {
  const long long HW = (long long)H * W;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < NP; p += (long long)gridDim.x * blockDim.x) {
    const long long n = p / HW, q = p % HW;
    float y[K], g[K], dot = 0.f, gs = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      y[k] = out[(n * K + k) * HW + q]; g[k] = dout[(n * K + k) * HW + q];
      dot += g[k] * y[k]; gs += g[k];
    }
    T *dp = vp<T>(dz, p, H, W, 0);
    for (int c = 0; c < dz.C; c += 8) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        o[i] = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k)
          if (c + i == k) o[i] = log_mode ? (g[k] - expf(y[k]) * gs) : y[k] * (g[k] - dot);
      }
      st8(dp + c, o);
    }
  }
}

// ---- Dropout2d -------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {   // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// mask[i] = 1/(1-p) with probability 1-p else 0; the stream is a pure function of (seed, step counter, i), so a
// captured CUDA graph draws fresh masks on every replay (the counter lives in device memory).
__global__ void dropout_mask_kernel(float *mask, long long n, float p, unsigned long long seed, const int *step_ptr) {
  const unsigned long long step = step_ptr ? (unsigned long long)(unsigned int)*step_ptr : 0ull;
  const float keep = 1.f / (1.f - p);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long r = mix64(mix64(seed ^ (step << 32)) + (unsigned long long)i);
    const float u = (float)(r >> 40) * (1.0f / 16777216.0f);
    mask[i] = (u >= p) ? keep : 0.f;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
channel_scale_kernel(View x, int H, int W, long long NP, const float *__restrict__ m) {
  const int CV = x.C / 8;
  const long long total = NP * CV, HW = (long long)H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / CV; const int c = (int)(i % CV) * 8;
    const long long n = p / HW;
    T *xp = vp<T>(x, p, H, W, c);
    float f[8]; ld8(xp, f);
    const float *mp = m + n * x.C + c;
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] *= mp[k];
    st8(xp, f);
  }
}

// ---- |a - b| ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
absdiff_fwd_kernel(View a, View b, View o, int H, int W, long long NP) {
  const int CV = a.C / 8;
  const long long total = NP * CV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / CV; const int c = (int)(i % CV) * 8;
    float fa[8], fb[8];
    ld8(vp<T>(a, p, H, W, c), fa); ld8(vp<T>(b, p, H, W, c), fb);
#pragma unroll
    for (int k = 0; k < 8; ++k) fa[k] = fabsf(fa[k] - fb[k]);
    st8(vp<T>(o, p, H, W, c), fa);
  }
}

// da (+)= sign(a-b) g ; db (+)= -sign(a-b) g      (aten abs backward: sign(0) = 0)
template <typename T>
__global__ void __launch_bounds__(256)
absdiff_bwd_kernel(View a, View b, View g, View da, View db, int acc_a, int acc_b, int H, int W, long long NP) {
  const int CV = a.C / 8;
  const long long total = NP * CV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / CV; const int c = (int)(i % CV) * 8;
    float fa[8], fb[8], fg[8], oa[8], ob[8];
    ld8(vp<T>(a, p, H, W, c), fa); ld8(vp<T>(b, p, H, W, c), fb); ld8(vp<T>(g, p, H, W, c), fg);
    if (acc_a) ld8(vp<T>(da, p, H, W, c), oa);
    if (acc_b) ld8(vp<T>(db, p, H, W, c), ob);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float d = fa[k] - fb[k];
      const float s = d > 0.f ? fg[k] : (d < 0.f ? -fg[k] : 0.f);
      oa[k] = acc_a ? oa[k] + s : s;
      ob[k] = acc_b ? ob[k] - s : -s;
    }
    st8(vp<T>(da, p, H, W, c), oa); st8(vp<T>(db, p, H, W, c), ob);
  }
}

static inline bool vec_ok(const ks_view_t *v, int esize) {
  return v && v->ptr && v->C > 0 && (v->C % 8 == 0) && (((uintptr_t)v->ptr % 16) == 0) && ((v->sn * esize) % 16 == 0) &&
         ((v->sh * esize) % 16 == 0) && ((v->sw * esize) % 16 == 0);
}
static inline int grid_for(long long work) {
  long long g = (work + 255) / 256; const long long cap = (long long)kNumSMs * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace ks

using namespace ks;

extern "C" int ks_softmax_head_fwd(int dtype, int N, int H, int W, const ks_view_t *z, int K, int log_mode,
                                   float *out, void *stream) {
  KS_CHECK_ARG(z && z->ptr && out && N > 0 && H > 0 && W > 0 && z->C >= K);
  if (K != 3) return KS_EUNSUPPORTED;
  const long long NP = (long long)N * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == KS_F32) softmax_head_fwd_kernel<float, 3><<<grid_for(NP), 256, 0, st>>>(to_view(*z), H, W, NP, log_mode, out);
  else if (dtype == KS_BF16) softmax_head_fwd_kernel<__nv_bfloat16, 3><<<grid_for(NP), 256, 0, st>>>(to_view(*z), H, W, NP, log_mode, out);
  else return KS_EINVAL;
  KS_LAUNCH_RET();
}

extern "C" int ks_softmax_head_bwd(int dtype, int N, int H, int W, const float *out, const float *dout, int K,
                                   int log_mode, const ks_view_t *dz, void *stream) {
  KS_CHECK_ARG(out && dout && N > 0 && H > 0 && W > 0 && dz && dz->C >= K);
  if (K != 3) return KS_EUNSUPPORTED;
  if (!vec_ok(dz, dtype == KS_F32 ? 4 : 2)) return KS_EUNSUPPORTED;
  const long long NP = (long long)N * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == KS_F32) softmax_head_bwd_kernel<float, 3><<<grid_for(NP), 256, 0, st>>>(to_view(*dz), H, W, NP, log_mode, out, dout);
  else if (dtype == KS_BF16) softmax_head_bwd_kernel<__nv_bfloat16, 3><<<grid_for(NP), 256, 0, st>>>(to_view(*dz), H, W, NP, log_mode, out, dout);
  else return KS_EINVAL;
  KS_LAUNCH_RET();
}

extern "C" int ks_dropout_mask(float *mask, int64_t n, float p, uint64_t seed, const int *step_ptr, void *stream) {
  KS_CHECK_ARG(mask && n > 0 && p >= 0.f && p < 1.f);
  dropout_mask_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(mask, n, p, seed, step_ptr);
  KS_LAUNCH_RET();
}

extern "C" int ks_channel_scale(int dtype, int N, int H, int W, const ks_view_t *x, const float *m, void *stream) {
  KS_CHECK_ARG(m && N > 0 && H > 0 && W > 0);
  if (!vec_ok(x, dtype == KS_F32 ? 4 : 2)) return KS_EUNSUPPORTED;
  const long long NP = (long long)N * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == KS_F32) channel_scale_kernel<float><<<grid_for(NP * (x->C / 8)), 256, 0, st>>>(to_view(*x), H, W, NP, m);
  else if (dtype == KS_BF16) channel_scale_kernel<__nv_bfloat16><<<grid_for(NP * (x->C / 8)), 256, 0, st>>>(to_view(*x), H, W, NP, m);
  else return KS_EINVAL;
  KS_LAUNCH_RET();
}

extern "C" int ks_absdiff_fwd(int dtype, int N, int H, int W, const ks_view_t *a, const ks_view_t *b,
                              const ks_view_t *out, void *stream) {
  KS_CHECK_ARG(N > 0 && H > 0 && W > 0);
  const int es = dtype == KS_F32 ? 4 : 2;
  if (!vec_ok(a, es) || !vec_ok(b, es) || !vec_ok(out, es)) return KS_EUNSUPPORTED;
  KS_CHECK_ARG(a->C == b->C && a->C == out->C);
  const long long NP = (long long)N * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  const int g = grid_for(NP * (a->C / 8));
  if (dtype == KS_F32) absdiff_fwd_kernel<float><<<g, 256, 0, st>>>(to_view(*a), to_view(*b), to_view(*out), H, W, NP);
  else if (dtype == KS_BF16) absdiff_fwd_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(to_view(*a), to_view(*b), to_view(*out), H, W, NP);
  else return KS_EINVAL;
  KS_LAUNCH_RET();
}

extern "C" int ks_absdiff_bwd(int dtype, int N, int H, int W, const ks_view_t *a, const ks_view_t *b, const ks_view_t *g,
                              const ks_view_t *da, int accumulate_a, const ks_view_t *db, int accumulate_b, void *stream) {
  KS_CHECK_ARG(N > 0 && H > 0 && W > 0);
  const int es = dtype == KS_F32 ? 4 : 2;
  if (!vec_ok(a, es) || !vec_ok(b, es) || !vec_ok(g, es) || !vec_ok(da, es) || !vec_ok(db, es)) return KS_EUNSUPPORTED;
  KS_CHECK_ARG(a->C == b->C && a->C == g->C && a->C == da->C && a->C == db->C);
  const long long NP = (long long)N * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  const int gr = grid_for(NP * (a->C / 8));
  if (dtype == KS_F32) absdiff_bwd_kernel<float><<<gr, 256, 0, st>>>(to_view(*a), to_view(*b), to_view(*g), to_view(*da), to_view(*db), accumulate_a, accumulate_b, H, W, NP);
  else if (dtype == KS_BF16) absdiff_bwd_kernel<__nv_bfloat16><<<gr, 256, 0, st>>>(to_view(*a), to_view(*b), to_view(*g), to_view(*da), to_view(*db), accumulate_a, accumulate_b, H, W, NP);
  else return KS_EINVAL;
  KS_LAUNCH_RET();
}
