// Stem convolution conv0_0.conv1 (models/snunet.py:75: Conv2d(in_channels=2|3, 32, 3, padding=1)):
// forward and weight gradient.  Cin is 2 (VV,VH) or 3 (+DEM): far too thin for tensor cores and
// purely HBM-bound (read 8 B/px of input, write/read 64 B/px of the 32-channel tensor), so both
// kernels read the NCHW fp32 network input directly (no layout pass) and put the OUTPUT channel on
// the warp lane: every activation access is one fully coalesced 32-lane row of an NHWC pixel, and
// the 3x3xCin input window comes from a shared-memory halo tile as broadcast reads.
#include "common.cuh"

namespace ks {

constexpr int ST_TH = 8, ST_TW = 32;  // pixels per block tile: 8 rows (one per warp) x 32 columns

template <int CIN>
__device__ __forceinline__ void load_x_tile(float4 (*xt)[ST_TW + 2], const float *__restrict__ x, int n, int H, int W,
                                            int h0, int w0, bool round_bf16) {
  for (int i = threadIdx.x; i < (ST_TH + 2) * (ST_TW + 2); i += blockDim.x) {
    const int r = i / (ST_TW + 2), c = i % (ST_TW + 2);
    const int h = h0 + r - 1, w = w0 + c - 1;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (h >= 0 && h < H && w >= 0 && w < W) {
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        float t = __ldg(x + (((long long)n * CIN + ci) * H + h) * W + w);
        v[ci] = round_bf16 ? __bfloat162float(__float2bfloat16_rn(t)) : t;
      }
    }
    xt[r][c] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// out[n,h,w,co] = bias[co] + sum_{tap,ci} x[n,ci,h+dy,w+dx] * w[co][ci][tap];  lane = co (Cout == 32)
template <typename T, int CIN>
__global__ void __launch_bounds__(256)
stem_fwd_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias, View dst,
                int N, int H, int W, int tiles_h, int tiles_w) {
  __shared__ float4 xt[ST_TH + 2][ST_TW + 2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr bool RB = sizeof(T) == 2;
  float wr[9][CIN];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const float v = __ldg(w + ((long long)lane * CIN + ci) * 9 + t);
      wr[t][ci] = RB ? __bfloat162float(__float2bfloat16_rn(v)) : v;
    }
  const float b = bias ? __ldg(bias + lane) : 0.f;
  const long long total = (long long)N * tiles_h * tiles_w;
  for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int tw = (int)(tile % tiles_w); long long r = tile / tiles_w;
    const int th = (int)(r % tiles_h); const int n = (int)(r / tiles_h);
    const int h0 = th * ST_TH, w0 = tw * ST_TW;
    __syncthreads();
    load_x_tile<CIN>(xt, x, n, H, W, h0, w0, RB);
    __syncthreads();
    const int h = h0 + warp;
    if (h < H) {
      T *orow = reinterpret_cast<T *>(dst.ptr) + ((long long)n * dst.sn + (long long)h * dst.sh + lane);
      const int wmax = min(ST_TW, W - w0);
      for (int c = 0; c < wmax; ++c) {
        float acc = b;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 v = xt[warp + t / 3][c + t % 3];
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) acc = fmaf(vv[ci], wr[t][ci], acc);
        }
        Cvt<T>::st(orow + (long long)(w0 + c) * dst.sw, acc);
      }
    }
  }
}

// dw[co][ci][tap] (+)= sum_px dy[px][co] * x[px+tap][ci];  lane = co, registers hold the 9*CIN partial sums
template <typename T, int CIN>
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(const float *__restrict__ x, View dy, float *dw, int N, int H, int W, int tiles_h, int tiles_w) {
  __shared__ float4 xt[ST_TH + 2][ST_TW + 2];
  __shared__ float red[8][9 * CIN][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr bool RB = sizeof(T) == 2;
  float acc[9][CIN];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) acc[t][ci] = 0.f;
  const long long total = (long long)N * tiles_h * tiles_w;
  for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int tw = (int)(tile % tiles_w); long long r = tile / tiles_w;
    const int th = (int)(r % tiles_h); const int n = (int)(r / tiles_h);
    const int h0 = th * ST_TH, w0 = tw * ST_TW;
    __syncthreads();
    load_x_tile<CIN>(xt, x, n, H, W, h0, w0, RB);
    __syncthreads();
    const int h = h0 + warp;
    if (h < H) {
      const T *grow = reinterpret_cast<const T *>(dy.ptr) + ((long long)n * dy.sn + (long long)h * dy.sh + lane);
      const int wmax = min(ST_TW, W - w0);
      for (int c = 0; c < wmax; ++c) {
        const float g = Cvt<T>::ld(grow + (long long)(w0 + c) * dy.sw);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 v = xt[warp + t / 3][c + t % 3];
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) acc[t][ci] = fmaf(g, vv[ci], acc[t][ci]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) red[warp][t * CIN + ci][lane] = acc[t][ci];
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * CIN * 32; i += blockDim.x) {
    const int co = i % 32, k = i / 32;       // k = t*CIN + ci
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s += red[wv][k][co];
    const int t = k / CIN, ci = k % CIN;
    atomicAdd(dw + ((long long)co * CIN + ci) * 9 + t, s);
  }
}

}  // namespace ks

using namespace ks;

extern "C" int ks_bn_stats(int dtype, int N, int H, int W, const ks_view_t *x, double *sums, void *stream);

template <typename T>
static int stem_fwd_dispatch(int Cin, dim3 grid, cudaStream_t st, const float *x, const float *w, const float *bias, View dst,
                             int N, int H, int W, int th, int tw) {
  switch (Cin) {
    case 1: stem_fwd_kernel<T, 1><<<grid, 256, 0, st>>>(x, w, bias, dst, N, H, W, th, tw); break;
    case 2: stem_fwd_kernel<T, 2><<<grid, 256, 0, st>>>(x, w, bias, dst, N, H, W, th, tw); break;
    case 3: stem_fwd_kernel<T, 3><<<grid, 256, 0, st>>>(x, w, bias, dst, N, H, W, th, tw); break;
    case 4: stem_fwd_kernel<T, 4><<<grid, 256, 0, st>>>(x, w, bias, dst, N, H, W, th, tw); break;
    default: return KS_EUNSUPPORTED;
  }
  return (int)cudaGetLastError();
}
template <typename T>
static int stem_wgrad_dispatch(int Cin, dim3 grid, cudaStream_t st, const float *x, View dy, float *dw, int N, int H, int W, int th, int tw) {
  switch (Cin) {
    case 1: stem_wgrad_kernel<T, 1><<<grid, 256, 0, st>>>(x, dy, dw, N, H, W, th, tw); break;
    case 2: stem_wgrad_kernel<T, 2><<<grid, 256, 0, st>>>(x, dy, dw, N, H, W, th, tw); break;
    case 3: stem_wgrad_kernel<T, 3><<<grid, 256, 0, st>>>(x, dy, dw, N, H, W, th, tw); break;
    case 4: stem_wgrad_kernel<T, 4><<<grid, 256, 0, st>>>(x, dy, dw, N, H, W, th, tw); break;
    default: return KS_EUNSUPPORTED;
  }
  return (int)cudaGetLastError();
}

extern "C" int ks_stem_conv3x3(int dtype, int N, int Cin, int H, int W, const float *x_nchw, const float *w_oihw,
                               const float *bias, const ks_view_t *dst, double *stats, void *stream) {
  KS_CHECK_ARG(x_nchw && w_oihw && dst && dst->ptr && N > 0 && H > 0 && W > 0);
  if (dst->C != 32 || Cin < 1 || Cin > 4) return KS_EUNSUPPORTED;
  const int th = (H + ST_TH - 1) / ST_TH, tw = (W + ST_TW - 1) / ST_TW;
  const long long total = (long long)N * th * tw;
  const long long cap = (long long)kNumSMs * 8;
  dim3 grid((unsigned)(total < cap ? total : cap));
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (dtype == KS_F32) rc = stem_fwd_dispatch<float>(Cin, grid, st, x_nchw, w_oihw, bias, to_view(*dst), N, H, W, th, tw);
  else if (dtype == KS_BF16) rc = stem_fwd_dispatch<__nv_bfloat16>(Cin, grid, st, x_nchw, w_oihw, bias, to_view(*dst), N, H, W, th, tw);
  else return KS_EINVAL;
  if (rc) return rc;
  if (stats) return ks_bn_stats(dtype, N, H, W, dst, stats, stream);
  return KS_OK;
}

extern "C" int ks_stem_wgrad3x3(int dtype, int N, int Cin, int H, int W, const float *x_nchw, const ks_view_t *dy,
                                float *dw_oihw, int accumulate, void *stream) {
  KS_CHECK_ARG(x_nchw && dy && dy->ptr && dw_oihw && N > 0 && H > 0 && W > 0);
  if (dy->C != 32 || Cin < 1 || Cin > 4) return KS_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) { cudaError_t e = cudaMemsetAsync(dw_oihw, 0, sizeof(float) * 32 * Cin * 9, st); if (e != cudaSuccess) return (int)e; }
  const int th = (H + ST_TH - 1) / ST_TH, tw = (W + ST_TW - 1) / ST_TW;
  const long long total = (long long)N * th * tw;
  const long long cap = (long long)kNumSMs * 4;
  dim3 grid((unsigned)(total < cap ? total : cap));
  if (dtype == KS_F32) return stem_wgrad_dispatch<float>(Cin, grid, st, x_nchw, to_view(*dy), dw_oihw, N, H, W, th, tw);
  if (dtype == KS_BF16) return stem_wgrad_dispatch<__nv_bfloat16>(Cin, grid, st, x_nchw, to_view(*dy), dw_oihw, N, H, W, th, tw);
  return KS_EINVAL;
}
