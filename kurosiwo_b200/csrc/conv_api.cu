// C-ABI entry points of the convolution engine: argument checks + dispatch between the
// tcgen05 engine (conv_tc.cu / wgrad_tc.cu, bf16) and the CUDA-core engine (conv_simt.cu).
#include "common.cuh"

namespace ks {
int conv2d_simt(int dtype, int N, int H, int W, int ksize, const ViewList &srcs, const void *weight, const float *bias,
                const ViewList &dsts, int acc_mask, double *stats, cudaStream_t st);
int wgrad_simt(int dtype, int N, int H, int W, int ksize, const ViewList &xs, const ViewList &dys, float *dw, cudaStream_t st);
int conv2d_tc(int N, int H, int W, int ksize, const ViewList &srcs, const void *weight, const float *bias,
              const ViewList &dsts, int acc_mask, double *stats, cudaStream_t st);
int wgrad_tc(int N, int H, int W, int ksize, const ViewList &xs, const ViewList &dys, float *dw, cudaStream_t st, float *dsum = nullptr,
             int dsum_mod = 0, bool *dsum_done = nullptr);
}  // namespace ks

using namespace ks;

extern "C" int ks_conv2d(int dtype, int N, int H, int W, int ksize, const ks_view_t *srcs, int n_src,
                         const void *weight, const float *bias, const ks_view_t *dsts, int n_dst,
                         const int *dst_accumulate, double *stats, int impl, void *stream) {
  KS_CHECK_ARG(N > 0 && H > 0 && W > 0 && (ksize == 1 || ksize == 3) && weight);
  KS_CHECK_ARG(dtype == KS_F32 || dtype == KS_BF16);
  ViewList sl, dl;
  int rc = make_view_list(srcs, n_src, sl); if (rc) return rc;
  rc = make_view_list(dsts, n_dst, dl); if (rc) return rc;
  int mask = 0;
  if (dst_accumulate) for (int i = 0; i < n_dst; ++i) if (dst_accumulate[i]) mask |= 1 << i;
  if (stats && mask) return KS_EINVAL;  // statistics are of the stored conv output
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == KS_IMPL_TC || (impl == KS_IMPL_AUTO && dtype == KS_BF16)) {
    if (dtype != KS_BF16) return KS_EUNSUPPORTED;
    rc = conv2d_tc(N, H, W, ksize, sl, weight, bias, dl, mask, stats, st);
    if (rc != KS_EUNSUPPORTED || impl == KS_IMPL_TC) return rc;
  }
  return conv2d_simt(dtype, N, H, W, ksize, sl, weight, bias, dl, mask, stats, st);
}

extern "C" int ks_conv2d_wgrad(int dtype, int N, int H, int W, int ksize, const ks_view_t *xs, int n_x,
                               const ks_view_t *dys, int n_dy, float *dw, int accumulate, int impl, void *stream) {
  KS_CHECK_ARG(N > 0 && H > 0 && W > 0 && (ksize == 1 || ksize == 3) && dw);
  KS_CHECK_ARG(dtype == KS_F32 || dtype == KS_BF16);
  ViewList xl, yl;
  int rc = make_view_list(xs, n_x, xl); if (rc) return rc;
  rc = make_view_list(dys, n_dy, yl); if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) {
    const size_t bytes = sizeof(float) * (size_t)ksize * ksize * xl.cstart[xl.n] * yl.cstart[yl.n];
    cudaError_t e = cudaMemsetAsync(dw, 0, bytes, st); if (e != cudaSuccess) return (int)e;
  }
  if (impl == KS_IMPL_TC || (impl == KS_IMPL_AUTO && dtype == KS_BF16)) {
    if (dtype != KS_BF16) return KS_EUNSUPPORTED;
    rc = wgrad_tc(N, H, W, ksize, xl, yl, dw, st);
    if (rc != KS_EUNSUPPORTED || impl == KS_IMPL_TC) return rc;
  }
  return wgrad_simt(dtype, N, H, W, ksize, xl, yl, dw, st);
}

extern "C" int ks_channel_sum(int dtype, int N, int H, int W, const ks_view_t *x, float *out, int accumulate, void *stream);

// Weight gradient + bias gradient of the same layer: dbias[(cstart_of_view + c) % bias_mod] (+)= sum_pixels dy[p][c] (bias_mod = 0: no
// folding).  nn.ConvTranspose2d(k2, s2) as four 1x1 phases has ONE bias per output channel for all four phase views: bias_mod = C.
extern "C" int ks_conv2d_wgrad_bias(int dtype, int N, int H, int W, int ksize, const ks_view_t *xs, int n_x, const ks_view_t *dys, int n_dy,
                                    float *dw, int accumulate, float *dbias, int bias_mod, int accumulate_bias, int impl, void *stream) {
  KS_CHECK_ARG(N > 0 && H > 0 && W > 0 && (ksize == 1 || ksize == 3) && dw && dbias && bias_mod >= 0);
  KS_CHECK_ARG(dtype == KS_F32 || dtype == KS_BF16);
  ViewList xl, yl;
  int rc = make_view_list(xs, n_x, xl); if (rc) return rc;
  rc = make_view_list(dys, n_dy, yl); if (rc) return rc;
  const int cout = yl.cstart[yl.n], blen = bias_mod > 0 ? bias_mod : cout;
  for (int i = 0; i < yl.n; ++i)
    if (bias_mod > 0 && (yl.cstart[i] % bias_mod) + yl.v[i].C > bias_mod) return KS_EINVAL;       // a view must not wrap around the fold
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if (!accumulate) { e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)ksize * ksize * xl.cstart[xl.n] * cout, st); if (e != cudaSuccess) return (int)e; }
  if (!accumulate_bias) { e = cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)blen, st); if (e != cudaSuccess) return (int)e; }
  bool done = false;
  rc = KS_EUNSUPPORTED;
  if (impl == KS_IMPL_TC || (impl == KS_IMPL_AUTO && dtype == KS_BF16)) {
    if (dtype != KS_BF16) return KS_EUNSUPPORTED;
    rc = wgrad_tc(N, H, W, ksize, xl, yl, dw, st, dbias, bias_mod, &done);
    if (rc != KS_OK && (rc != KS_EUNSUPPORTED || impl == KS_IMPL_TC)) return rc;
  }
  if (rc == KS_EUNSUPPORTED) { rc = wgrad_simt(dtype, N, H, W, ksize, xl, yl, dw, st); if (rc) return rc; done = false; }
  if (!done)
    for (int i = 0; i < n_dy; ++i) {
      rc = ks_channel_sum(dtype, N, H, W, &dys[i], dbias + (bias_mod > 0 ? yl.cstart[i] % bias_mod : yl.cstart[i]), 1, stream);
      if (rc) return rc;
    }
  return KS_OK;
}
