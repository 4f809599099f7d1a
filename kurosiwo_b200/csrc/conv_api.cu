// C-ABI entry points of the convolution engine: argument checks + dispatch between the
// tcgen05 engine (conv_tc.cu / wgrad_tc.cu, bf16) and the CUDA-core engine (conv_simt.cu).
#include "common.cuh"

namespace ks {
int conv2d_simt(int dtype, int N, int H, int W, int ksize, const ViewList &srcs, const void *weight, const float *bias,
                const ViewList &dsts, int acc_mask, double *stats, cudaStream_t st);
int wgrad_simt(int dtype, int N, int H, int W, int ksize, const ViewList &xs, const ViewList &dys, float *dw, cudaStream_t st);
int conv2d_tc(int N, int H, int W, int ksize, const ViewList &srcs, const void *weight, const float *bias,
              const ViewList &dsts, int acc_mask, double *stats, cudaStream_t st);
int wgrad_tc(int N, int H, int W, int ksize, const ViewList &xs, const ViewList &dys, float *dw, cudaStream_t st);
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
