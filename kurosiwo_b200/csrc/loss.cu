// Fused softmax -> CE(weight, ignore_index) + Dice loss: forward, gradient and argmax in two streaming passes
// (reduce | gradient) chained by a programmatic dependent launch.
//
// Reference semantics (parity target):
//   utilities/bce_and_dice.py:18-24   loss = DiceLoss(preds, lbl) + CrossEntropyLoss(preds, lbl)
//   utilities/dice.py:111-137         target*(target!=ignore) -> one_hot(+1e-6) -> softmax ->
//                                     per-sample sums over (C,H,W) -> mean_n(1 - 2I/(S+1e-6))
//   utilities/dice.py:57-59           one_hot = zeros.scatter_(1, y, 1.0) + 1e-6   (fp32 {1e-6, 1+1e-6})
// The reference lowers this to >=12 ATen kernels and materialises an int64 one-hot
// tensor; here each logit/label is read from HBM once (phase 2 re-reads hit L2:
// 64 MB at bs=64 < 126 MB L2) and each gradient written once: 32 B/pixel.
#include "common.cuh"

namespace ks {

struct LossWs {
  // doubles: [N][2] (I_n, S_n), then [2] (ce_num, ce_den); then counters
  double *acc; double *ce; unsigned int *counter;
};

// exp / log on the MUFU unit (ex2.approx / lg2.approx, relative error ~2^-22): with expf / logf the kernel issued 23 M warp
// instructions at bs=64 (IPC 2.1 of 4 over its whole run); the loss and the gradient stay inside the 1e-5 bars of the tests.
__device__ __forceinline__ float fast_exp(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}
__device__ __forceinline__ float fast_log(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y * 0.6931471805599453f;
}

template <int C, bool NEED_LSE = true>
__device__ __forceinline__ void softmax_px(const float (&z)[C], float (&p)[C], float &m, float &lse) {
  m = z[0];
#pragma unroll
  for (int c = 1; c < C; ++c) m = fmaxf(m, z[c]);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) { p[c] = fast_exp(z[c] - m); s += p[c]; }
  float inv = __fdividef(1.0f, s);
#pragma unroll
  for (int c = 0; c < C; ++c) p[c] *= inv;
  lse = NEED_LSE ? fast_log(s) : 0.f;
}

// ---- pass 1: per-sample Dice sums, CE numerator / denominator, argmax.  Reads logits + labels once (20 B/pixel). ----
template <int C, int VEC>
__device__ __forceinline__ void load_px(const float *zb, const long long *lb, long long HW, long long px, float (&zz)[C][VEC], long long (&yy)[VEC]) {
  if (VEC == 4) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float4 t = __ldcg(reinterpret_cast<const float4 *>(zb + c * HW + px));
      zz[c][0] = t.x; zz[c][1] = t.y; zz[c][2] = t.z; zz[c][3] = t.w;
    }
    const longlong2 a = __ldcg(reinterpret_cast<const longlong2 *>(lb + px));
    const longlong2 b = __ldcg(reinterpret_cast<const longlong2 *>(lb + px + 2));
    yy[0] = a.x; yy[1] = a.y; yy[2] = b.x; yy[3] = b.y;
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) zz[c][0] = __ldcg(zb + c * HW + px);
    yy[0] = __ldcg(lb + px);
  }
}

template <int C, int VEC>
__global__ void __launch_bounds__(256)
ce_dice_reduce_kernel(const float *__restrict__ logits, const long long *__restrict__ labels,
                      int N, long long HW, const float *__restrict__ cw, int ignore_index,
                      float *__restrict__ loss_out, unsigned char *__restrict__ pred, double *acc, double *ce,
                      unsigned int *counter, float dice_weight) {
  // the gradient pass may be launched as a programmatic dependent: let it get resident (its prologue waits on griddepcontrol)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int n = blockIdx.y;
  const long long chunk = (((HW + gridDim.x - 1) / gridDim.x) + VEC - 1) / VEC * VEC;
  const long long p0 = (long long)blockIdx.x * chunk;
  const long long p1 = min(HW, p0 + chunk);
  const float *zb = logits + (long long)n * C * HW;
  const long long *lb = labels + (long long)n * HW;
  const float T0 = 1e-6f, T1 = 1.0f + 1e-6f;  // one_hot(...)+eps in fp32 (dice.py:59)
  float w[C];
#pragma unroll
  for (int c = 0; c < C; ++c) w[c] = cw[c];
  float fI = 0.f, fS = 0.f, fnum = 0.f, fden = 0.f;
  for (long long px = p0 + (long long)threadIdx.x * VEC; px < p1; px += (long long)blockDim.x * VEC) {
    float zz[C][VEC]; long long yy[VEC];
    load_px<C, VEC>(zb, lb, HW, px, zz, yy);
    unsigned char pr[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float z[C], p[C], m, lse;
#pragma unroll
      for (int c = 0; c < C; ++c) z[c] = zz[c][i];
      softmax_px<C>(z, p, m, lse);
      const int y = (int)yy[i];
      const bool valid = (y != ignore_index);
      const int yd = valid ? y : 0;  // dice.py:116-119: ignored pixels become class 0
      int am = 0; float best = z[0];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float t = (c == yd) ? T1 : T0;
        fI += p[c] * t;
        fS += p[c] + t;
        if (c > 0 && z[c] > best) { best = z[c]; am = c; }
      }
      if (valid) {
        float zy = z[0], wy = w[0];
#pragma unroll
        for (int c = 1; c < C; ++c) if (c == y) { zy = z[c]; wy = w[c]; }
        fnum += -wy * (zy - m - lse);
        fden += wy;
      }
      pr[i] = (unsigned char)am;
    }
    if (pred != nullptr) {
      if (VEC == 4) *reinterpret_cast<uchar4 *>(pred + (long long)n * HW + px) = make_uchar4(pr[0], pr[1], pr[2], pr[3]);
      else pred[(long long)n * HW + px] = pr[0];
    }
  }
  __shared__ double red[4][8];
  __shared__ unsigned int ticket;
  const double d0 = warp_sum_d((double)fI), d1 = warp_sum_d((double)fS);
  const double d2 = warp_sum_d((double)fnum), d3 = warp_sum_d((double)fden);
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][wid] = d0; red[1][wid] = d1; red[2][wid] = d2; red[3][wid] = d3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[threadIdx.x][i];
    double *dst = (threadIdx.x < 2) ? (acc + 2 * n + threadIdx.x) : (ce + (threadIdx.x - 2));
    atomicAdd(dst, s);
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) ticket = atomicAdd(counter, 1u);
  __syncthreads();
  // ---- loss value: one warp of the LAST block to finish (every other block's sums are visible: fence + atomic ticket) ----
  if (ticket == gridDim.x * gridDim.y - 1 && threadIdx.x < 32) {
    __threadfence();
    double dice = 0.0;
    for (int i = threadIdx.x; i < N; i += 32) {
      const double I = __ldcg(acc + 2 * i), S = __ldcg(acc + 2 * i + 1);
      dice += 1.0 - 2.0 * I / (S + 1e-6);
    }
    dice = warp_sum_d(dice);
    if (threadIdx.x == 0) {
      dice /= (double)N;
      const double cel = __ldcg(ce) / __ldcg(ce + 1);  // NaN when every pixel is ignored (as torch)
      loss_out[0] = (float)((double)dice_weight * dice + cel); loss_out[1] = (float)dice; loss_out[2] = (float)cel;
    }
  }
}

// ---- pass 2: gradient.  Re-reads logits + labels (L2 hits: 64 MB at bs=64 against 126 MB of L2), writes 12 B/pixel. ----
template <int C, int VEC>
__global__ void __launch_bounds__(256)
ce_dice_grad_kernel(const float *__restrict__ logits, const long long *__restrict__ labels, int N, long long HW,
                    const float *__restrict__ cw, int ignore_index, float grad_scale, float *__restrict__ dlogits,
                    const double *acc, const double *ce, float dice_weight) {
  const int n = blockIdx.y;
  const long long chunk = (((HW + gridDim.x - 1) / gridDim.x) + VEC - 1) / VEC * VEC;
  const long long p0 = (long long)blockIdx.x * chunk;
  const long long p1 = min(HW, p0 + chunk);
  const float *zb = logits + (long long)n * C * HW;
  const long long *lb = labels + (long long)n * HW;
  const float T0 = 1e-6f, T1 = 1.0f + 1e-6f;
  float w[C];
#pragma unroll
  for (int c = 0; c < C; ++c) w[c] = cw[c];
  // first tile of inputs does not depend on pass 1: issue it before waiting for the sums
  long long px = p0 + (long long)threadIdx.x * VEC;
  float zz[C][VEC]; long long yy[VEC];
  if (px < p1) load_px<C, VEC>(zb, lb, HW, px, zz, yy);
  asm volatile("griddepcontrol.wait;" ::: "memory");      // pass 1 complete and flushed (no-op without a programmatic dependency)
  const double In = __ldcg(acc + 2 * n), Sn = __ldcg(acc + 2 * n + 1) + 1e-6;
  const float a_n = (float)(-2.0 / ((double)N * Sn)) * grad_scale * dice_weight;      // dL/dp = a_n*t + b_n
  const float b_n = (float)(2.0 * In / ((double)N * Sn * Sn)) * grad_scale * dice_weight;
  const float inv_den = (float)(1.0 / __ldcg(ce + 1)) * grad_scale;
  float *gb = dlogits + (long long)n * C * HW;
  while (px < p1) {
    const long long nx = px + (long long)blockDim.x * VEC;
    float zn[C][VEC]; long long yn[VEC];
    if (nx < p1) load_px<C, VEC>(zb, lb, HW, nx, zn, yn);   // software pipeline: next tile in flight while this one computes
    float gg[C][VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float z[C], p[C], m, lse;
#pragma unroll
      for (int c = 0; c < C; ++c) z[c] = zz[c][i];
      softmax_px<C, false>(z, p, m, lse);
      const int y = (int)yy[i];
      const bool valid = (y != ignore_index);
      const int yd = valid ? y : 0;
      float g[C], dot = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        g[c] = a_n * ((c == yd) ? T1 : T0) + b_n;
        dot += g[c] * p[c];
      }
      float wy = 0.f;
      if (valid) {
        wy = w[0];
#pragma unroll
        for (int c = 1; c < C; ++c) if (c == y) wy = w[c];
        wy *= inv_den;
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float d = p[c] * (g[c] - dot);
        d += wy * (p[c] - ((valid && c == y) ? 1.0f : 0.0f));
        gg[c][i] = d;
      }
    }
    if (VEC == 4) {
#pragma unroll
      for (int c = 0; c < C; ++c)
        __stcs(reinterpret_cast<float4 *>(gb + c * HW + px), make_float4(gg[c][0], gg[c][1], gg[c][2], gg[c][3]));
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) gb[c * HW + px] = gg[c][0];
    }
    px = nx;
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int i = 0; i < VEC; ++i) zz[c][i] = zn[c][i];
#pragma unroll
    for (int i = 0; i < VEC; ++i) yy[i] = yn[i];
  }
}

template <int C, int VEC>
static int launch_ce_dice(const float *logits, const int64_t *labels, int N, int64_t HW,
                          const float *cw, int ignore_index, float grad_scale, float *loss_out,
                          float *dlogits, uint8_t *pred, void *workspace, cudaStream_t st, float dice_weight) {
  // Two plain launches (no cooperative grid, no software grid barrier: the kernel boundary is the barrier, and with the
  // programmatic-dependent-launch attribute the gradient pass is resident and has its first loads in flight when pass 1 drains).
  // grid: ~8 CTAs of 256 threads per SM, at least 1024 pixels per CTA.
  int chunks = (kNumSMs * 8 + N - 1) / N; if (chunks < 1) chunks = 1;
  long long maxchunks = (HW + 1023) / 1024; if (chunks > maxchunks) chunks = (int)maxchunks;
  if (chunks < 1) chunks = 1;
  if (g_opt.loss_chunks > 0) chunks = g_opt.loss_chunks;
  const int64_t ws_bytes = ks_ce_dice_workspace_bytes(N);
  cudaError_t e = cudaMemsetAsync(workspace, 0, (size_t)ws_bytes, st);
  if (e != cudaSuccess) return (int)e;
  double *acc = (double *)workspace; double *ce = acc + 2 * (size_t)N;
  unsigned int *counter = (unsigned int *)(ce + 2);
  const long long *lab = (const long long *)labels;
  ce_dice_reduce_kernel<C, VEC><<<dim3(chunks, N), 256, 0, st>>>(logits, lab, N, (long long)HW, cw, ignore_index, loss_out, pred, acc, ce,
                                                                  counter, dice_weight);
  e = cudaGetLastError();
  if (e != cudaSuccess || dlogits == nullptr) return (int)e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(chunks, N); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_opt.loss_no_pdl ? 0 : 1;
  e = cudaLaunchKernelEx(&cfg, ce_dice_grad_kernel<C, VEC>, logits, lab, N, (long long)HW, cw, ignore_index, grad_scale, dlogits,
                         (const double *)acc, (const double *)ce, dice_weight);
  return (int)e;
}

}  // namespace ks

extern "C" int64_t ks_ce_dice_workspace_bytes(int N) {
  return (int64_t)((2 * (int64_t)N + 2) * 8 + 64);
}

extern "C" int ks_ce_dice_fwd_bwd_ex(const float *logits, const int64_t *labels, int N, int C, int64_t HW,
                                     const float *class_weights, int ignore_index, float grad_scale, float dice_weight,
                                     float *loss_out, float *dlogits, uint8_t *pred, void *workspace, void *stream);

extern "C" int ks_ce_dice_fwd_bwd(const float *logits, const int64_t *labels, int N, int C, int64_t HW,
                                  const float *class_weights, int ignore_index, float grad_scale,
                                  float *loss_out, float *dlogits, uint8_t *pred,
                                  void *workspace, void *stream) {
  return ks_ce_dice_fwd_bwd_ex(logits, labels, N, C, HW, class_weights, ignore_index, grad_scale, 1.0f, loss_out, dlogits, pred,
                               workspace, stream);
}

extern "C" int ks_ce_dice_fwd_bwd_ex(const float *logits, const int64_t *labels, int N, int C, int64_t HW,
                                     const float *class_weights, int ignore_index, float grad_scale, float dice_weight,
                                     float *loss_out, float *dlogits, uint8_t *pred, void *workspace, void *stream) {
  KS_CHECK_ARG(logits && labels && class_weights && loss_out && workspace);
  KS_CHECK_ARG(N > 0 && HW > 0);
  if (C != 3) return KS_EUNSUPPORTED;  // num_classes is 3 on every reference config (configs/config.json:13)
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (HW % 4 == 0) && (((uintptr_t)logits & 15) == 0) && (((uintptr_t)labels & 15) == 0) &&
                   (dlogits == nullptr || ((uintptr_t)dlogits & 15) == 0) && (pred == nullptr || ((uintptr_t)pred & 3) == 0);
  if (vec) return ks::launch_ce_dice<3, 4>(logits, labels, N, HW, class_weights, ignore_index, grad_scale, loss_out, dlogits, pred, workspace, st, dice_weight);
  return ks::launch_ce_dice<3, 1>(logits, labels, N, HW, class_weights, ignore_index, grad_scale, loss_out, dlogits, pred, workspace, st, dice_weight);
}

// ---- in-loop metrics ----------------------------------------------------------------------
// The reference keeps 5 torchmetrics objects (multiclass, num_classes 4, ignore_index 3, average 'none'; utilities/utilities.py:
// 228-265) and updates each one every iteration (training/change_detection_trainer.py:184-199): ~25 small kernels per step.
// Everything they report derives from ONE KxK confusion matrix: mat[target][pred] += 1 over the pixels whose target is not ignored.
namespace ks {
template <int K>
__global__ void __launch_bounds__(256)
confusion_kernel(const unsigned char *__restrict__ pred, const long long *__restrict__ labels, long long n, int ignore_index,
                 unsigned long long *mat) {
  __shared__ unsigned int sm[K * K];
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) sm[i] = 0u;
  __syncthreads();
  unsigned int cnt[K * K];
#pragma unroll
  for (int i = 0; i < K * K; ++i) cnt[i] = 0u;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long t = labels[i];
    const int p = pred[i];
    if (t != ignore_index && t >= 0 && t < K && p < K) {
      const int idx = (int)t * K + p;
#pragma unroll
      for (int q = 0; q < K * K; ++q) cnt[q] += (q == idx) ? 1u : 0u;
    }
  }
#pragma unroll
  for (int q = 0; q < K * K; ++q) {
    unsigned int v = cnt[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sm[q], v);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * K; i += blockDim.x)
    if (sm[i]) atomicAdd(mat + i, (unsigned long long)sm[i]);
}
}  // namespace ks

namespace ks {
// Grouped variant: the batch's per-sample keys route every sample's counts to up to two extra families of matrices (per
// activation / AOI and per climate zone: change_detection_trainer.py:184-199, :445-472) next to the global one - ONE launch
// instead of 5 torchmetrics updates per (family, key present in the batch).  grid = (chunks, samples): a block sees one sample.
template <int K>
__global__ void __launch_bounds__(256)
confusion_grouped_kernel(const unsigned char *__restrict__ pred, const long long *__restrict__ labels, long long per_sample,
                         int ignore_index, const int *__restrict__ key_a, int n_a, const int *__restrict__ key_b, int n_b,
                         unsigned long long *mat, unsigned long long *mat_a, unsigned long long *mat_b) {
  __shared__ unsigned int sm[K * K];
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) sm[i] = 0u;
  __syncthreads();
  const long long base = (long long)blockIdx.y * per_sample;
  unsigned int cnt[K * K];
#pragma unroll
  for (int i = 0; i < K * K; ++i) cnt[i] = 0u;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_sample; i += (long long)gridDim.x * blockDim.x) {
    const long long t = labels[base + i];
    const int p = pred[base + i];
    if (t != ignore_index && t >= 0 && t < K && p < K) {
      const int idx = (int)t * K + p;
#pragma unroll
      for (int q = 0; q < K * K; ++q) cnt[q] += (q == idx) ? 1u : 0u;
    }
  }
#pragma unroll
  for (int q = 0; q < K * K; ++q) {
    unsigned int v = cnt[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sm[q], v);
  }
  __syncthreads();
  const int ka = key_a ? key_a[blockIdx.y] : -1, kb = key_b ? key_b[blockIdx.y] : -1;
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) {
    const unsigned long long v = sm[i];
    if (!v) continue;
    if (mat) atomicAdd(mat + i, v);
    if (mat_a && ka >= 0 && ka < n_a) atomicAdd(mat_a + (size_t)ka * K * K + i, v);
    if (mat_b && kb >= 0 && kb < n_b) atomicAdd(mat_b + (size_t)kb * K * K + i, v);
  }
}
}  // namespace ks

extern "C" int ks_confusion_update_grouped(const uint8_t *pred, const int64_t *labels, int n_samples, int64_t per_sample,
                                           int num_classes_with_ignore, int ignore_index, const int32_t *key_a, int n_a,
                                           const int32_t *key_b, int n_b, int64_t *mat, int64_t *mat_a, int64_t *mat_b, void *stream) {
  KS_CHECK_ARG(pred && labels && n_samples > 0 && per_sample > 0 && (mat || mat_a || mat_b));
  KS_CHECK_ARG((!mat_a || (key_a && n_a > 0)) && (!mat_b || (key_b && n_b > 0)));
  if (num_classes_with_ignore != 4) return KS_EUNSUPPORTED;
  long long g = (per_sample + 256 * 16 - 1) / (256 * 16);
  const long long cap = (ks::kNumSMs * 8 + n_samples - 1) / n_samples;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  dim3 grid((unsigned)g, (unsigned)n_samples);
  ks::confusion_grouped_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(pred, (const long long *)labels, per_sample, ignore_index, key_a, n_a,
                                                                            key_b, n_b, (unsigned long long *)mat, (unsigned long long *)mat_a,
                                                                            (unsigned long long *)mat_b);
  KS_LAUNCH_RET();
}

extern "C" int ks_confusion_update(const uint8_t *pred, const int64_t *labels, int64_t n, int num_classes_with_ignore, int ignore_index,
                                   int64_t *mat, void *stream) {
  KS_CHECK_ARG(pred && labels && mat && n > 0);
  if (num_classes_with_ignore != 4) return KS_EUNSUPPORTED;
  long long g = (n + 256 * 16 - 1) / (256 * 16);
  if (g > ks::kNumSMs * 8) g = ks::kNumSMs * 8;
  if (g < 1) g = 1;
  ks::confusion_kernel<4><<<(int)g, 256, 0, (cudaStream_t)stream>>>(pred, (const long long *)labels, n, ignore_index, (unsigned long long *)mat);
  KS_LAUNCH_RET();
}
