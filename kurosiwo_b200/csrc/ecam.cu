// ECAM head of SNUNet (ensemble channel attention + 1x1 classifier), forward and backward.
//
// Reference: models/snunet.py:49-62 (ChannelAttention), :146-151
//   out   = cat(x0_1..x0_4)                      (J*Cb channels, never materialised here)
//   intra = x0_1 + x0_2 + x0_3 + x0_4            (Cb channels, never materialised here)
//   ca    = CA_{J*Cb}(out), ca1 = CA_{Cb}(intra) (global avg+max pool -> fc1 -> relu -> fc2 -> sigmoid)
//   logits = conv_final( ca * (out + ca1.repeat(1,J,1,1)) )
// The reference runs ~14 ATen kernels and materialises `out`, `intra`, the repeat and the
// gated tensor (4 x 128ch x 224^2 per sample).  Here: one pooling pass, one tiny gate kernel
// and one pass that reads the four block outputs and writes the 3 logit planes.
#include "common.cuh"

namespace ks {

__device__ __forceinline__ unsigned long long pack_max(float v, unsigned int idx) {
  unsigned int b = __float_as_uint(v);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)b << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ float unpack_max_val(unsigned long long k) {
  unsigned int b = (unsigned int)(k >> 32);
  b = (b & 0x80000000u) ? (b & 0x7FFFFFFFu) : ~b;
  return __uint_as_float(b);
}

template <typename T>
__device__ __forceinline__ const T *vptr(const View &v, int n, int h, int w, int c) {
  return reinterpret_cast<const T *>(v.ptr) + ((long long)n * v.sn + (long long)h * v.sh + (long long)w * v.sw + c);
}

constexpr int kMaxJ = 4;

// ---- pooling: per-(n,c) sum and max of the J views and of intra = sum_j x_j -------------------------------
// Thread = (pixel lane, 8-channel vector) and handles ALL J views of its channels, so the intra sum is thread-local
// (no shuffles).  Only VALUES are tracked here (80 registers of state); the index of the first maximum - needed by
// the backward pass - is recovered by ecam_final, which reads the same tensors anyway (equality test + atomicMin).
__device__ __forceinline__ unsigned int fkey(float v) {          // order-preserving float -> uint
  const unsigned int b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

template <typename T>
__global__ void __launch_bounds__(256)
ecam_pool_kernel(ViewList xs, int J, int Cb, int H, int W, float *pooled, unsigned int *maxkey) {
  extern __shared__ unsigned char smem_raw[];
  const int CC = J * Cb, CT = (J + 1) * Cb;
  unsigned int *smax = reinterpret_cast<unsigned int *>(smem_raw);
  float *ssum = reinterpret_cast<float *>(smax + CT);
  for (int i = threadIdx.x; i < CT; i += blockDim.x) { smax[i] = 0u; ssum[i] = 0.f; }
  __syncthreads();
  const int n = blockIdx.y;
  const int CVb = Cb / 8, rows = blockDim.x / CVb;
  const int tx = threadIdx.x % CVb, ty = threadIdx.x / CVb;
  const int HW = H * W;
  float sum[kMaxJ][8], best[kMaxJ + 1][8];
#pragma unroll
  for (int j = 0; j <= kMaxJ; ++j)
#pragma unroll
    for (int k = 0; k < 8; ++k) { if (j < kMaxJ) sum[j][k] = 0.f; best[j][k] = -INFINITY; }
  for (int p = blockIdx.x * rows + ty; p < HW; p += gridDim.x * rows) {
    const int h = p / W, w = p % W;
    float f[kMaxJ][8];
#pragma unroll
    for (int j = 0; j < kMaxJ; ++j)
      if (j < J) ld8(vptr<T>(xs.v[j], n, h, w, tx * 8), f[j]);
    float it[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) it[k] = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxJ; ++j)
      if (j < J) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { sum[j][k] += f[j][k]; best[j][k] = fmaxf(best[j][k], f[j][k]); it[k] += f[j][k]; }
      }
#pragma unroll
    for (int k = 0; k < 8; ++k) best[kMaxJ][k] = fmaxf(best[kMaxJ][k], it[k]);
  }
  // combine the pixel lanes of a warp (lanes sharing tx), then one smem atomic per warp and channel
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j <= kMaxJ; ++j) {
    if (j < J || j == kMaxJ) {
      const int jj = (j == kMaxJ) ? J : j;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float mv = best[j][k];
        float sv = (j < kMaxJ) ? sum[j][k] : 0.f;
        for (int o = CVb; o < 32; o <<= 1) {
          mv = fmaxf(mv, __shfl_xor_sync(0xffffffffu, mv, o));
          sv += __shfl_xor_sync(0xffffffffu, sv, o);
        }
        if (lane < CVb) {
          const int c = jj * Cb + tx * 8 + k;
          atomicMax(&smax[c], fkey(mv));
          if (j < kMaxJ) { atomicAdd(&ssum[c], sv); atomicAdd(&ssum[CC + tx * 8 + k], sv); }   // avg(intra) = sum_j avg_j
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < CT; i += blockDim.x) {
    atomicAdd(pooled + ((size_t)n * 2 + 0) * CT + i, ssum[i]);
    atomicMax(maxkey + (size_t)n * CT + i, smax[i]);
  }
}

__global__ void ecam_pool_finalize_kernel(int N, int CT, int HW, float *pooled, const unsigned int *maxkey, int *argmax) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * CT) return;
  const int n = i / CT, c = i % CT;
  pooled[((size_t)n * 2 + 0) * CT + c] *= (1.0f / (float)HW);
  pooled[((size_t)n * 2 + 1) * CT + c] = fkey_inv(maxkey[i]);
  argmax[i] = 0x7fffffff;       // filled by ecam_final (first pixel whose value equals the maximum)
}

// ---- gates: one block per sample -----------------------------------------------------------
__device__ __forceinline__ void ca_forward(int Cin, int hid, const float *avg, const float *mx,
                                           const float *w1 /*[hid][Cin]*/, const float *w2 /*[Cin][hid]*/,
                                           float *gate, float *h_avg, float *h_max, float *sh /* smem 2*hid */) {
  for (int q = threadIdx.x; q < 2 * hid; q += blockDim.x) {
    const int qq = q % hid; const float *src = (q < hid) ? avg : mx;
    float s = 0.f;
    for (int c = 0; c < Cin; ++c) s += w1[qq * Cin + c] * src[c];
    sh[q] = s;
    if (q < hid) h_avg[qq] = s; else h_max[qq] = s;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cin; c += blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < hid; ++q) s += w2[c * hid + q] * (fmaxf(sh[q], 0.f) + fmaxf(sh[hid + q], 0.f));
    gate[c] = 1.0f / (1.0f + expf(-s));
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
ecam_gates_kernel(int Cb, int J, int hid, int hid1, const float *__restrict__ pooled,
                  const float *__restrict__ w_fc1, const float *__restrict__ w_fc2,
                  const float *__restrict__ w1_fc1, const float *__restrict__ w1_fc2,
                  float *gates, float *hidden) {
  __shared__ float sh[128];
  const int n = blockIdx.x, CT = (J + 1) * Cb, CC = J * Cb, HT = hid + hid1;
  const float *avg = pooled + ((size_t)n * 2 + 0) * CT, *mx = pooled + ((size_t)n * 2 + 1) * CT;
  float *g = gates + (size_t)n * CT;
  float *ha = hidden + ((size_t)n * 2 + 0) * HT, *hm = hidden + ((size_t)n * 2 + 1) * HT;
  ca_forward(CC, hid, avg, mx, w_fc1, w_fc2, g, ha, hm, sh);
  ca_forward(Cb, hid1, avg + CC, mx + CC, w1_fc1, w1_fc2, g + CC, ha + hid, hm + hid, sh);
}

// ---- final: gated sum + 1x1 classifier -> NCHW fp32 logits (+ arg-max discovery) -----------------------------
// Thread = (pixel lane, 8-channel vector), all J views: 4 x 16-byte loads, effective weights ca*wf from shared memory
// (LDS.128, conflict-free: the 4 vector lanes read 4 adjacent 32-byte runs), partial logits reduced over the CVb vector
// lanes.  While the values are in registers, any element equal to its channel's pooled maximum records its pixel with
// atomicMin -> argmax[n][c] = FIRST maximum (what aten adaptive_max_pool2d's backward routes to).
template <typename T, int K>
__global__ void __launch_bounds__(256)
ecam_final_kernel(ViewList xs, int J, int Cb, int H, int W, const float *__restrict__ gates,
                  const float *__restrict__ wf, const float *__restrict__ bf, const float *__restrict__ pooled,
                  int *argmax, float *logits) {
  extern __shared__ float sm[];
  const int n = blockIdx.y, CC = J * Cb, CT = (J + 1) * Cb, HW = H * W;
  float *weff = sm;                 // [J][K][Cb]   ca*wf, view-major so that a thread's weights are contiguous runs
  float *cst = weff + K * CC;       // [K]
  float *smx = cst + 4;             // [CT] pooled maxima of this sample
  const float *g = gates + (size_t)n * CT;
  for (int i = threadIdx.x; i < K * CC; i += blockDim.x) {
    const int j = i / (K * Cb), k = (i / Cb) % K, cb = i % Cb;
    weff[i] = wf[k * CC + j * Cb + cb] * g[j * Cb + cb];
  }
  for (int i = threadIdx.x; i < CT; i += blockDim.x) smx[i] = pooled ? pooled[((size_t)n * 2 + 1) * CT + i] : 0.f;
  __syncthreads();
  if (threadIdx.x < K) {
    float s = bf[threadIdx.x];
    for (int j = 0; j < J; ++j)
      for (int cb = 0; cb < Cb; ++cb) s += weff[(j * K + threadIdx.x) * Cb + cb] * g[CC + cb];
    cst[threadIdx.x] = s;
  }
  __syncthreads();
  const int CVb = Cb / 8, rows = blockDim.x / CVb;
  const int tx = threadIdx.x % CVb, ty = threadIdx.x / CVb;
  int *am = argmax ? argmax + (size_t)n * CT : nullptr;
  float mx[kMaxJ + 1][8];           // pooled maxima of this thread's channels (register-resident)
#pragma unroll
  for (int j = 0; j <= kMaxJ; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) mx[j][i] = (j < J) ? smx[j * Cb + tx * 8 + i] : (j == kMaxJ ? smx[CC + tx * 8 + i] : 0.f);
  const int pmax = ((HW + rows - 1) / rows) * rows;
  for (int p = blockIdx.x * rows + ty; p < pmax; p += gridDim.x * rows) {
    const bool ok = p < HW;
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.f;
    if (ok) {
      const int h = p / W, w = p % W;
      float f[kMaxJ][8], it[8];
#pragma unroll
      for (int j = 0; j < kMaxJ; ++j)
        if (j < J) ld8(vptr<T>(xs.v[j], n, h, w, tx * 8), f[j]);
#pragma unroll
      for (int i = 0; i < 8; ++i) it[i] = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxJ; ++j)
        if (j < J) {
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const float4 w0 = *reinterpret_cast<const float4 *>(weff + (j * K + k) * Cb + tx * 8);
            const float4 w1 = *reinterpret_cast<const float4 *>(weff + (j * K + k) * Cb + tx * 8 + 4);
            acc[k] = fmaf(f[j][0], w0.x, acc[k]); acc[k] = fmaf(f[j][1], w0.y, acc[k]);
            acc[k] = fmaf(f[j][2], w0.z, acc[k]); acc[k] = fmaf(f[j][3], w0.w, acc[k]);
            acc[k] = fmaf(f[j][4], w1.x, acc[k]); acc[k] = fmaf(f[j][5], w1.y, acc[k]);
            acc[k] = fmaf(f[j][6], w1.z, acc[k]); acc[k] = fmaf(f[j][7], w1.w, acc[k]);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            it[i] += f[j][i];
            if (am && f[j][i] == mx[j][i]) atomicMin(am + j * Cb + tx * 8 + i, p);
          }
        }
      if (am) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (it[i] == mx[kMaxJ][i]) atomicMin(am + CC + tx * 8 + i, p);
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k)
      for (int o = 1; o < CVb; o <<= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (ok) {
#pragma unroll
      for (int k = 0; k < K; ++k)
        if ((k % CVb) == tx) logits[((size_t)n * K + k) * HW + p] = acc[k] + cst[k];
    }
  }
}

// ---- backward reductions: B[n][k][c] = sum_px dl[k]*x[c],  D[n][k] = sum_px dl[k] ---------
template <typename T, int K>
__global__ void __launch_bounds__(256)
ecam_bwd_reduce_kernel(ViewList xs, int J, int Cb, int H, int W, const float *__restrict__ dlogits, double *red) {
  extern __shared__ float sacc[];  // [K*CC + K]
  const int n = blockIdx.y, CC = J * Cb, HW = H * W, RT = K * CC + K;
  for (int i = threadIdx.x; i < RT; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  // warp -> view j (a warp load covers 32/CVb consecutive pixels of ONE view: 512 contiguous bytes at Cb = 32);
  // lane -> (pixel lane, 8-channel vector)
  const int CVb = Cb / 8, ppw = 32 / CVb;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = warp % J, tx = lane % CVb;
  const int rows = (blockDim.x >> 5) / J * ppw, ty = (warp / J) * ppw + lane / CVb;
  const int cb0 = tx * 8, c0 = j * Cb + cb0;
  const View &xv = xs.v[j];
  const T *xp = reinterpret_cast<const T *>(xv.ptr) + (long long)n * xv.sn + cb0;
  float B[K][8], D[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    D[k] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) B[k][i] = 0.f;
  }
  constexpr int UNR = 4;
  const int stride = gridDim.x * rows;
  for (int p0 = blockIdx.x * rows + ty; p0 < HW; p0 += stride * UNR) {
    float f[UNR][8], dl[UNR][K];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int p = p0 + u * stride;
      if (p < HW) {
        ld8(xp + (long long)(p / W) * xv.sh + (long long)(p % W) * xv.sw, f[u]);
#pragma unroll
        for (int k = 0; k < K; ++k) dl[u][k] = __ldg(dlogits + ((size_t)n * K + k) * HW + p);
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) dl[u][k] = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) f[u][i] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u)
#pragma unroll
      for (int k = 0; k < K; ++k) {
        D[k] += dl[u][k];
#pragma unroll
        for (int i = 0; i < 8; ++i) B[k][i] = fmaf(dl[u][k], f[u][i], B[k][i]);
      }
  }
  // the pixel lanes of a warp (lanes sharing tx) are combined by shuffles, then one smem atomic per warp and value
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = B[k][i];
      for (int o = CVb; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane < CVb) atomicAdd(&sacc[k * CC + c0 + i], v);
    }
    float d = D[k];
    for (int o = CVb; o < 32; o <<= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0 && j == 0) atomicAdd(&sacc[K * CC + k], d);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < RT; i += blockDim.x) atomicAdd(red + (size_t)n * RT + i, (double)sacc[i]);
}

// ---- gates backward: one block per sample; weight gradients are combined with fp32 atomics ----------
__global__ void __launch_bounds__(256)
ecam_gates_bwd_kernel(int N, int Cb, int J, int hid, int hid1, int K,
                      const float *__restrict__ pooled, const float *__restrict__ hidden,
                      const float *__restrict__ gates, const double *__restrict__ red,
                      const float *__restrict__ wf, const float *__restrict__ w_fc1, const float *__restrict__ w_fc2,
                      const float *__restrict__ w1_fc1, const float *__restrict__ w1_fc2,
                      float *dpooled, float *dwf, float *dbf, float *dw_fc1, float *dw_fc2,
                      float *dw1_fc1, float *dw1_fc2) {
  extern __shared__ float sm[];
  const int CC = J * Cb, CT = (J + 1) * Cb, HT = hid + hid1, RT = K * CC + K;
  float *ds = sm;            // CT   (d pre-sigmoid)
  float *dh = ds + CT;       // 2*HT (d hidden pre-relu: avg | max)
  const int n = blockIdx.x;
  const double *R = red + (size_t)n * RT;
  const float *g = gates + (size_t)n * CT;
  const float *avg = pooled + ((size_t)n * 2 + 0) * CT, *mx = pooled + ((size_t)n * 2 + 1) * CT;
  const float *ha = hidden + ((size_t)n * 2 + 0) * HT, *hm = hidden + ((size_t)n * 2 + 1) * HT;
  for (int c = threadIdx.x; c < CC; c += blockDim.x) {
    const float ca1 = g[CC + (c % Cb)];
    float dca = 0.f;
    for (int k = 0; k < K; ++k) {
      const float A = (float)R[k * CC + c] + ca1 * (float)R[K * CC + k];
      dca += wf[k * CC + c] * A;
      if (dwf) atomicAdd(dwf + k * CC + c, g[c] * A);
    }
    ds[c] = dca * g[c] * (1.f - g[c]);
  }
  for (int cb = threadIdx.x; cb < Cb; cb += blockDim.x) {
    float d = 0.f;
    for (int j = 0; j < J; ++j) {
      const int c = j * Cb + cb;
      float t = 0.f;
      for (int k = 0; k < K; ++k) t += wf[k * CC + c] * (float)R[K * CC + k];
      d += g[c] * t;
    }
    const float gg = g[CC + cb];
    ds[CC + cb] = d * gg * (1.f - gg);
  }
  if (threadIdx.x < K && dbf) atomicAdd(dbf + threadIdx.x, (float)R[K * CC + threadIdx.x]);
  __syncthreads();
  for (int q = threadIdx.x; q < HT; q += blockDim.x) {
    float s = 0.f;
    if (q < hid) { for (int c = 0; c < CC; ++c) s += ds[c] * w_fc2[c * hid + q]; }
    else { const int qq = q - hid; for (int c = 0; c < Cb; ++c) s += ds[CC + c] * w1_fc2[c * hid1 + qq]; }
    dh[q] = (ha[q] > 0.f) ? s : 0.f;
    dh[HT + q] = (hm[q] > 0.f) ? s : 0.f;
  }
  if (dw_fc2) for (int i = threadIdx.x; i < CC * hid; i += blockDim.x) {
    const int c = i / hid, q = i % hid;
    atomicAdd(dw_fc2 + i, ds[c] * (fmaxf(ha[q], 0.f) + fmaxf(hm[q], 0.f)));
  }
  if (dw1_fc2) for (int i = threadIdx.x; i < Cb * hid1; i += blockDim.x) {
    const int c = i / hid1, q = i % hid1;
    atomicAdd(dw1_fc2 + i, ds[CC + c] * (fmaxf(ha[hid + q], 0.f) + fmaxf(hm[hid + q], 0.f)));
  }
  __syncthreads();
  if (dw_fc1) for (int i = threadIdx.x; i < hid * CC; i += blockDim.x) {
    const int q = i / CC, c = i % CC;
    atomicAdd(dw_fc1 + i, dh[q] * avg[c] + dh[HT + q] * mx[c]);
  }
  if (dw1_fc1) for (int i = threadIdx.x; i < hid1 * Cb; i += blockDim.x) {
    const int q = i / Cb, c = i % Cb;
    atomicAdd(dw1_fc1 + i, dh[hid + q] * avg[CC + c] + dh[HT + hid + q] * mx[CC + c]);
  }
  for (int c = threadIdx.x; c < CT; c += blockDim.x) {
    float da = 0.f, dm = 0.f;
    if (c < CC) { for (int q = 0; q < hid; ++q) { da += dh[q] * w_fc1[q * CC + c]; dm += dh[HT + q] * w_fc1[q * CC + c]; } }
    else { const int cb = c - CC; for (int q = 0; q < hid1; ++q) { da += dh[hid + q] * w1_fc1[q * Cb + cb]; dm += dh[HT + hid + q] * w1_fc1[q * Cb + cb]; } }
    dpooled[((size_t)n * 2 + 0) * CT + c] = da;
    dpooled[((size_t)n * 2 + 1) * CT + c] = dm;
  }
}

// ---- backward apply: gradient of the four block outputs -------------------------------------
// Thread = (pixel lane, 8-channel group): all per-channel coefficients (ca*wf, pool gradients, argmax pixels) in registers.
template <typename T, int K>
__global__ void __launch_bounds__(256)
ecam_bwd_apply_kernel(ViewList dxs, int J, int Cb, int H, int W, const float *__restrict__ gates,
                      const float *__restrict__ wf, const float *__restrict__ dlogits,
                      const float *__restrict__ dpooled, const int *__restrict__ argmax) {
  const int n = blockIdx.y, CC = J * Cb, CT = (J + 1) * Cb, HW = H * W;
  const int CVb = Cb / 8, ppw = 32 / CVb;               // warp -> view, lane -> (pixel lane, 8-channel vector)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = warp % J, tx = lane % CVb;
  const int rows = (blockDim.x >> 5) / J * ppw, ty = (warp / J) * ppw + lane / CVb;
  const int cb0 = tx * 8, c0 = j * Cb + cb0;
  const float *gt = gates + (size_t)n * CT;
  const float *da = dpooled + ((size_t)n * 2 + 0) * CT, *dm = dpooled + ((size_t)n * 2 + 1) * CT;
  const int *am = argmax + (size_t)n * CT;
  const float inv = 1.0f / (float)HW;
  float we[K][8], base[8], dmc[8], dmi[8]; int amc[8], ami[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int k = 0; k < K; ++k) we[k][i] = wf[k * CC + c0 + i] * gt[c0 + i];
    base[i] = (da[c0 + i] + da[CC + cb0 + i]) * inv;
    dmc[i] = dm[c0 + i]; amc[i] = am[c0 + i];
    dmi[i] = dm[CC + cb0 + i]; ami[i] = am[CC + cb0 + i];
  }
  const View &dv = dxs.v[j];
  T *dp = reinterpret_cast<T *>(dv.ptr) + (long long)n * dv.sn + cb0;
  const bool flat = dv.sh == (long long)W * dv.sw;      // pixel-dense view (the planar slots): no h / w decomposition per pixel
  constexpr int UNR = 4;
  const int stride = gridDim.x * rows;
  for (int p0 = blockIdx.x * rows + ty; p0 < HW; p0 += stride * UNR) {
    float dl[UNR][K];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int p = min(p0 + u * stride, HW - 1);
#pragma unroll
      for (int k = 0; k < K; ++k) dl[u][k] = __ldg(dlogits + ((size_t)n * K + k) * HW + p);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int p = p0 + u * stride;
      if (p < HW) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float v = base[i];
#pragma unroll
          for (int k = 0; k < K; ++k) v = fmaf(we[k][i], dl[u][k], v);
          if (amc[i] == p) v += dmc[i];
          if (ami[i] == p) v += dmi[i];
          o[i] = v;
        }
        st8(dp + (flat ? (long long)p * dv.sw : (long long)(p / W) * dv.sh + (long long)(p % W) * dv.sw), o);
      }
    }
  }
}


// =====================================================================================================================
// Bulk-staged variants (default when every view is pixel-dense: sw == Cb, sh == W*Cb, the planar slot layout).
// The register-resident formulations above run at 2 CTAs x 8 warps per SM with 4 x 16-byte loads in flight per thread:
// ~32 KB per SM, i.e. latency-bound at ~1.1-1.6 TB/s (ncu: 80 % of the samples in long-scoreboard stalls).  Here one thread per CTA
// issues cp.async.bulk copies of whole pixel chunks (P pixels x Cb channels of all J views = up to 32 KB) into a 3-stage shared-memory
// ring guarded by mbarriers, so ~64-96 KB per CTA are in flight regardless of register pressure; the arithmetic is unchanged.
// =====================================================================================================================
constexpr int BULK_STAGES = 3;

__device__ __forceinline__ uint32_t bk_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bk_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bk_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bk_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bk_copy(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Stage layout: [J][P][Cb] elements of T, then (optionally) [K][P] floats of dlogits.
struct BulkPipe {
  uint32_t bars;        // smem address of BULK_STAGES mbarriers
  uint32_t stage0;      // smem address of stage 0
  uint32_t stage_bytes;
  __device__ __forceinline__ void init(unsigned char *barmem, unsigned char *stagemem, uint32_t sb) {
    bars = bk_smem_u32(barmem); stage0 = bk_smem_u32(stagemem); stage_bytes = sb;
    if (threadIdx.x == 0) {
      for (int i = 0; i < BULK_STAGES; ++i) bk_mbar_init(bars + 8u * i, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
};

template <typename T>
__device__ __forceinline__ void bulk_issue(const BulkPipe &bp, int stage, const ViewList &xs, int J, int Cb, int n, int P, int p0, int npx,
                                           const float *dl, int K, long long HW) {
  // one thread: expect the bytes, then one copy per view (+ one per logit plane)
  const uint32_t bar = bp.bars + 8u * stage, dst0 = bp.stage0 + (uint32_t)stage * bp.stage_bytes;
  const uint32_t vb = (uint32_t)npx * Cb * sizeof(T);
  const uint32_t lb = dl ? (uint32_t)npx * 4u : 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  bk_expect_tx(bar, vb * J + lb * K);
  for (int j = 0; j < J; ++j) {
    const View &v = xs.v[j];
    bk_copy(dst0 + (uint32_t)j * P * Cb * sizeof(T), reinterpret_cast<const T *>(v.ptr) + ((long long)n * v.sn + (long long)p0 * Cb), vb, bar);
  }
  if (dl)
    for (int k = 0; k < K; ++k)
      bk_copy(dst0 + (uint32_t)J * P * Cb * sizeof(T) + (uint32_t)k * P * 4u, dl + ((long long)n * K + k) * HW + p0, lb, bar);
}

template <typename T>
__global__ void __launch_bounds__(256, 2)
ecam_pool_bulk_kernel(ViewList xs, int J, int Cb, int HW, int P, float *pooled, unsigned int *maxkey) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int CC = J * Cb, CT = (J + 1) * Cb;
  const uint32_t stage_bytes = (uint32_t)J * P * Cb * sizeof(T);
  unsigned char *stages = smraw;
  unsigned char *barmem = smraw + BULK_STAGES * stage_bytes;
  unsigned int *smax = reinterpret_cast<unsigned int *>(barmem + 64);
  float *ssum = reinterpret_cast<float *>(smax + CT);
  BulkPipe bp; bp.init(barmem, stages, stage_bytes);
  for (int i = threadIdx.x; i < CT; i += blockDim.x) { smax[i] = 0u; ssum[i] = 0.f; }
  __syncthreads();
  const int n = blockIdx.y, CVb = Cb / 8, tx = threadIdx.x % CVb;
  const int nchunks = (HW + P - 1) / P;
  float sum[kMaxJ][8], best[kMaxJ + 1][8];
#pragma unroll
  for (int j = 0; j <= kMaxJ; ++j)
#pragma unroll
    for (int k = 0; k < 8; ++k) { if (j < kMaxJ) sum[j][k] = 0.f; best[j][k] = -INFINITY; }
  if (threadIdx.x == 0)
    for (int s = 0; s < BULK_STAGES - 1; ++s) {
      const int c = blockIdx.x + s * gridDim.x;
      if (c < nchunks) bulk_issue<T>(bp, s, xs, J, Cb, n, P, c * P, min(P, HW - c * P), nullptr, 0, HW);
    }
  int it = 0;
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x, ++it) {
    const int s = it % BULK_STAGES;
    if (threadIdx.x == 0) {
      const int cn = c + (BULK_STAGES - 1) * gridDim.x;
      if (cn < nchunks) bulk_issue<T>(bp, (it + BULK_STAGES - 1) % BULK_STAGES, xs, J, Cb, n, P, cn * P, min(P, HW - cn * P), nullptr, 0, HW);
    }
    bk_wait(bp.bars + 8u * s, (uint32_t)((it / BULK_STAGES) & 1));
    const T *st = reinterpret_cast<const T *>(stages + (size_t)s * stage_bytes);
    const int npx = min(P, HW - c * P);
    for (int i = threadIdx.x; i < npx * CVb; i += blockDim.x) {
      const int px = i / CVb;
      float f[kMaxJ][8], itv[8];
#pragma unroll
      for (int j = 0; j < kMaxJ; ++j)
        if (j < J) ld8(st + ((size_t)j * P + px) * Cb + tx * 8, f[j]);
#pragma unroll
      for (int k = 0; k < 8; ++k) itv[k] = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxJ; ++j)
        if (j < J) {
#pragma unroll
          for (int k = 0; k < 8; ++k) { sum[j][k] += f[j][k]; best[j][k] = fmaxf(best[j][k], f[j][k]); itv[k] += f[j][k]; }
        }
#pragma unroll
      for (int k = 0; k < 8; ++k) best[kMaxJ][k] = fmaxf(best[kMaxJ][k], itv[k]);
    }
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j <= kMaxJ; ++j) {
    if (j < J || j == kMaxJ) {
      const int jj = (j == kMaxJ) ? J : j;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float mv = best[j][k];
        float sv = (j < kMaxJ) ? sum[j][k] : 0.f;
        for (int o = CVb; o < 32; o <<= 1) {
          mv = fmaxf(mv, __shfl_xor_sync(0xffffffffu, mv, o));
          sv += __shfl_xor_sync(0xffffffffu, sv, o);
        }
        if (lane < CVb) {
          const int c = jj * Cb + tx * 8 + k;
          atomicMax(&smax[c], fkey(mv));
          if (j < kMaxJ) { atomicAdd(&ssum[c], sv); atomicAdd(&ssum[CC + tx * 8 + k], sv); }
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < CT; i += blockDim.x) {
    atomicAdd(pooled + ((size_t)n * 2 + 0) * CT + i, ssum[i]);
    atomicMax(maxkey + (size_t)n * CT + i, smax[i]);
  }
}

template <typename T, int K>
__global__ void __launch_bounds__(256, 2)
ecam_final_bulk_kernel(ViewList xs, int J, int Cb, int HW, int P, const float *__restrict__ gates, const float *__restrict__ wf,
                       const float *__restrict__ bf, const float *__restrict__ pooled, int *argmax, float *logits) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int n = blockIdx.y, CC = J * Cb, CT = (J + 1) * Cb;
  const uint32_t stage_bytes = (uint32_t)J * P * Cb * sizeof(T);
  unsigned char *stages = smraw;
  unsigned char *barmem = smraw + BULK_STAGES * stage_bytes;
  float *weff = reinterpret_cast<float *>(barmem + 64);   // [J][K][Cb]
  float *cst = weff + K * CC;                              // [K] (+1 pad)
  float *smx = cst + 4;                                    // [CT]
  float *slog = smx + CT;                                  // [K][P] logits of the chunk (coalesced plane stores)
  BulkPipe bp; bp.init(barmem, stages, stage_bytes);
  const float *g = gates + (size_t)n * CT;
  for (int i = threadIdx.x; i < K * CC; i += blockDim.x) {
    const int j = i / (K * Cb), k = (i / Cb) % K, cb = i % Cb;
    weff[i] = wf[k * CC + j * Cb + cb] * g[j * Cb + cb];
  }
  for (int i = threadIdx.x; i < CT; i += blockDim.x) smx[i] = pooled ? pooled[((size_t)n * 2 + 1) * CT + i] : 0.f;
  __syncthreads();
  if (threadIdx.x < K) {
    float s = bf[threadIdx.x];
    for (int j = 0; j < J; ++j)
      for (int cb = 0; cb < Cb; ++cb) s += weff[(j * K + threadIdx.x) * Cb + cb] * g[CC + cb];
    cst[threadIdx.x] = s;
  }
  __syncthreads();
  const int CVb = Cb / 8, tx = threadIdx.x % CVb;
  int *am = argmax ? argmax + (size_t)n * CT : nullptr;
  float mx[kMaxJ + 1][8];
#pragma unroll
  for (int j = 0; j <= kMaxJ; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) mx[j][i] = (j < J) ? smx[j * Cb + tx * 8 + i] : (j == kMaxJ ? smx[CC + tx * 8 + i] : 0.f);
  const int nchunks = (HW + P - 1) / P;
  if (threadIdx.x == 0)
    for (int s = 0; s < BULK_STAGES - 1; ++s) {
      const int c = blockIdx.x + s * gridDim.x;
      if (c < nchunks) bulk_issue<T>(bp, s, xs, J, Cb, n, P, c * P, min(P, HW - c * P), nullptr, 0, HW);
    }
  int it = 0;
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x, ++it) {
    const int s = it % BULK_STAGES;
    if (threadIdx.x == 0) {
      const int cn = c + (BULK_STAGES - 1) * gridDim.x;
      if (cn < nchunks) bulk_issue<T>(bp, (it + BULK_STAGES - 1) % BULK_STAGES, xs, J, Cb, n, P, cn * P, min(P, HW - cn * P), nullptr, 0, HW);
    }
    bk_wait(bp.bars + 8u * s, (uint32_t)((it / BULK_STAGES) & 1));
    const T *st = reinterpret_cast<const T *>(stages + (size_t)s * stage_bytes);
    const int p0 = c * P, npx = min(P, HW - p0);
    const int nitems = ((npx * CVb + 31) / 32) * 32;           // whole warps run the shuffles
    for (int i = threadIdx.x; i < nitems; i += blockDim.x) {
      const int px = i / CVb;
      const bool ok = px < npx;
      const int p = p0 + px;
      float acc[K];
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] = 0.f;
      if (ok) {
        float f[kMaxJ][8], itv[8];
        bool hit = false;
#pragma unroll
        for (int j = 0; j < kMaxJ; ++j)
          if (j < J) ld8(st + ((size_t)j * P + px) * Cb + tx * 8, f[j]);
#pragma unroll
        for (int q = 0; q < 8; ++q) itv[q] = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxJ; ++j)
          if (j < J) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
              const float4 w0 = *reinterpret_cast<const float4 *>(weff + (j * K + k) * Cb + tx * 8);
              const float4 w1 = *reinterpret_cast<const float4 *>(weff + (j * K + k) * Cb + tx * 8 + 4);
              acc[k] = fmaf(f[j][0], w0.x, acc[k]); acc[k] = fmaf(f[j][1], w0.y, acc[k]);
              acc[k] = fmaf(f[j][2], w0.z, acc[k]); acc[k] = fmaf(f[j][3], w0.w, acc[k]);
              acc[k] = fmaf(f[j][4], w1.x, acc[k]); acc[k] = fmaf(f[j][5], w1.y, acc[k]);
              acc[k] = fmaf(f[j][6], w1.z, acc[k]); acc[k] = fmaf(f[j][7], w1.w, acc[k]);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) { itv[q] += f[j][q]; hit |= (f[j][q] == mx[j][q]); }
          }
#pragma unroll
        for (int q = 0; q < 8; ++q) hit |= (itv[q] == mx[kMaxJ][q]);
        if (am && hit) {      // rare: some element equals its channel's pooled maximum -> record the FIRST such pixel
#pragma unroll
          for (int j = 0; j < kMaxJ; ++j)
            if (j < J) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                if (f[j][q] == mx[j][q]) atomicMin(am + j * Cb + tx * 8 + q, p);
            }
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (itv[q] == mx[kMaxJ][q]) atomicMin(am + CC + tx * 8 + q, p);
        }
      }
#pragma unroll
      for (int k = 0; k < K; ++k)
        for (int o = 1; o < CVb; o <<= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
      if (ok && tx == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) slog[k * P + px] = acc[k] + cst[k];
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * npx; i += blockDim.x) {
      const int k = i / npx, px = i % npx;
      logits[((size_t)n * K + k) * HW + p0 + px] = slog[k * P + px];
    }
    __syncthreads();
  }
}

// Tensor-core variant of the final pass (bf16 storage, Cb == 32): the CUDA-core kernel above spends ~330 issue slots per
// (pixel, 8-channel) item - 96 FMAs behind 24 LDS.128 of weights, 32 bf16->fp32 conversions, 40 compares - and ran at 1.7 TB/s.
// Here the 1x1 classifier is a warp-level MMA per 16 pixels: the A fragments ARE the raw 16-byte loads (thread (g, t) holds channels
// 8t..8t+7 of pixels g and g+8 of every view; the K index is a free permutation, the B fragments use the same one), the effective
// weights ca*wf are split into bf16 hi + lo parts that sit in neighbouring accumulator columns (n = 2 class + part), so the products
// are exact and thread t < 3 ends up with logit t = c0 + c1 without a shuffle; the arg-max discovery compares packed bf16 pairs.
__device__ __forceinline__ void ecam_mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t ecam_pack2(float lo, float hi) {
  const __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t *>(&h2);
}
__device__ __forceinline__ float2 ecam_unpack2(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u));
}

template <int K>
__global__ void __launch_bounds__(256, 2)
ecam_final_mma_kernel(ViewList xs, int J, int HW, int P, const float *__restrict__ gates, const float *__restrict__ wf,
                      const float *__restrict__ bf, const float *__restrict__ pooled, int *argmax, float *logits) {
  typedef __nv_bfloat16 T;
  constexpr int Cb = 32;
  static_assert(K == 3, "hi/lo columns of 3 classes fill 6 of the 8 accumulator columns");
  extern __shared__ __align__(128) unsigned char smraw[];
  const int n = blockIdx.y, CC = J * Cb, CT = (J + 1) * Cb;
  const uint32_t stage_bytes = (uint32_t)J * P * Cb * sizeof(T);
  unsigned char *stages = smraw;
  unsigned char *barmem = smraw + BULK_STAGES * stage_bytes;
  float *weff = reinterpret_cast<float *>(barmem + 64);   // [J][K][Cb]
  float *cst = weff + K * CC;                              // [K] (+1 pad)
  float *smx = cst + 4;                                    // [CT]
  float *slog = smx + CT;                                  // [K][P] logits of the chunk (coalesced plane stores)
  BulkPipe bp; bp.init(barmem, stages, stage_bytes);
  const float *gt = gates + (size_t)n * CT;
  for (int i = threadIdx.x; i < K * CC; i += blockDim.x) {
    const int j = i / (K * Cb), k = (i / Cb) % K, cb = i % Cb;
    weff[i] = wf[k * CC + j * Cb + cb] * gt[j * Cb + cb];
  }
  for (int i = threadIdx.x; i < CT; i += blockDim.x) smx[i] = pooled ? pooled[((size_t)n * 2 + 1) * CT + i] : 0.f;
  __syncthreads();
  if (threadIdx.x < K) {
    float sacc = bf[threadIdx.x];
    for (int j = 0; j < J; ++j)
      for (int cb = 0; cb < Cb; ++cb) sacc += weff[(j * K + threadIdx.x) * Cb + cb] * gt[CC + cb];
    cst[threadIdx.x] = sacc;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  int *am = argmax ? argmax + (size_t)n * CT : nullptr;
  // B fragments: column n = g -> class g >> 1, part g & 1 (0 = hi, 1 = lo); k-step s of view j covers channels 8t + 4s .. 8t + 4s + 3
  uint32_t bfr[kMaxJ][2][2];
#pragma unroll
  for (int j = 0; j < kMaxJ; ++j)
#pragma unroll
    for (int sst = 0; sst < 2; ++sst)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float v[2] = {0.f, 0.f};
        if (j < J && g < 2 * K) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float wv = weff[(j * K + (g >> 1)) * Cb + 8 * t + 4 * sst + 2 * r + e];
            const float hi = __bfloat162float(__float2bfloat16_rn(wv));
            v[e] = (g & 1) ? (wv - hi) : hi;
          }
        }
        bfr[j][sst][r] = ecam_pack2(v[0], v[1]);
      }
  uint32_t mxp[kMaxJ][4];      // pooled maxima of this thread's channels as packed bf16 pairs (exact: they ARE bf16 values)
  float mxi[8];                // intra maxima (fp32 sums)
#pragma unroll
  for (int j = 0; j < kMaxJ; ++j)
#pragma unroll
    for (int r = 0; r < 4; ++r) mxp[j][r] = (j < J) ? ecam_pack2(smx[j * Cb + t * 8 + 2 * r], smx[j * Cb + t * 8 + 2 * r + 1]) : 0u;
#pragma unroll
  for (int i = 0; i < 8; ++i) mxi[i] = smx[CC + t * 8 + i];
  const float my_cst = (t < K) ? cst[t] : 0.f;
  const int nchunks = (HW + P - 1) / P;
  if (threadIdx.x == 0)
    for (int s = 0; s < BULK_STAGES - 1; ++s) {
      const int c = blockIdx.x + s * gridDim.x;
      if (c < nchunks) bulk_issue<T>(bp, s, xs, J, Cb, n, P, c * P, min(P, HW - c * P), nullptr, 0, HW);
    }
  int it = 0;
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x, ++it) {
    const int s = it % BULK_STAGES;
    if (threadIdx.x == 0) {
      const int cn = c + (BULK_STAGES - 1) * gridDim.x;
      if (cn < nchunks) bulk_issue<T>(bp, (it + BULK_STAGES - 1) % BULK_STAGES, xs, J, Cb, n, P, cn * P, min(P, HW - cn * P), nullptr, 0, HW);
    }
    bk_wait(bp.bars + 8u * s, (uint32_t)((it / BULK_STAGES) & 1));
    const T *st = reinterpret_cast<const T *>(stages + (size_t)s * stage_bytes);
    const int p0 = c * P, npx = min(P, HW - p0);
    for (int base = warp * 16; base < npx; base += 128) {          // warp-uniform
      const int px[2] = {base + g, base + g + 8};
      const bool ok[2] = {px[0] < npx, px[1] < npx};
      uint4 d[2][kMaxJ];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < kMaxJ; ++j)
          d[h][j] = (j < J && ok[h]) ? *reinterpret_cast<const uint4 *>(st + ((size_t)j * P + px[h]) * Cb + t * 8) : make_uint4(0u, 0u, 0u, 0u);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < kMaxJ; ++j)
        if (j < J) {
          ecam_mma_16816(acc, d[0][j].x, d[1][j].x, d[0][j].y, d[1][j].y, bfr[j][0][0], bfr[j][0][1]);
          ecam_mma_16816(acc, d[0][j].z, d[1][j].z, d[0][j].w, d[1][j].w, bfr[j][1][0], bfr[j][1][1]);
        }
      if (t < K) {
        if (ok[0]) slog[t * P + px[0]] = acc[0] + acc[1] + my_cst;
        if (ok[1]) slog[t * P + px[1]] = acc[2] + acc[3] + my_cst;
      }
      if (am) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (!ok[h]) continue;
          float itv[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) itv[q] = 0.f;
          unsigned int hit = 0u;
#pragma unroll
          for (int j = 0; j < kMaxJ; ++j)
            if (j < J) {
              const uint32_t dw[4] = {d[h][j].x, d[h][j].y, d[h][j].z, d[h][j].w};
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                hit |= __heq2_mask(*reinterpret_cast<const __nv_bfloat162 *>(&dw[r]), *reinterpret_cast<const __nv_bfloat162 *>(&mxp[j][r]));
                const float2 f2 = ecam_unpack2(dw[r]);
                itv[2 * r] += f2.x; itv[2 * r + 1] += f2.y;
              }
            }
          bool hit_i = false;
#pragma unroll
          for (int q = 0; q < 8; ++q) hit_i |= (itv[q] == mxi[q]);
          if (hit != 0u || hit_i) {      // rare: some element equals its channel's pooled maximum -> record the FIRST such pixel
            const int p = p0 + px[h];
#pragma unroll
            for (int j = 0; j < kMaxJ; ++j)
              if (j < J) {
                const uint32_t dw[4] = {d[h][j].x, d[h][j].y, d[h][j].z, d[h][j].w};
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                  const float2 f2 = ecam_unpack2(dw[r]), m2 = ecam_unpack2(mxp[j][r]);
                  if (f2.x == m2.x) atomicMin(am + j * Cb + t * 8 + 2 * r, p);
                  if (f2.y == m2.y) atomicMin(am + j * Cb + t * 8 + 2 * r + 1, p);
                }
              }
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (itv[q] == mxi[q]) atomicMin(am + CC + t * 8 + q, p);
          }
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * npx; i += blockDim.x) {
      const int k = i / npx, pxl = i % npx;
      logits[((size_t)n * K + k) * HW + p0 + pxl] = slog[k * P + pxl];
    }
    __syncthreads();
  }
}

template <typename T, int K>
__global__ void __launch_bounds__(256, 2)
ecam_bwd_reduce_bulk_kernel(ViewList xs, int J, int Cb, int HW, int P, const float *__restrict__ dlogits, double *red) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int n = blockIdx.y, CC = J * Cb, RT = K * CC + K;
  const uint32_t stage_bytes = (uint32_t)J * P * Cb * sizeof(T) + (uint32_t)K * P * 4u;
  unsigned char *stages = smraw;
  unsigned char *barmem = smraw + BULK_STAGES * stage_bytes;
  float *sacc = reinterpret_cast<float *>(barmem + 64);
  BulkPipe bp; bp.init(barmem, stages, stage_bytes);
  for (int i = threadIdx.x; i < RT; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  // warp -> view j; lane -> (pixel lane, 8-channel vector)
  const int CVb = Cb / 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwj = (blockDim.x >> 5) / J;                      // warps per view
  const int j = warp % J, tx = lane % CVb;
  const int cb0 = tx * 8, c0 = j * Cb + cb0;
  float B[K][8], D[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    D[k] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) B[k][i] = 0.f;
  }
  const int nchunks = (HW + P - 1) / P;
  if (threadIdx.x == 0)
    for (int s = 0; s < BULK_STAGES - 1; ++s) {
      const int c = blockIdx.x + s * gridDim.x;
      if (c < nchunks) bulk_issue<T>(bp, s, xs, J, Cb, n, P, c * P, min(P, HW - c * P), dlogits, K, HW);
    }
  int it = 0;
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x, ++it) {
    const int s = it % BULK_STAGES;
    if (threadIdx.x == 0) {
      const int cn = c + (BULK_STAGES - 1) * gridDim.x;
      if (cn < nchunks) bulk_issue<T>(bp, (it + BULK_STAGES - 1) % BULK_STAGES, xs, J, Cb, n, P, cn * P, min(P, HW - cn * P), dlogits, K, HW);
    }
    bk_wait(bp.bars + 8u * s, (uint32_t)((it / BULK_STAGES) & 1));
    const T *st = reinterpret_cast<const T *>(stages + (size_t)s * stage_bytes) + (size_t)j * P * Cb;
    const float *sdl = reinterpret_cast<const float *>(stages + (size_t)s * stage_bytes + (size_t)J * P * Cb * sizeof(T));
    const int npx = min(P, HW - c * P);
    for (int i = (warp / J) * 32 + lane; i < npx * CVb; i += nwj * 32) {
      const int px = i / CVb;
      float f[8], dl[K];
      ld8(st + (size_t)px * Cb + cb0, f);
#pragma unroll
      for (int k = 0; k < K; ++k) dl[k] = sdl[k * P + px];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        D[k] += dl[k];
#pragma unroll
        for (int q = 0; q < 8; ++q) B[k][q] = fmaf(dl[k], f[q], B[k][q]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = B[k][i];
      for (int o = CVb; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane < CVb) atomicAdd(&sacc[k * CC + c0 + i], v);
    }
    float d = D[k];
    for (int o = CVb; o < 32; o <<= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0 && j == 0) atomicAdd(&sacc[K * CC + k], d);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < RT; i += blockDim.x) atomicAdd(red + (size_t)n * RT + i, (double)sacc[i]);
}

// host-side: can the bulk path run?  (pixel-dense views, 16-byte multiples everywhere)
static bool bulk_views_ok(const ks_view_t *xs, int J, int Cb, int W, long long HW, int esize) {
  if (g_opt.debug & 64) return false;
  if (HW % 4) return false;
  for (int j = 0; j < J; ++j)
    if (xs[j].sw != Cb || xs[j].sh != (long long)W * Cb || (xs[j].sn * esize) % 16) return false;
  return true;
}
static int bulk_chunk_px(int J, int Cb, int esize) {
  int P = 128;
  while (P > 16 && (size_t)J * P * Cb * esize > 32 * 1024) P >>= 1;
  return P;
}

static int check_views(const ks_view_t *xs, int J, int &Cb, int esize) {
  if (!xs || J < 1 || J > kMaxJ) return KS_EINVAL;
  Cb = xs[0].C;
  if (Cb % 8 != 0 || Cb > 64 || ((Cb / 8) & (Cb / 8 - 1)) != 0) return KS_EUNSUPPORTED;
  for (int j = 0; j < J; ++j) {
    if (xs[j].C != Cb || !xs[j].ptr) return KS_EINVAL;
    if (((uintptr_t)xs[j].ptr % 16) || (xs[j].sn * esize) % 16 || (xs[j].sh * esize) % 16 || (xs[j].sw * esize) % 16) return KS_EUNSUPPORTED;
  }
  return KS_OK;
}

}  // namespace ks

using namespace ks;

extern "C" int ks_ecam_pool(int dtype, int N, int H, int W, const ks_view_t *xs, int J,
                            float *pooled, int *argmax, unsigned long long *scratch, void *stream) {
  KS_CHECK_ARG(pooled && argmax && scratch && N > 0 && H > 0 && W > 0);
  int Cb; int rc = check_views(xs, J, Cb, dtype == KS_F32 ? 4 : 2); if (rc) return rc;
  ViewList vl; rc = make_view_list(xs, J, vl); if (rc) return rc;
  const int CT = (J + 1) * Cb;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int *maxkey = reinterpret_cast<unsigned int *>(scratch);     // first 4 bytes of each 8-byte scratch slot pair
  cudaError_t e = cudaMemsetAsync(pooled, 0, sizeof(float) * (size_t)N * 2 * CT, st); if (e) return (int)e;
  e = cudaMemsetAsync(maxkey, 0, sizeof(unsigned int) * (size_t)N * CT, st); if (e) return (int)e;
  const int es = dtype == KS_F32 ? 4 : 2;
  if (bulk_views_ok(xs, J, Cb, W, (long long)H * W, es) && (dtype == KS_F32 || dtype == KS_BF16)) {
    const int P = bulk_chunk_px(J, Cb, es);
    const int nch = (H * W + P - 1) / P;
    int gx = (kNumSMs * 2) / N; if (gx > nch) gx = nch; if (gx < 1) gx = 1;
    const size_t smem = (size_t)BULK_STAGES * J * P * Cb * es + 64 + (size_t)CT * 8;
    if (dtype == KS_F32) {
      static bool a = false; if (!a) { cudaFuncSetAttribute(ecam_pool_bulk_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024); a = true; }
      ecam_pool_bulk_kernel<float><<<dim3(gx, N), 256, smem, st>>>(vl, J, Cb, H * W, P, pooled, maxkey);
    } else {
      static bool a = false; if (!a) { cudaFuncSetAttribute(ecam_pool_bulk_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024); a = true; }
      ecam_pool_bulk_kernel<__nv_bfloat16><<<dim3(gx, N), 256, smem, st>>>(vl, J, Cb, H * W, P, pooled, maxkey);
    }
    ecam_pool_finalize_kernel<<<(N * CT + 255) / 256, 256, 0, st>>>(N, CT, H * W, pooled, maxkey, argmax);
    KS_LAUNCH_RET();
  }
  const int rows = 256 / (Cb / 8);
  int chunks = (H * W + rows * 8 - 1) / (rows * 8); if (chunks < 1) chunks = 1;
  const int cap = (kNumSMs * 6 + N - 1) / N; if (chunks > cap) chunks = cap;
  const size_t smem = (size_t)CT * (sizeof(unsigned int) + sizeof(float));
  if (dtype == KS_F32) ecam_pool_kernel<float><<<dim3(chunks, N), 256, smem, st>>>(vl, J, Cb, H, W, pooled, maxkey);
  else if (dtype == KS_BF16) ecam_pool_kernel<__nv_bfloat16><<<dim3(chunks, N), 256, smem, st>>>(vl, J, Cb, H, W, pooled, maxkey);
  else return KS_EINVAL;
  ecam_pool_finalize_kernel<<<(N * CT + 255) / 256, 256, 0, st>>>(N, CT, H * W, pooled, maxkey, argmax);
  KS_LAUNCH_RET();
}

extern "C" int ks_ecam_gates(int N, int Cb, int J, int hid, int hid1, const float *pooled,
                             const float *w_fc1, const float *w_fc2, const float *w1_fc1, const float *w1_fc2,
                             float *gates, float *hidden, void *stream) {
  KS_CHECK_ARG(N > 0 && Cb > 0 && J > 0 && hid > 0 && hid1 > 0 && 2 * hid <= 128 && 2 * hid1 <= 128);
  KS_CHECK_ARG(pooled && w_fc1 && w_fc2 && w1_fc1 && w1_fc2 && gates && hidden);
  ecam_gates_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(Cb, J, hid, hid1, pooled, w_fc1, w_fc2, w1_fc1, w1_fc2, gates, hidden);
  KS_LAUNCH_RET();
}

extern "C" int ks_ecam_final(int dtype, int N, int H, int W, const ks_view_t *xs, int J,
                             const float *gates, const float *wf, const float *bf, int K,
                             float *logits, const float *pooled, int *argmax, void *stream) {
  KS_CHECK_ARG(gates && wf && bf && logits && N > 0 && H > 0 && W > 0);
  KS_CHECK_ARG((pooled == nullptr) == (argmax == nullptr));
  if (K != 3) return KS_EUNSUPPORTED;
  int Cb; int rc = check_views(xs, J, Cb, dtype == KS_F32 ? 4 : 2); if (rc) return rc;
  ViewList vl; rc = make_view_list(xs, J, vl); if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int CC = J * Cb, CT = (J + 1) * Cb;
  const int es = dtype == KS_F32 ? 4 : 2;
  if (bulk_views_ok(xs, J, Cb, W, (long long)H * W, es) && (dtype == KS_F32 || dtype == KS_BF16)) {
    const int P = bulk_chunk_px(J, Cb, es);
    const int nch = (H * W + P - 1) / P;
    int gx = (kNumSMs * 2) / N; if (gx > nch) gx = nch; if (gx < 1) gx = 1;
    const size_t smem = (size_t)BULK_STAGES * J * P * Cb * es + 64 + sizeof(float) * (size_t)(K * CC + 4 + CT + K * P);
    if (dtype == KS_BF16 && Cb == 32 && !g_opt.ecam_simt) {
      static bool a = false; if (!a) { cudaFuncSetAttribute(ecam_final_mma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024); a = true; }
      ecam_final_mma_kernel<3><<<dim3(gx, N), 256, smem, st>>>(vl, J, H * W, P, gates, wf, bf, pooled, argmax, logits);
    } else if (dtype == KS_F32) {
      static bool a = false; if (!a) { cudaFuncSetAttribute(ecam_final_bulk_kernel<float, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024); a = true; }
      ecam_final_bulk_kernel<float, 3><<<dim3(gx, N), 256, smem, st>>>(vl, J, Cb, H * W, P, gates, wf, bf, pooled, argmax, logits);
    } else {
      static bool a = false; if (!a) { cudaFuncSetAttribute(ecam_final_bulk_kernel<__nv_bfloat16, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024); a = true; }
      ecam_final_bulk_kernel<__nv_bfloat16, 3><<<dim3(gx, N), 256, smem, st>>>(vl, J, Cb, H * W, P, gates, wf, bf, pooled, argmax, logits);
    }
    KS_LAUNCH_RET();
  }
  const int rows_f = 256 / (Cb / 8);
  int chunks = (H * W + rows_f * 8 - 1) / (rows_f * 8); if (chunks < 1) chunks = 1;
  const int cap = (kNumSMs * 8 + N - 1) / N; if (chunks > cap) chunks = cap;
  const size_t smem = sizeof(float) * (size_t)(K * CC + 4 + CT);
  if (dtype == KS_F32) ecam_final_kernel<float, 3><<<dim3(chunks, N), 256, smem, st>>>(vl, J, Cb, H, W, gates, wf, bf, pooled, argmax, logits);
  else if (dtype == KS_BF16) ecam_final_kernel<__nv_bfloat16, 3><<<dim3(chunks, N), 256, smem, st>>>(vl, J, Cb, H, W, gates, wf, bf, pooled, argmax, logits);
  else return KS_EINVAL;
  KS_LAUNCH_RET();
}

extern "C" int ks_ecam_bwd_reduce(int dtype, int N, int H, int W, const ks_view_t *xs, int J, int K,
                                  const float *dlogits, double *red, void *stream) {
  KS_CHECK_ARG(dlogits && red && N > 0 && H > 0 && W > 0);
  if (K != 3) return KS_EUNSUPPORTED;
  int Cb; int rc = check_views(xs, J, Cb, dtype == KS_F32 ? 4 : 2); if (rc) return rc;
  ViewList vl; rc = make_view_list(xs, J, vl); if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int RT = K * J * Cb + K;
  cudaError_t e = cudaMemsetAsync(red, 0, sizeof(double) * (size_t)N * RT, st); if (e) return (int)e;
  const int es = dtype == KS_F32 ? 4 : 2;
  if (bulk_views_ok(xs, J, Cb, W, (long long)H * W, es) && (dtype == KS_F32 || dtype == KS_BF16) && 8 % J == 0 && (((uintptr_t)dlogits) % 16) == 0) {
    const int P = bulk_chunk_px(J, Cb, es);
    const int nch = (H * W + P - 1) / P;
    int gx = (kNumSMs * 2) / N; if (gx > nch) gx = nch; if (gx < 1) gx = 1;
    const size_t smem = (size_t)BULK_STAGES * ((size_t)J * P * Cb * es + (size_t)K * P * 4) + 64 + sizeof(float) * (size_t)RT;
    if (dtype == KS_F32) {
      static bool a = false; if (!a) { cudaFuncSetAttribute(ecam_bwd_reduce_bulk_kernel<float, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024); a = true; }
      ecam_bwd_reduce_bulk_kernel<float, 3><<<dim3(gx, N), 256, smem, st>>>(vl, J, Cb, H * W, P, dlogits, red);
    } else {
      static bool a = false; if (!a) { cudaFuncSetAttribute(ecam_bwd_reduce_bulk_kernel<__nv_bfloat16, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024); a = true; }
      ecam_bwd_reduce_bulk_kernel<__nv_bfloat16, 3><<<dim3(gx, N), 256, smem, st>>>(vl, J, Cb, H * W, P, dlogits, red);
    }
    KS_LAUNCH_RET();
  }
  const int nthr = 32 * J * (8 / J), rows = (8 / J) * (32 / (Cb / 8));
  int chunks = (H * W + rows * 16 - 1) / (rows * 16); if (chunks < 1) chunks = 1;
  const int cap = (kNumSMs * 8 + N - 1) / N; if (chunks > cap) chunks = cap;
  const size_t smem = sizeof(float) * (size_t)RT;
  if (dtype == KS_F32) ecam_bwd_reduce_kernel<float, 3><<<dim3(chunks, N), nthr, smem, st>>>(vl, J, Cb, H, W, dlogits, red);
  else if (dtype == KS_BF16) ecam_bwd_reduce_kernel<__nv_bfloat16, 3><<<dim3(chunks, N), nthr, smem, st>>>(vl, J, Cb, H, W, dlogits, red);
  else return KS_EINVAL;
  KS_LAUNCH_RET();
}

extern "C" int ks_ecam_gates_bwd(int N, int Cb, int J, int hid, int hid1, int K, const float *pooled,
                                 const float *hidden, const float *gates, const double *red, const float *wf,
                                 const float *w_fc1, const float *w_fc2, const float *w1_fc1, const float *w1_fc2,
                                 float *dpooled, float *dwf, float *dbf, float *dw_fc1, float *dw_fc2,
                                 float *dw1_fc1, float *dw1_fc2, int accumulate, void *stream) {
  KS_CHECK_ARG(N > 0 && Cb > 0 && J > 0 && hid > 0 && hid1 > 0 && K > 0);
  KS_CHECK_ARG(pooled && hidden && gates && red && wf && w_fc1 && w_fc2 && w1_fc1 && w1_fc2 && dpooled);
  const int CC = J * Cb, CT = (J + 1) * Cb, HT = hid + hid1;
  const size_t smem = sizeof(float) * (size_t)(CT + 2 * HT);
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) {
    struct { float *p; size_t n; } z[6] = {{dwf, (size_t)K * CC}, {dbf, (size_t)K}, {dw_fc1, (size_t)hid * CC}, {dw_fc2, (size_t)CC * hid},
                                           {dw1_fc1, (size_t)hid1 * Cb}, {dw1_fc2, (size_t)Cb * hid1}};
    for (auto &e : z) if (e.p) { cudaError_t er = cudaMemsetAsync(e.p, 0, sizeof(float) * e.n, st); if (er != cudaSuccess) return (int)er; }
  }
  ecam_gates_bwd_kernel<<<N, 256, smem, st>>>(N, Cb, J, hid, hid1, K, pooled, hidden, gates, red, wf,
      w_fc1, w_fc2, w1_fc1, w1_fc2, dpooled, dwf, dbf, dw_fc1, dw_fc2, dw1_fc1, dw1_fc2);
  KS_LAUNCH_RET();
}

extern "C" int ks_ecam_bwd_apply(int dtype, int N, int H, int W, int J, int Cb, const float *gates, const float *wf, int K,
                                 const float *dlogits, const float *dpooled, const int *argmax,
                                 const ks_view_t *dxs, void *stream) {
  KS_CHECK_ARG(gates && wf && dlogits && dpooled && argmax && N > 0 && H > 0 && W > 0);
  if (K != 3) return KS_EUNSUPPORTED;
  int Cb2; int rc = check_views(dxs, J, Cb2, dtype == KS_F32 ? 4 : 2); if (rc) return rc;
  if (Cb2 != Cb) return KS_EINVAL;
  ViewList vl; rc = make_view_list(dxs, J, vl); if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int nthr = 32 * J * (8 / J), rows_a = (8 / J) * (32 / (Cb / 8));
  int chunks = (H * W + rows_a * 16 - 1) / (rows_a * 16); if (chunks < 1) chunks = 1;
  const int cap = (kNumSMs * 16 + N - 1) / N; if (chunks > cap) chunks = cap;
  const size_t smem = 0;
  if (dtype == KS_F32) ecam_bwd_apply_kernel<float, 3><<<dim3(chunks, N), nthr, smem, st>>>(vl, J, Cb, H, W, gates, wf, dlogits, dpooled, argmax);
  else if (dtype == KS_BF16) ecam_bwd_apply_kernel<__nv_bfloat16, 3><<<dim3(chunks, N), nthr, smem, st>>>(vl, J, Cb, H, W, gates, wf, dlogits, dpooled, argmax);
  else return KS_EINVAL;
  KS_LAUNCH_RET();
}
