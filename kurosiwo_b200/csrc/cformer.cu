// Passes of ChangeFormerV6 (models/changeformer.py) that are not a stride-1 3x3 / 1x1 convolution, LayerNorm, GELU or BatchNorm
// (those run on the engines in conv_*.cu, vit.cu, elementwise.cu):
//   ks_conv2d_strided / _dgrad / _wgrad   OverlapPatchEmbed.proj 7x7 s4|s2 p3 (:281,288) and Attention.sr k = s = sr (:166,193)
//   ks_xattention_fwd / _bwd              spatial-reduction attention: N queries x (<= 64) reduced keys (:186-208)
//   ks_dwconv3x3_fwd / _bwd               Mix-FFN depth-wise 3x3 conv (:84-96)
//   ks_bilinear_nhwc_fwd / _bwd           F.interpolate(..., mode='bilinear', align_corners=False) of NHWC maps (:585-608)
//   ks_relu_fwd / _bwd, ks_sigmoid_head_fwd / _bwd   ReLU outside a BatchNorm pass (:31-38,471-483), final Sigmoid (:635-639)
//   ks_dropout_apply, ks_branch_add / _scale         Dropout / DropPath with a stateless RNG (:129-132,206,246-247,652-654)
// Token matrices [B*N, C] ARE NHWC images [B, H, W, C] (N = H*W row-major), so no transposes exist on this path.
// First correct path: the strided convolutions and the attention are exact-fp32 CUDA-core kernels for both storage dtypes
// (6 % of the model's FLOPs, models/changeformer.py encoder); the 94 % in the decoder's 256-channel 3x3 convs run on tcgen05.
#include "common.cuh"

namespace ks {

// stateless RNG of the stochastic regularisers (see the dropout section below)
__device__ __forceinline__ unsigned long long cf_mix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ unsigned long long cf_key(unsigned long long seed, const int *step_ptr, int site) {
  const unsigned long long step = step_ptr ? (unsigned long long)(unsigned int)*step_ptr : 0ull;
  return cf_mix64(seed ^ (step << 32) ^ ((unsigned long long)(unsigned int)site * 0x632BE59BD9B4E019ull));
}
// keep/(1-p) factor of element idx: 0 with probability p, 1/(1-p) otherwise
__device__ __forceinline__ float cf_keep(unsigned long long key, unsigned long long idx, float p, float inv_keep) {
  const unsigned long long r = cf_mix64(key + idx);
  return ((float)(r >> 40) * (1.0f / 16777216.0f) >= p) ? inv_keep : 0.f;
}

// ---------------------------------------------------------------------------------------------------------
// Generic strided convolution, implicit GEMM 64x64x16 on CUDA cores.
//   MODE 0 (forward):  out[n,ho,wo,co] = bias[co] + sum_{ky,kx,ci} x[n, ho*s-p+ky, wo*s-p+kx, ci] * w[ky*k+kx][co][ci]
//   MODE 1 (data grad): dx[n,h,w,ci] (+)= sum_{ky,kx,co} dy[n, (h+p-ky)/s, (w+p-kx)/s, co] * w[ky*k+kx][co][ci]   (divisible only)
// ---------------------------------------------------------------------------------------------------------
constexpr int GBM = 64, GBN = 64, GBK = 16;

template <typename T, int MODE>
__global__ void __launch_bounds__(256)
conv_gen_kernel(View src, View dst, int N, int Hs, int Ws, int Hd, int Wd, int ksize, int stride, int pad, int Cin, int Cout,
                const T *__restrict__ weight, const float *__restrict__ bias, int accumulate) {
  // src/dst: MODE 0 src = x [N,Hs,Ws,Cin], dst = out [N,Hd,Wd,Cout];  MODE 1 src = dy [N,Hs,Ws,Cout], dst = dx [N,Hd,Wd,Cin]
  // MODE 1 runs one grid.z slice per input-pixel PHASE (h % stride, w % stride): all pixels of a slice share the set of taps that
  // divide evenly, (ky, kx) = (ky0 + i*stride, kx0 + j*stride), so only those k^2/stride^2 taps are visited (1 tap when k == stride).
  __shared__ float As[GBK][GBM + 4];
  __shared__ float Bs[GBK][GBN + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int ph = (MODE == 1) ? (int)blockIdx.z / stride : 0, pw = (MODE == 1) ? (int)blockIdx.z % stride : 0;
  const int Hp = (MODE == 1) ? (Hd - ph + stride - 1) / stride : Hd, Wp = (MODE == 1) ? (Wd - pw + stride - 1) / stride : Wd;
  const long long M = (long long)N * Hp * Wp;
  const long long m0 = (long long)blockIdx.x * GBM;
  if (m0 >= M) return;
  const int n0 = blockIdx.y * GBN;
  const int CK = (MODE == 0) ? Cin : Cout;      // reduction channels
  const int CN = (MODE == 0) ? Cout : Cin;      // output channels
  const int lrow = tid / 4, lk = (tid % 4) * 4;
  const long long lm = m0 + lrow;
  const bool lm_ok = lm < M;
  int ln = 0, lh = 0, lw = 0;
  if (lm_ok) { lw = (int)(lm % Wp); long long r = lm / Wp; lh = (int)(r % Hp); ln = (int)(r / Hp); }
  if (MODE == 1) { lh = lh * stride + ph; lw = lw * stride + pw; }
  const int lcn = n0 + lrow;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int ky0 = (MODE == 1) ? (ph + pad) % stride : 0, kx0 = (MODE == 1) ? (pw + pad) % stride : 0;
  const int kstep = (MODE == 1) ? stride : 1;
  for (int ky = ky0; ky < ksize; ky += kstep) {
    for (int kx = kx0; kx < ksize; kx += kstep) {
      const int tap = ky * ksize + kx;
      int hh, ww;
      if (MODE == 0) { hh = lh * stride - pad + ky; ww = lw * stride - pad + kx; }
      else { hh = (lh + pad - ky) / stride; ww = (lw + pad - kx) / stride; }       // exact by construction; may be negative -> masked
      const bool pix_ok = lm_ok && (MODE == 0 || (lh + pad - ky >= 0 && lw + pad - kx >= 0)) && hh >= 0 && hh < Hs && ww >= 0 && ww < Ws;
      const T *xp = reinterpret_cast<const T *>(src.ptr) + ((long long)ln * src.sn + (long long)hh * src.sh + (long long)ww * src.sw);
      const T *wt = weight + (long long)tap * Cout * Cin;
      for (int c0 = 0; c0 < CK; c0 += GBK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = c0 + lk + i;
          float a = 0.f, b = 0.f;
          if (c < CK) {
            if (pix_ok) a = Cvt<T>::ld(xp + c);
            if (lcn < CN) b = Cvt<T>::ld(MODE == 0 ? (wt + (long long)lcn * Cin + c) : (wt + (long long)c * Cin + lcn));
          }
          As[lk + i][lrow] = a;
          Bs[lk + i][lrow] = b;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GBK; ++k) {
          const float4 a4 = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
          const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
          const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    int w = (int)(m % Wp); long long r = m / Wp; int h = (int)(r % Hp); const int n = (int)(r / Hp);
    if (MODE == 1) { h = h * stride + ph; w = w * stride + pw; }
    T *op = reinterpret_cast<T *>(dst.ptr) + ((long long)n * dst.sn + (long long)h * dst.sh + (long long)w * dst.sw);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cn = n0 + tx * 4 + j;
      if (cn >= CN) continue;
      float v = acc[i][j] + ((MODE == 0 && bias) ? __ldg(bias + cn) : 0.f);
      if (accumulate) v += Cvt<T>::ld(op + cn);
      Cvt<T>::st(op + cn, v);
    }
  }
}

// dw[tap][co][ci] (+)= sum_{n,ho,wo} dy[n,ho,wo,co] * x[n, ho*s-p+ky, wo*s-p+kx, ci];  grid (ci blocks, co blocks, taps*splits)
template <typename T>
__global__ void __launch_bounds__(256)
conv_gen_wgrad_kernel(View x, View dy, int N, int Hi, int Wi, int Ho, int Wo, int ksize, int stride, int pad, int Cin, int Cout,
                      int splits, float *dw) {
  __shared__ float As[GBK][GBM + 4];  // dY tile [px][co]
  __shared__ float Bs[GBK][GBN + 4];  // X tile  [px][ci]
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int tap = blockIdx.z / splits, split = blockIdx.z % splits;
  const int ky = tap / ksize, kx = tap % ksize;
  const int ci0 = blockIdx.x * 64, co0 = blockIdx.y * 64;
  const long long M = (long long)N * Ho * Wo;
  const long long per = ((M + splits - 1) / splits + GBK - 1) / GBK * GBK;
  const long long p0 = (long long)split * per, p1 = min(M, p0 + per);
  const int lp = tid / 16, lc = (tid % 16) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long pb = p0; pb < p1; pb += GBK) {
    const long long p = pb + lp;
    float a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
    if (p < p1) {
      const int wo = (int)(p % Wo); long long r = p / Wo; const int ho = (int)(r % Ho); const int n = (int)(r / Ho);
      const T *yp = reinterpret_cast<const T *>(dy.ptr) + ((long long)n * dy.sn + (long long)ho * dy.sh + (long long)wo * dy.sw + co0 + lc);
#pragma unroll
      for (int i = 0; i < 4; ++i) if (co0 + lc + i < Cout) a[i] = Cvt<T>::ld(yp + i);
      const int hh = ho * stride - pad + ky, ww = wo * stride - pad + kx;
      if (hh >= 0 && hh < Hi && ww >= 0 && ww < Wi) {
        const T *xp = reinterpret_cast<const T *>(x.ptr) + ((long long)n * x.sn + (long long)hh * x.sh + (long long)ww * x.sw + ci0 + lc);
#pragma unroll
        for (int i = 0; i < 4; ++i) if (ci0 + lc + i < Cin) b[i] = Cvt<T>::ld(xp + i);
      }
    }
    *reinterpret_cast<float4 *>(&As[lp][lc]) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4 *>(&Bs[lp][lc]) = make_float4(b[0], b[1], b[2], b[3]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
      const float aa[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci >= Cin) continue;
      atomicAdd(dw + ((long long)tap * Cout + co) * Cin + ci, acc[i][j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Spatial-reduction attention (exact fp32 arithmetic).  q: [B*Nq, ldq] (head h at columns h*dh..), kv: [B*Nk, ldkv] with
// k at columns h*dh.. and v at heads*dh + h*dh.. ('reshape(B,-1,2,heads,d)', :195-198); Nk <= 64, dh <= 96.
// One CTA per (row chunk, head, image); K/V of the head in shared memory (pitch dh+1); one warp per query row.
// ---------------------------------------------------------------------------------------------------------
constexpr int XA_WARPS = 8, XA_DMAX = 96, XA_KMAX = 64;

template <typename T>
__device__ __forceinline__ void xa_load(float *dst, const T *src, long long ld, int rows, int dh) {
  for (int i = threadIdx.x; i < rows * dh; i += blockDim.x) {
    const int r = i / dh, c = i % dh;
    dst[r * (dh + 1) + c] = Cvt<T>::ld(src + (long long)r * ld + c);
  }
}

template <typename T>
__global__ void __launch_bounds__(XA_WARPS * 32)
xattention_fwd_kernel(int Nq, int Nk, int heads, int dh, const T *__restrict__ q, long long ldq, const T *__restrict__ kv, long long ldkv,
                      float scale, T *__restrict__ out, long long ldo, T *__restrict__ probs, int rows_per_cta,
                      float pdrop, unsigned long long seed, const int *step_ptr, int site) {
  extern __shared__ float sm[];
  const int P = dh + 1;
  const unsigned long long dkey = cf_key(seed, step_ptr, site);
  const float ikeep = 1.f / (1.f - pdrop);
  float *Ks = sm, *Vs = Ks + Nk * P, *Qw = Vs + Nk * P, *Pw = Qw + XA_WARPS * XA_DMAX;
  const int b = blockIdx.z, h = blockIdx.y, inner = heads * dh;
  const T *kvb = kv + (long long)b * Nk * ldkv + h * dh;
  xa_load(Ks, kvb, ldkv, Nk, dh);
  xa_load(Vs, kvb + inner, ldkv, Nk, dh);
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float *qw = Qw + w * XA_DMAX, *pw = Pw + w * XA_KMAX;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(Nq, r0 + rows_per_cta);
  for (int i = r0 + w; i < r1; i += XA_WARPS) {
    const T *qrow = q + ((long long)b * Nq + i) * ldq + h * dh;
    for (int d = lane; d < dh; d += 32) qw[d] = Cvt<T>::ld(qrow + d);
    __syncwarp();
    float s[2], mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = jj * 32 + lane;
      s[jj] = -INFINITY;
      if (j < Nk) {
        float a = 0.f;
        const float *kp = Ks + j * P;
        for (int d = 0; d < dh; ++d) a = fmaf(qw[d], kp[d], a);
        s[jj] = a * scale; mx = fmaxf(mx, s[jj]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) { s[jj] = (jj * 32 + lane < Nk) ? expf(s[jj] - mx) : 0.f; sum += s[jj]; }
    const float inv = 1.f / warp_sum(sum);
    T *prow = probs + (((long long)b * heads + h) * Nq + i) * Nk;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = jj * 32 + lane;
      if (j < Nk) {
        const float pq = round_as<T>(s[jj] * inv);
        Cvt<T>::st(prow + j, pq);                  // the softmax output is kept; attn_drop (:203) acts on the copy used for P.V
        pw[j] = (pdrop > 0.f) ? pq * cf_keep(dkey, (unsigned long long)((((long long)b * heads + h) * Nq + i) * Nk + j), pdrop, ikeep) : pq;
      }
    }
    __syncwarp();
    T *orow = out + ((long long)b * Nq + i) * ldo + h * dh;
    for (int d = lane; d < dh; d += 32) {
      float o = 0.f;
      for (int j = 0; j < Nk; ++j) o = fmaf(pw[j], Vs[j * P + d], o);
      Cvt<T>::st(orow + d, o);
    }
    __syncwarp();
  }
}

template <typename T>
__global__ void __launch_bounds__(XA_WARPS * 32)
xattention_bwd_kernel(int Nq, int Nk, int heads, int dh, const T *__restrict__ q, long long ldq, const T *__restrict__ kv, long long ldkv,
                      const T *__restrict__ probs, const T *__restrict__ dout, long long ldo, float scale, T *__restrict__ dq, long long lddq,
                      float *__restrict__ dkv, int rows_per_cta, float pdrop, unsigned long long seed, const int *step_ptr, int site) {
  extern __shared__ float sm[];
  const int P = dh + 1;
  const unsigned long long dkey = cf_key(seed, step_ptr, site);
  const float ikeep = 1.f / (1.f - pdrop);
  float *Ks = sm, *Vs = Ks + Nk * P, *dKs = Vs + Nk * P, *dVs = dKs + Nk * P, *Qw = dVs + Nk * P, *Ow = Qw + XA_WARPS * XA_DMAX,
        *Pw = Ow + XA_WARPS * XA_DMAX, *Sw = Pw + XA_WARPS * XA_KMAX;
  const int b = blockIdx.z, h = blockIdx.y, inner = heads * dh;
  const T *kvb = kv + (long long)b * Nk * ldkv + h * dh;
  xa_load(Ks, kvb, ldkv, Nk, dh);
  xa_load(Vs, kvb + inner, ldkv, Nk, dh);
  for (int i = threadIdx.x; i < 2 * Nk * P; i += blockDim.x) dKs[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float *qw = Qw + w * XA_DMAX, *ow = Ow + w * XA_DMAX, *pw = Pw + w * XA_KMAX, *sw = Sw + w * XA_KMAX;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(Nq, r0 + rows_per_cta);
  for (int i = r0 + w; i < r1; i += XA_WARPS) {
    const T *qrow = q + ((long long)b * Nq + i) * ldq + h * dh;
    const T *dorow = dout + ((long long)b * Nq + i) * ldo + h * dh;
    const T *prow = probs + (((long long)b * heads + h) * Nq + i) * Nk;
    for (int d = lane; d < dh; d += 32) { qw[d] = Cvt<T>::ld(qrow + d); ow[d] = Cvt<T>::ld(dorow + d); }
    __syncwarp();
    float dp[2], pv[2], kf[2], dot = 0.f;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = jj * 32 + lane;
      dp[jj] = 0.f; pv[jj] = 0.f; kf[jj] = 1.f;
      if (j < Nk) {
        float a = 0.f;
        const float *vp = Vs + j * P;
        for (int d = 0; d < dh; ++d) a = fmaf(ow[d], vp[d], a);
        if (pdrop > 0.f) kf[jj] = cf_keep(dkey, (unsigned long long)((((long long)b * heads + h) * Nq + i) * Nk + j), pdrop, ikeep);
        dp[jj] = a * kf[jj]; pv[jj] = Cvt<T>::ld(prow + j);         // d(softmax output) = d(dropped copy) * keep/(1-p)
        dot += dp[jj] * pv[jj];
      }
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = jj * 32 + lane;
      if (j < Nk) { pw[j] = pv[jj] * kf[jj]; sw[j] = pv[jj] * (dp[jj] - dot) * scale; }     // pw: the dropped probabilities (dV = Pdrop^T dO)
    }
    __syncwarp();
    T *dqrow = dq + ((long long)b * Nq + i) * lddq + h * dh;
    for (int d = lane; d < dh; d += 32) {
      float o = 0.f;
      const float qd = qw[d], od = ow[d];
      for (int j = 0; j < Nk; ++j) {
        o = fmaf(sw[j], Ks[j * P + d], o);
        atomicAdd(&dKs[j * P + d], sw[j] * qd);
        atomicAdd(&dVs[j * P + d], pw[j] * od);
      }
      Cvt<T>::st(dqrow + d, o);
    }
    __syncwarp();
  }
  __syncthreads();
  float *dkvb = dkv + (long long)b * Nk * (2 * inner) + h * dh;
  for (int i = threadIdx.x; i < Nk * dh; i += blockDim.x) {
    const int j = i / dh, d = i % dh;
    atomicAdd(dkvb + (long long)j * (2 * inner) + d, dKs[j * P + d]);
    atomicAdd(dkvb + (long long)j * (2 * inner) + inner + d, dVs[j * P + d]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// bf16, dh == 64, Nk <= 64: the forward on mma.sync.m16n8k16 (same scheme as the ViT kernels in vit.cu): K and V of the (batch, head)
// in shared memory as bf16 [64][72] (rows >= Nk zero), one warp per 16-query strip, the 16 x 64 score block in accumulator fragments,
// softmax in registers (MUFU exp), P (bf16, pre-dropout) stored for the backward, dropped P re-used in registers as the A operand of
// P.V.  The CUDA-core kernel above stays the fp32 / generic path.
// ---------------------------------------------------------------------------------------------------------
constexpr int XT_KP = 72;
__device__ __forceinline__ void xt_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void xt_ldsm(uint32_t (&r)[4], const void *p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void xt_ldsm_t(uint32_t (&r)[4], const void *p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ uint32_t xt_pack(float lo, float hi) {
  const __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t *>(&h2);
}
__device__ __forceinline__ void xt_load_kv(__nv_bfloat16 *dst, const __nv_bfloat16 *src, long long ld, int rows) {
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < rows) v = *reinterpret_cast<const uint4 *>(src + (long long)r * ld + c);
    *reinterpret_cast<uint4 *>(dst + r * XT_KP + c) = v;
  }
}

__global__ void __launch_bounds__(XA_WARPS * 32)
xattention_fwd_tc_kernel(int Nq, int Nk, int heads, const __nv_bfloat16 *__restrict__ q, long long ldq, const __nv_bfloat16 *__restrict__ kv,
                         long long ldkv, float scale, __nv_bfloat16 *__restrict__ out, long long ldo, __nv_bfloat16 *__restrict__ probs,
                         int rows_per_cta, float pdrop, unsigned long long seed, const int *step_ptr, int site) {
  __shared__ __align__(16) __nv_bfloat16 Ks[64 * XT_KP], Vs[64 * XT_KP];
  const unsigned long long dkey = cf_key(seed, step_ptr, site);
  const float ikeep = 1.f / (1.f - pdrop);
  const int b = blockIdx.z, h = blockIdx.y, inner = heads * 64;
  const __nv_bfloat16 *kvb = kv + (long long)b * Nk * ldkv + h * 64;
  xt_load_kv(Ks, kvb, ldkv, Nk);
  xt_load_kv(Vs, kvb + inner, ldkv, Nk);
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(Nq, r0 + rows_per_cta);
  const float sl2 = scale * 1.4426950408889634f;
  for (int i0 = r0 + w * 16; i0 < r1; i0 += XA_WARPS * 16) {
    const int ia = i0 + g, ib = i0 + g + 8;
    const bool ok0 = ia < r1, ok1 = ib < r1;
    // A fragments of the 16 query rows (rows past the range are clamped for the load and never stored)
    const __nv_bfloat16 *qa0 = q + ((long long)b * Nq + min(ia, Nq - 1)) * ldq + h * 64 + 2 * t;
    const __nv_bfloat16 *qa1 = q + ((long long)b * Nq + min(ib, Nq - 1)) * ldq + h * 64 + 2 * t;
    uint32_t qa[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      qa[kk][0] = *reinterpret_cast<const uint32_t *>(qa0 + kk * 16);
      qa[kk][1] = *reinterpret_cast<const uint32_t *>(qa1 + kk * 16);
      qa[kk][2] = *reinterpret_cast<const uint32_t *>(qa0 + kk * 16 + 8);
      qa[kk][3] = *reinterpret_cast<const uint32_t *>(qa1 + kk * 16 + 8);
    }
    float s[8][4];
#pragma unroll
    for (int jt = 0; jt < 8; ++jt) s[jt][0] = s[jt][1] = s[jt][2] = s[jt][3] = 0.f;
    const __nv_bfloat16 *krow = Ks + (lane & 7) * XT_KP + (lane >> 3) * 8;
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
#pragma unroll
      for (int jt = 0; jt < 8; jt += 2) {
        uint32_t ra[4], rb[4];
        xt_ldsm(ra, krow + jt * 8 * XT_KP + h2 * 32);
        xt_ldsm(rb, krow + (jt + 1) * 8 * XT_KP + h2 * 32);
        xt_mma(s[jt], qa[2 * h2], ra[0], ra[1]);
        xt_mma(s[jt + 1], qa[2 * h2], rb[0], rb[1]);
        xt_mma(s[jt], qa[2 * h2 + 1], ra[2], ra[3]);
        xt_mma(s[jt + 1], qa[2 * h2 + 1], rb[2], rb[3]);
      }
    }
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int jt = 0; jt < 8; ++jt) {
      const int c = jt * 8 + 2 * t;
      if (c < Nk) { m0 = fmaxf(m0, s[jt][0]); m1 = fmaxf(m1, s[jt][2]); }
      if (c + 1 < Nk) { m0 = fmaxf(m0, s[jt][1]); m1 = fmaxf(m1, s[jt][3]); }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    const float m0s = m0 * sl2, m1s = m1 * sl2;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int jt = 0; jt < 8; ++jt) {
      const int c = jt * 8 + 2 * t;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool in = (c + (e & 1)) < Nk;
        float y;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(fmaf(s[jt][e], sl2, (e < 2) ? -m0s : -m1s)));
        s[jt][e] = in ? y : 0.f;
      }
      l0 += s[jt][0] + s[jt][1]; l1 += s[jt][2] + s[jt][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    uint32_t pk[8][2];
    const long long pbase0 = (((long long)b * heads + h) * Nq + ia) * Nk, pbase1 = (((long long)b * heads + h) * Nq + ib) * Nk;
#pragma unroll
    for (int jt = 0; jt < 8; ++jt) {
      const int c = jt * 8 + 2 * t;
      float pq[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) pq[e] = round_as<__nv_bfloat16>(s[jt][e] * ((e < 2) ? inv0 : inv1));
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int cc = c + (e & 1);
        const bool rowok = (e < 2) ? ok0 : ok1;
        if (rowok && cc < Nk) {
          const long long idx = ((e < 2) ? pbase0 : pbase1) + cc;
          probs[idx] = __float2bfloat16_rn(pq[e]);                 // Nk is odd in general: 2-byte stores
          if (pdrop > 0.f) pq[e] *= cf_keep(dkey, (unsigned long long)idx, pdrop, ikeep);
        } else pq[e] = 0.f;
      }
      pk[jt][0] = xt_pack(pq[0], pq[1]);
      pk[jt][1] = xt_pack(pq[2], pq[3]);
    }
    float o[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t a[4] = {pk[2 * kk][0], pk[2 * kk][1], pk[2 * kk + 1][0], pk[2 * kk + 1][1]};
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t r[4];
        xt_ldsm_t(r, Vs + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * XT_KP + (dp * 2 + (lane >> 4)) * 8);
        xt_mma(o[2 * dp], a, r[0], r[1]);
        xt_mma(o[2 * dp + 1], a, r[2], r[3]);
      }
    }
    __nv_bfloat16 *o0 = out + ((long long)b * Nq + ia) * ldo + h * 64 + 2 * t, *o1 = out + ((long long)b * Nq + ib) * ldo + h * 64 + 2 * t;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      if (ok0) *reinterpret_cast<uint32_t *>(o0 + dt * 8) = xt_pack(o[dt][0], o[dt][1]);
      if (ok1) *reinterpret_cast<uint32_t *>(o1 + dt * 8) = xt_pack(o[dt][2], o[dt][3]);
    }
  }
}

// Backward of the same (bf16, dh == 64, Nk <= 64), all four products on mma.sync.  Per batch of 128 query rows of the CTA's range:
//   stage  Q and dO rows in shared memory ([128][72] bf16);
//   pass A (a warp per 16 rows): dP = dO V^T, dS = P*(dP*keep - rowsum)*scale and Pdrop = P*keep in registers, dQ = dS K -> global,
//          dS and Pdrop -> shared memory ([128][72] bf16);
//   pass B (warp = 16 keys x 32 d): dK += dS^T Q, dV += Pdrop^T dO with ldmatrix.trans operands, accumulators live in registers over
//          all batches and are added to the fp32 dkv buffer once at the end.
constexpr int XT_ROWS = 128;

__global__ void __launch_bounds__(XA_WARPS * 32, 2)
xattention_bwd_tc_kernel(int Nq, int Nk, int heads, const __nv_bfloat16 *__restrict__ q, long long ldq, const __nv_bfloat16 *__restrict__ kv,
                         long long ldkv, const __nv_bfloat16 *__restrict__ probs, const __nv_bfloat16 *__restrict__ dout, long long ldo,
                         float scale, __nv_bfloat16 *__restrict__ dq, long long lddq, float *__restrict__ dkv, int rows_per_cta, float pdrop,
                         unsigned long long seed, const int *step_ptr, int site) {
  extern __shared__ __align__(16) unsigned char xsm[];
  __nv_bfloat16 *Ks = reinterpret_cast<__nv_bfloat16 *>(xsm), *Vs = Ks + 64 * XT_KP, *Qb = Vs + 64 * XT_KP, *Ob = Qb + XT_ROWS * XT_KP,
                *Sb = Ob + XT_ROWS * XT_KP, *Pb = Sb + XT_ROWS * XT_KP;
  const unsigned long long dkey = cf_key(seed, step_ptr, site);
  const float ikeep = 1.f / (1.f - pdrop);
  const int b = blockIdx.z, h = blockIdx.y, inner = heads * 64;
  const __nv_bfloat16 *kvb = kv + (long long)b * Nk * ldkv + h * 64;
  xt_load_kv(Ks, kvb, ldkv, Nk);
  xt_load_kv(Vs, kvb + inner, ldkv, Nk);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int mt = w & 3, dhalf = w >> 2;                         // pass B ownership
  float aK[4][4], aV[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { aK[i][0] = aK[i][1] = aK[i][2] = aK[i][3] = 0.f; aV[i][0] = aV[i][1] = aV[i][2] = aV[i][3] = 0.f; }
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(Nq, r0 + rows_per_cta);
  for (int rb = r0; rb < r1; rb += XT_ROWS) {
    __syncthreads();                                            // K/V staged (first batch); pass B of the previous batch done
    for (int i = threadIdx.x; i < XT_ROWS * 8; i += blockDim.x) {
      const int r = i >> 3, c = (i & 7) * 8, row = rb + r;
      uint4 vq = make_uint4(0u, 0u, 0u, 0u), vo = vq;
      if (row < r1) {
        vq = *reinterpret_cast<const uint4 *>(q + ((long long)b * Nq + row) * ldq + h * 64 + c);
        vo = *reinterpret_cast<const uint4 *>(dout + ((long long)b * Nq + row) * ldo + h * 64 + c);
      }
      *reinterpret_cast<uint4 *>(Qb + r * XT_KP + c) = vq;
      *reinterpret_cast<uint4 *>(Ob + r * XT_KP + c) = vo;
    }
    __syncthreads();
    {   // ---------------- pass A: this warp's 16 rows ----------------
      const int sr = w * 16, ia = rb + sr + g, ib = ia + 8;
      const bool ok0 = ia < r1, ok1 = ib < r1;
      uint32_t da[4][4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) xt_ldsm(da[kk], Ob + (sr + (lane & 7) + ((lane >> 3) & 1) * 8) * XT_KP + kk * 16 + (lane >> 4) * 8);
      float dp[8][4];
#pragma unroll
      for (int jt = 0; jt < 8; ++jt) dp[jt][0] = dp[jt][1] = dp[jt][2] = dp[jt][3] = 0.f;
      const __nv_bfloat16 *vrow = Vs + (lane & 7) * XT_KP + (lane >> 3) * 8;
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
#pragma unroll
        for (int jt = 0; jt < 8; jt += 2) {
          uint32_t ra[4], rb4[4];
          xt_ldsm(ra, vrow + jt * 8 * XT_KP + h2 * 32);
          xt_ldsm(rb4, vrow + (jt + 1) * 8 * XT_KP + h2 * 32);
          xt_mma(dp[jt], da[2 * h2], ra[0], ra[1]);
          xt_mma(dp[jt + 1], da[2 * h2], rb4[0], rb4[1]);
          xt_mma(dp[jt], da[2 * h2 + 1], ra[2], ra[3]);
          xt_mma(dp[jt + 1], da[2 * h2 + 1], rb4[2], rb4[3]);
        }
      }
      const long long pbase0 = (((long long)b * heads + h) * Nq + ia) * Nk, pbase1 = (((long long)b * heads + h) * Nq + ib) * Nk;
      float pd[8][4];                                            // P * keep
      float dot0 = 0.f, dot1 = 0.f;
#pragma unroll
      for (int jt = 0; jt < 8; ++jt) {
        const int c = jt * 8 + 2 * t;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int cc = c + (e & 1);
          const bool ok = ((e < 2) ? ok0 : ok1) && cc < Nk;
          float pv = 0.f, kf = 1.f;
          if (ok) {
            const long long idx = ((e < 2) ? pbase0 : pbase1) + cc;
            pv = __bfloat162float(probs[idx]);
            if (pdrop > 0.f) kf = cf_keep(dkey, (unsigned long long)idx, pdrop, ikeep);
          }
          pd[jt][e] = pv * kf;                                   // dropped probability; d(softmax out) = dP * keep
          if (e < 2) dot0 = fmaf(dp[jt][e], pd[jt][e], dot0); else dot1 = fmaf(dp[jt][e], pd[jt][e], dot1);
          dp[jt][e] = ok ? dp[jt][e] * kf : 0.f;                 // dP wrt the softmax output
          pd[jt][e] = ok ? pd[jt][e] : 0.f;
        }
      }
      dot0 += __shfl_xor_sync(0xffffffffu, dot0, 1); dot0 += __shfl_xor_sync(0xffffffffu, dot0, 2);
      dot1 += __shfl_xor_sync(0xffffffffu, dot1, 1); dot1 += __shfl_xor_sync(0xffffffffu, dot1, 2);
      uint32_t sk[8][2];
#pragma unroll
      for (int jt = 0; jt < 8; ++jt) {
        const int c = jt * 8 + 2 * t;
        // P itself = pd / keep where kept; where dropped (kf = 0) dS is 0 anyway... but P*(dP - dot) needs the UNdropped P: reload
        float ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int cc = c + (e & 1);
          const bool ok = ((e < 2) ? ok0 : ok1) && cc < Nk;
          const float pv = ok ? __bfloat162float(probs[((e < 2) ? pbase0 : pbase1) + cc]) : 0.f;
          ds[e] = pv * (dp[jt][e] - ((e < 2) ? dot0 : dot1)) * scale;
        }
        sk[jt][0] = xt_pack(ds[0], ds[1]); sk[jt][1] = xt_pack(ds[2], ds[3]);
        *reinterpret_cast<uint32_t *>(Sb + (sr + g) * XT_KP + c) = sk[jt][0];
        *reinterpret_cast<uint32_t *>(Sb + (sr + g + 8) * XT_KP + c) = sk[jt][1];
        *reinterpret_cast<uint32_t *>(Pb + (sr + g) * XT_KP + c) = xt_pack(pd[jt][0], pd[jt][1]);
        *reinterpret_cast<uint32_t *>(Pb + (sr + g + 8) * XT_KP + c) = xt_pack(pd[jt][2], pd[jt][3]);
      }
      float o[8][4];
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {                           // dQ = dS K
        const uint32_t a[4] = {sk[2 * kk][0], sk[2 * kk][1], sk[2 * kk + 1][0], sk[2 * kk + 1][1]};
#pragma unroll
        for (int dp2 = 0; dp2 < 4; ++dp2) {
          uint32_t r[4];
          xt_ldsm_t(r, Ks + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * XT_KP + (dp2 * 2 + (lane >> 4)) * 8);
          xt_mma(o[2 * dp2], a, r[0], r[1]);
          xt_mma(o[2 * dp2 + 1], a, r[2], r[3]);
        }
      }
      __nv_bfloat16 *o0 = dq + ((long long)b * Nq + ia) * lddq + h * 64 + 2 * t, *o1 = dq + ((long long)b * Nq + ib) * lddq + h * 64 + 2 * t;
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        if (ok0) *reinterpret_cast<uint32_t *>(o0 + dt * 8) = xt_pack(o[dt][0], o[dt][1]);
        if (ok1) *reinterpret_cast<uint32_t *>(o1 + dt * 8) = xt_pack(o[dt][2], o[dt][3]);
      }
    }
    __syncthreads();
    // ---------------- pass B: dK / dV tile of this warp over the batch's 128 rows ----------------
    const int nk16 = (min(r1 - rb, XT_ROWS) + 15) >> 4;          // 16-row steps that hold rows of this batch (the rest is zero)
#pragma unroll 2
    for (int kk = 0; kk < nk16; ++kk) {
      uint32_t as[4], ap[4];
      const int aoff = (kk * 16 + (lane & 7) + (lane >> 4) * 8) * XT_KP + mt * 16 + ((lane >> 3) & 1) * 8;
      xt_ldsm_t(as, Sb + aoff);
      xt_ldsm_t(ap, Pb + aoff);
      const int boff = (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * XT_KP + dhalf * 32 + (lane >> 4) * 8;
#pragma unroll
      for (int dp2 = 0; dp2 < 2; ++dp2) {
        uint32_t rq[4], ro[4];
        xt_ldsm_t(rq, Qb + boff + dp2 * 16);
        xt_ldsm_t(ro, Ob + boff + dp2 * 16);
        xt_mma(aK[2 * dp2], as, rq[0], rq[1]);
        xt_mma(aV[2 * dp2], ap, ro[0], ro[1]);
        xt_mma(aK[2 * dp2 + 1], as, rq[2], rq[3]);
        xt_mma(aV[2 * dp2 + 1], ap, ro[2], ro[3]);
      }
    }
  }
  float *dkvb = dkv + (long long)b * Nk * (2 * inner) + h * 64;
#pragma unroll
  for (int nt8 = 0; nt8 < 4; ++nt8) {
    const int d = dhalf * 32 + nt8 * 8 + 2 * t;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int key = mt * 16 + g + ((e >> 1) ? 8 : 0);
      if (key < Nk) {
        atomicAdd(dkvb + (long long)key * (2 * inner) + d + (e & 1), aK[nt8][e]);
        atomicAdd(dkvb + (long long)key * (2 * inner) + inner + d + (e & 1), aV[nt8][e]);
      }
    }
  }
}

// dh == 64 variant: the same per-row pass (one warp per query row: dP, softmax backward, dQ) writes the row's q, dO, dropped P and
// dS into a 32-row shared-memory batch; then ALL threads fold the batch into register accumulators - thread = (16 keys, one d):
// dK[j][d] += sum_r dS[r][j] q[r][d], dV[j][d] += sum_r Pdrop[r][j] dO[r][d] - with float4 broadcast reads.  (The generic kernel above
// issues 4 shared-memory atomics per (row, key, d): 9 ms of the 60 ms ChangeFormer step.)
constexpr int XB_ROWS = 32, XB_D = 64;

template <typename T>
__global__ void __launch_bounds__(XA_WARPS * 32)
xattention_bwd64_kernel(int Nq, int Nk, int heads, const T *__restrict__ q, long long ldq, const T *__restrict__ kv, long long ldkv,
                        const T *__restrict__ probs, const T *__restrict__ dout, long long ldo, float scale, T *__restrict__ dq, long long lddq,
                        float *__restrict__ dkv, int rows_per_cta, float pdrop, unsigned long long seed, const int *step_ptr, int site) {
  extern __shared__ __align__(16) float sm[];
  constexpr int dh = XB_D, P = XB_D + 1;
  const unsigned long long dkey = cf_key(seed, step_ptr, site);
  const float ikeep = 1.f / (1.f - pdrop);
  float *Sb = sm, *Pb = Sb + XB_ROWS * XA_KMAX, *Qb = Pb + XB_ROWS * XA_KMAX, *Ob = Qb + XB_ROWS * XB_D, *Ks = Ob + XB_ROWS * XB_D,
        *Vs = Ks + Nk * P;
  const int b = blockIdx.z, h = blockIdx.y, inner = heads * dh;
  const T *kvb = kv + (long long)b * Nk * ldkv + h * dh;
  xa_load(Ks, kvb, ldkv, Nk, dh);
  xa_load(Vs, kvb + inner, ldkv, Nk, dh);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int td = threadIdx.x & 63, tj = (threadIdx.x >> 6) * 16;          // phase 2: this thread's d and its 16-key block
  float aK[16], aV[16];
#pragma unroll
  for (int m = 0; m < 16; ++m) { aK[m] = 0.f; aV[m] = 0.f; }
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(Nq, r0 + rows_per_cta);
  for (int rb = r0; rb < r1; rb += XB_ROWS) {
    __syncthreads();                                   // K/V loaded (first batch); phase 2 of the previous batch done
    for (int slot = w; slot < XB_ROWS; slot += XA_WARPS) {
      const int i = rb + slot;
      float *qw = Qb + slot * XB_D, *ow = Ob + slot * XB_D, *pw = Pb + slot * XA_KMAX, *sw = Sb + slot * XA_KMAX;
      if (i >= r1) {                                   // ragged tail: an all-zero row adds nothing in phase 2
        for (int j = lane; j < XA_KMAX; j += 32) { pw[j] = 0.f; sw[j] = 0.f; }
        for (int d = lane; d < dh; d += 32) { qw[d] = 0.f; ow[d] = 0.f; }
        continue;
      }
      const T *qrow = q + ((long long)b * Nq + i) * ldq + h * dh;
      const T *dorow = dout + ((long long)b * Nq + i) * ldo + h * dh;
      const T *prow = probs + (((long long)b * heads + h) * Nq + i) * Nk;
      for (int d = lane; d < dh; d += 32) { qw[d] = Cvt<T>::ld(qrow + d); ow[d] = Cvt<T>::ld(dorow + d); }
      __syncwarp();
      float dp[2], pv[2], kf[2], dot = 0.f;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = jj * 32 + lane;
        dp[jj] = 0.f; pv[jj] = 0.f; kf[jj] = 1.f;
        if (j < Nk) {
          float a = 0.f;
          const float *vp = Vs + j * P;
#pragma unroll 16
          for (int d = 0; d < dh; ++d) a = fmaf(ow[d], vp[d], a);
          if (pdrop > 0.f) kf[jj] = cf_keep(dkey, (unsigned long long)((((long long)b * heads + h) * Nq + i) * Nk + j), pdrop, ikeep);
          dp[jj] = a * kf[jj]; pv[jj] = Cvt<T>::ld(prow + j);
          dot += dp[jj] * pv[jj];
        }
      }
      dot = warp_sum(dot);
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = jj * 32 + lane;
        const bool in = j < Nk;
        pw[j] = in ? pv[jj] * kf[jj] : 0.f;
        sw[j] = in ? pv[jj] * (dp[jj] - dot) * scale : 0.f;
      }
      __syncwarp();
      T *dqrow = dq + ((long long)b * Nq + i) * lddq + h * dh;
      for (int d = lane; d < dh; d += 32) {
        float o = 0.f;
        for (int j = 0; j < Nk; ++j) o = fmaf(sw[j], Ks[j * P + d], o);
        Cvt<T>::st(dqrow + d, o);
      }
    }
    __syncthreads();
    if (tj < Nk) {
#pragma unroll 4
      for (int r = 0; r < XB_ROWS; ++r) {
        const float qd = Qb[r * XB_D + td], od = Ob[r * XB_D + td];
        const float4 *s4 = reinterpret_cast<const float4 *>(Sb + r * XA_KMAX + tj), *p4 = reinterpret_cast<const float4 *>(Pb + r * XA_KMAX + tj);
#pragma unroll
        for (int m4 = 0; m4 < 4; ++m4) {
          const float4 sv = s4[m4], pv4 = p4[m4];
          aK[4 * m4] = fmaf(sv.x, qd, aK[4 * m4]); aK[4 * m4 + 1] = fmaf(sv.y, qd, aK[4 * m4 + 1]);
          aK[4 * m4 + 2] = fmaf(sv.z, qd, aK[4 * m4 + 2]); aK[4 * m4 + 3] = fmaf(sv.w, qd, aK[4 * m4 + 3]);
          aV[4 * m4] = fmaf(pv4.x, od, aV[4 * m4]); aV[4 * m4 + 1] = fmaf(pv4.y, od, aV[4 * m4 + 1]);
          aV[4 * m4 + 2] = fmaf(pv4.z, od, aV[4 * m4 + 2]); aV[4 * m4 + 3] = fmaf(pv4.w, od, aV[4 * m4 + 3]);
        }
      }
    }
  }
  float *dkvb = dkv + (long long)b * Nk * (2 * inner) + h * dh;
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    const int j = tj + m;
    if (j < Nk) {
      atomicAdd(dkvb + (long long)j * (2 * inner) + td, aK[m]);
      atomicAdd(dkvb + (long long)j * (2 * inner) + inner + td, aV[m]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Depth-wise 3x3 convolution (padding 1) on dense NHWC [N,H,W,C]; weights [9][C] fp32 (tap-major), bias [C].
// ---------------------------------------------------------------------------------------------------------
template <typename T, bool FLIP>
__global__ void __launch_bounds__(256)
dwconv_kernel(int N, int H, int W, int C, const T *__restrict__ x, const float *__restrict__ w9, const float *__restrict__ bias, T *__restrict__ y) {
  const int CV = C / 8;
  const unsigned total = (unsigned)N * H * W * CV;            // < 2^31 (checked by the launcher): 32-bit index arithmetic - with
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {   // 64-bit div/mod the kernel was issue-bound
    const int c = (int)(i % CV) * 8; const unsigned p = i / CV;
    const int w = (int)(p % W), h = (int)((p / W) % H); const long long n = p / ((unsigned)W * H);
    float acc[8];
    if (bias) ld8(bias + c, acc); else {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int dy = t / 3 - 1, dx = t % 3 - 1;
      const int hh = FLIP ? h - dy : h + dy, ww = FLIP ? w - dx : w + dx;
      if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
      float f[8], wv[8];
      ld8(x + ((n * H + hh) * W + ww) * C + c, f);
      ld8(w9 + (long long)t * C + c, wv);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(f[k], wv[k], acc[k]);
    }
    st8(y + (long long)p * C + c, acc);
  }
}

// dw9[t][c] += sum_px dy[px][c] * x[px+t][c];  dbias[c] += sum_px dy[px][c].   grid (pixel chunks, C/256)
template <typename T>
__global__ void __launch_bounds__(256)
dwconv_wgrad_kernel(int N, int H, int W, int C, const T *__restrict__ x, const T *__restrict__ dy, float *__restrict__ dw9, float *__restrict__ dbias) {
  __shared__ float red[10][256];
  for (int i = threadIdx.x; i < 10 * 256; i += blockDim.x) (&red[0][0])[i] = 0.f;
  __syncthreads();
  const int cbase = blockIdx.y * 256, cw = min(256, C - cbase), CV = cw / 8;
  const int tx = threadIdx.x % CV, ty = threadIdx.x / CV, rows = blockDim.x / CV;
  const int c = cbase + tx * 8;
  float acc[10][8];
#pragma unroll
  for (int t = 0; t < 10; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[t][k] = 0.f;
  const long long npix = (long long)N * H * W;
  if (ty < rows) {
    for (unsigned p = blockIdx.x * rows + ty; p < (unsigned)npix; p += gridDim.x * rows) {
      const int w = (int)(p % W), h = (int)((p / W) % H); const long long n = p / ((unsigned)W * H);
      float g[8];
      ld8(dy + (long long)p * C + c, g);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[9][k] += g[k];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int hh = h + t / 3 - 1, ww = w + t % 3 - 1;
        if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
        float f[8];
        ld8(x + ((n * H + hh) * W + ww) * C + c, f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[t][k] = fmaf(g[k], f[k], acc[t][k]);
      }
    }
#pragma unroll
    for (int t = 0; t < 10; ++t)
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(&red[t][tx * 8 + k], acc[t][k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 10 * cw; i += blockDim.x) {
    const int t = i / cw, cc = i % cw;
    if (t < 9) atomicAdd(dw9 + (long long)t * C + cbase + cc, red[t][cc]);
    else if (dbias) atomicAdd(dbias + cbase + cc, red[9][cc]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// 2x2-block variants (round 2b).  The kernels above spend 27 loads per output vector (9 taps x (16 B of x + 32 B of weights)) and
// four integer divisions: 3.2 + 1.1 ms per ChangeFormer step against an HBM floor of ~0.6 ms.  Here a warp owns 32 channel vectors
// (512 contiguous bytes per pixel) and walks 2x2 output blocks: the 4x4 input window is loaded ONCE (16 loads for 4 outputs instead of
// 36), every input vector feeds the up-to-four outputs it belongs to, and the 9 x 8 weights of the thread's channels (forward / data
// gradient) or its 9 x 8 partial weight gradients stay in registers for the thread's lifetime (the channel vector of a thread is fixed).
// ---------------------------------------------------------------------------------------------------------
template <typename T, bool FLIP>
__global__ void __launch_bounds__(256, 2)
dwconv_block_kernel(int N, int H, int W, int C, const T *__restrict__ x, const float *__restrict__ w9, const float *__restrict__ bias, T *__restrict__ y) {
  const int CV = C / 8, cv = blockIdx.y * 32 + (threadIdx.x & 31);
  if (cv >= CV) return;                                     // no barriers below
  const int c = cv * 8;
  float wr[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t) ld8(w9 + (long long)(FLIP ? 8 - t : t) * C + c, wr[t]);      // data gradient = correlation with the flipped taps
  float bs[8];
  if (bias) ld8(bias + c, bs); else {
#pragma unroll
    for (int k = 0; k < 8; ++k) bs[k] = 0.f;
  }
  const int HB = (H + 1) >> 1, WB = (W + 1) >> 1;
  const unsigned nblk = (unsigned)N * HB * WB;
  for (unsigned q = blockIdx.x * 8 + (threadIdx.x >> 5); q < nblk; q += gridDim.x * 8) {
    const int wb = (int)(q % WB), hb = (int)((q / WB) % HB); const long long n = q / ((unsigned)WB * HB);
    const int h0 = 2 * hb, w0 = 2 * wb;
    // every output starts from the bias and adds its taps in the order of dwconv_kernel (t = 0..8 of the UNflipped weights): the results
    // are bit-identical to the one-output-per-thread kernel, so the end-to-end drift pins do not move with the kernel choice
    float acc[2][2][8];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[a][b][k] = bs[k];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int r = FLIP ? 3 - rr : rr;
      const int hh = h0 - 1 + r;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int ss = 0; ss < 4; ++ss) {
        const int sx = FLIP ? 3 - ss : ss;
        const int ww = w0 - 1 + sx;
        if (ww < 0 || ww >= W) continue;
        float f[8];
        ld8(x + ((n * H + hh) * W + ww) * C + c, f);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          if (r - a < 0 || r - a > 2) continue;
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            if (sx - b < 0 || sx - b > 2) continue;
            const int t = (r - a) * 3 + (sx - b);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[a][b][k] = fmaf(f[k], wr[t][k], acc[a][b][k]);
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
        if (h0 + a < H && w0 + b < W) st8(y + ((n * H + h0 + a) * W + w0 + b) * C + c, acc[a][b]);
  }
}

// dw9[t][c] += sum_px dy[px][c] * x[px+t][c];  dbias[c] += sum_px dy[px][c]: the same 2x2 blocks, 4 dy vectors + the 4x4 x window per block
template <typename T>
__global__ void __launch_bounds__(256, 2)
dwconv_wgrad_block_kernel(int N, int H, int W, int C, const T *__restrict__ x, const T *__restrict__ dy, float *__restrict__ dw9,
                          float *__restrict__ dbias) {
  __shared__ float red[10][256];
  for (int i = threadIdx.x; i < 10 * 256; i += blockDim.x) (&red[0][0])[i] = 0.f;
  __syncthreads();
  const int CV = C / 8, lane = threadIdx.x & 31, cv = blockIdx.y * 32 + lane;
  const bool live = cv < CV;
  const int c = cv * 8;
  float acc[9][8], ab[8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[t][k] = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) ab[k] = 0.f;
  const int HB = (H + 1) >> 1, WB = (W + 1) >> 1;
  const unsigned nblk = (unsigned)N * HB * WB;
  if (live) {
    for (unsigned q = blockIdx.x * 8 + (threadIdx.x >> 5); q < nblk; q += gridDim.x * 8) {
      const int wb = (int)(q % WB), hb = (int)((q / WB) % HB); const long long n = q / ((unsigned)WB * HB);
      const int h0 = 2 * hb, w0 = 2 * wb;
      float g[2][2][8];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          if (h0 + a < H && w0 + b < W) {
            ld8(dy + ((n * H + h0 + a) * W + w0 + b) * C + c, g[a][b]);
#pragma unroll
            for (int k = 0; k < 8; ++k) ab[k] += g[a][b][k];
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) g[a][b][k] = 0.f;
          }
        }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int hh = h0 - 1 + r;
        if (hh < 0 || hh >= H) continue;
#pragma unroll
        for (int sx = 0; sx < 4; ++sx) {
          const int ww = w0 - 1 + sx;
          if (ww < 0 || ww >= W) continue;
          float f[8];
          ld8(x + ((n * H + hh) * W + ww) * C + c, f);
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            if (r - a < 0 || r - a > 2) continue;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
              if (sx - b < 0 || sx - b > 2) continue;
              const int t = (r - a) * 3 + (sx - b);
#pragma unroll
              for (int k = 0; k < 8; ++k) acc[t][k] = fmaf(g[a][b][k], f[k], acc[t][k]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(&red[t][lane * 8 + k], acc[t][k]);
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&red[9][lane * 8 + k], ab[k]);
  }
  __syncthreads();
  const int cbase = blockIdx.y * 256, cw = min(256, C - cbase);
  for (int i = threadIdx.x; i < 10 * cw; i += blockDim.x) {
    const int t = i / cw, cc = i % cw;
    if (t < 9) atomicAdd(dw9 + (long long)t * C + cbase + cc, red[t][cc]);
    else if (dbias) atomicAdd(dbias + cbase + cc, red[9][cc]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Shared-memory tile variants for bf16 (round 2c, the default).  The register-window kernels above keep one 4x4 window per warp in flight:
// the loads of the next block are issued only after the previous block's arithmetic, so a launch is bound by memory latency (1.3 TB/s of
// algorithmic traffic).  Here a CTA owns a 256-channel group (lane = 8 channels, 512 contiguous bytes per pixel) and walks 7x7 output
// tiles (56 / 28 / 14 / 7 are multiples of 7): the 9x9 input halo tile (41.5 KB) of the NEXT tile streams into the second shared-memory
// buffer with cp.async (16 bytes per request, zero fill outside the image = the padding) while the warps compute the current one, i.e. up to
// 2 x 41.5 KB per SM are in flight without holding registers.  Every output starts from the bias and adds its nine taps with fmaf in the
// order of dwconv_kernel (out-of-image taps add 0 * w): results are bit-identical to the one-output-per-thread kernel.
// ---------------------------------------------------------------------------------------------------------
constexpr int DWT = 7, DWI = DWT + 2;
constexpr int DW_PIX_BYTES = 512;                                   // 32 lanes x 8 bf16
constexpr int DW_IN_BYTES = DWI * DWI * DW_PIX_BYTES;               // 41 472
constexpr int DW_OUT_BYTES = DWT * DWT * DW_PIX_BYTES;              // 25 088

// 8 bf16 of a staged pixel -> 8 floats, one ALU instruction per value (the low half shifts up, the high half is masked in place)
__device__ __forceinline__ void dw_ld8s(const unsigned char *p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4 *>(p);
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}

// a RH x RW pixel rectangle whose top-left pixel is (hs, ws) of image n -> shared memory, pixel-major; warps take whole pixels
template <int RH, int RW>
__device__ __forceinline__ void dw_stage_rect(uint32_t sdst, const __nv_bfloat16 *__restrict__ src, long long n, int hs, int ws, int H, int W, int C, int c,
                                              bool cok) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int p = wid; p < RH * RW; p += nw) {
    const int hh = hs + p / RW, ww = ws + p % RW;
    const bool ok = cok && hh >= 0 && hh < H && ww >= 0 && ww < W;
    const __nv_bfloat16 *g = ok ? src + ((n * H + hh) * W + ww) * C + c : src;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdst + (uint32_t)(p * DW_PIX_BYTES + lane * 16)), "l"(g), "r"(ok ? 16u : 0u) : "memory");
  }
}

template <bool FLIP>
__global__ void __launch_bounds__(256, 2)
dwconv_tile_kernel(int N, int H, int W, int C, const __nv_bfloat16 *__restrict__ x, const float *__restrict__ w9, const float *__restrict__ bias,
                   __nv_bfloat16 *__restrict__ y) {
  extern __shared__ __align__(16) unsigned char dw_smem[];
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(dw_smem);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c = blockIdx.y * 256 + lane * 8;
  const bool cok = c < C;
  float wr[9][8], bs[8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    if (cok) ld8(w9 + (long long)t * C + c, wr[t]); else {
#pragma unroll
      for (int k = 0; k < 8; ++k) wr[t][k] = 0.f;
    }
  }
  if (cok && bias) ld8(bias + c, bs); else {
#pragma unroll
    for (int k = 0; k < 8; ++k) bs[k] = 0.f;
  }
  const int TH = (H + DWT - 1) / DWT, TW = (W + DWT - 1) / DWT;
  const int ntile = N * TH * TW;
  auto stage = [&](int tile, int buf) {
    const int tw = tile % TW, th = (tile / TW) % TH, n = tile / (TW * TH);
    dw_stage_rect<DWI, DWI>(sbase + (uint32_t)buf * DW_IN_BYTES, x, n, th * DWT - 1, tw * DWT - 1, H, W, C, c, cok);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if ((int)blockIdx.x < ntile) stage(blockIdx.x, 0);
  int buf = 0;
  for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x, buf ^= 1) {
    const bool more = tile + (int)gridDim.x < ntile;
    if (more) stage(tile + gridDim.x, buf ^ 1);                  // buffer buf^1 was released by the barrier that ended the previous iteration
    if (more) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const int tw = tile % TW, th = (tile / TW) % TH; const long long n = tile / (TW * TH);
    const int h0 = th * DWT, w0 = tw * DWT;
    const unsigned char *sb = dw_smem + buf * DW_IN_BYTES + lane * 16;
    if (cok) {
      // pixel p = a * 7 + b of the tile; a warp takes p = wid, wid + 8, ...: (a, b) advance by (1, 1) with a carry out of b
      for (int a = wid / DWT, b = wid % DWT; a < DWT; ++a, ++b) {
        if (b >= DWT) { b -= DWT; if (++a >= DWT) break; }
        if (h0 + a >= H || w0 + b >= W) continue;
        const unsigned char *sp = sb + ((a + 1) * DWI + b + 1) * DW_PIX_BYTES;      // the centre tap
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = bs[k];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int dy = t / 3 - 1, dx = t % 3 - 1;
          float f[8];
          dw_ld8s(sp + ((FLIP ? -dy : dy) * DWI + (FLIP ? -dx : dx)) * DW_PIX_BYTES, f);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = fmaf(f[k], wr[t][k], acc[k]);
        }
        st8(y + ((n * H + h0 + a) * W + w0 + b) * C + c, acc);
      }
    }
    __syncthreads();                                             // every warp is done with buffer buf before the next iteration refills it
  }
}

// dw9[t][c] += sum_px dy[px][c] * x[px+t][c];  dbias[c] += sum_px dy[px][c]: the x halo tile and the dy tile of the next 7x7 tile stream in
// (2 x 65 KB, one CTA of 16 warps per SM) while the warps accumulate the current one; zero-filled dy pixels / x halos contribute nothing
__global__ void __launch_bounds__(512, 1)
dwconv_wgrad_tile_kernel(int N, int H, int W, int C, const __nv_bfloat16 *__restrict__ x, const __nv_bfloat16 *__restrict__ dy, float *__restrict__ dw9,
                         float *__restrict__ dbias) {
  extern __shared__ __align__(16) unsigned char dw_smem[];
  __shared__ float red[10][256];
  for (int i = threadIdx.x; i < 10 * 256; i += blockDim.x) (&red[0][0])[i] = 0.f;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(dw_smem);
  constexpr int STAGE = DW_IN_BYTES + DW_OUT_BYTES;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c = blockIdx.y * 256 + lane * 8;
  const bool cok = c < C;
  float acc[9][8], ab[8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[t][k] = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) ab[k] = 0.f;
  const int TH = (H + DWT - 1) / DWT, TW = (W + DWT - 1) / DWT;
  const int ntile = N * TH * TW;
  auto stage = [&](int tile, int buf) {
    const int tw = tile % TW, th = (tile / TW) % TH, n = tile / (TW * TH);
    dw_stage_rect<DWI, DWI>(sbase + (uint32_t)buf * STAGE, x, n, th * DWT - 1, tw * DWT - 1, H, W, C, c, cok);
    dw_stage_rect<DWT, DWT>(sbase + (uint32_t)buf * STAGE + DW_IN_BYTES, dy, n, th * DWT, tw * DWT, H, W, C, c, cok);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if ((int)blockIdx.x < ntile) stage(blockIdx.x, 0);
  int buf = 0;
  for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x, buf ^= 1) {
    const bool more = tile + (int)gridDim.x < ntile;
    if (more) stage(tile + gridDim.x, buf ^ 1);
    if (more) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const unsigned char *sx = dw_smem + buf * STAGE + lane * 16, *sg = sx + DW_IN_BYTES;
    if (cok) {
      // pixel p = a * 7 + b; a warp takes p = wid, wid + 16, ...: (a, b) advance by (2, 2) with a carry out of b
      for (int a = wid / DWT, b = wid % DWT; a < DWT; a += 2, b += 2) {
        if (b >= DWT) { b -= DWT; if (++a >= DWT) break; }
        float g[8];
        dw_ld8s(sg + (a * DWT + b) * DW_PIX_BYTES, g);
#pragma unroll
        for (int k = 0; k < 8; ++k) ab[k] += g[k];
        const unsigned char *sp = sx + (a * DWI + b) * DW_PIX_BYTES;                 // tap (0, 0) of output pixel (a, b)
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          float f[8];
          dw_ld8s(sp + ((t / 3) * DWI + t % 3) * DW_PIX_BYTES, f);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[t][k] = fmaf(g[k], f[k], acc[t][k]);
        }
      }
    }
    __syncthreads();
  }
  __syncthreads();                                               // also orders the zeroing of red[] for a CTA that had no tile
  if (cok) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(&red[t][lane * 8 + k], acc[t][k]);
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&red[9][lane * 8 + k], ab[k]);
  }
  __syncthreads();
  const int cbase = blockIdx.y * 256, cw = min(256, C - cbase);
  for (int i = threadIdx.x; i < 10 * cw; i += blockDim.x) {
    const int t = i / cw, cc = i % cw;
    if (t < 9) atomicAdd(dw9 + (long long)t * C + cbase + cc, red[t][cc]);
    else if (dbias) atomicAdd(dbias + cbase + cc, red[9][cc]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Bilinear resize of dense NHWC maps, align_corners=False (aten upsample_bilinear2d).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bl_src(int d, float ratio, int in, int &i0, int &i1, float &l1) {
  float s = ratio * ((float)d + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s; if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + ((i0 < in - 1) ? 1 : 0);
  l1 = s - (float)i0;
}

// I = unsigned (item count < 2^31, the launcher's default) or long long for the index decomposition (see im2col_kernel)
template <typename T, typename I>
__global__ void __launch_bounds__(256)
bilinear_nhwc_fwd_kernel(int N, int Hi, int Wi, int Ho, int Wo, int C, const T *__restrict__ src, T *__restrict__ dst, int accumulate) {
  const int CV = C / 8;
  const I total = (I)N * (I)Ho * (I)Wo * (I)CV;
  const float ry = (float)Hi / (float)Ho, rx = (float)Wi / (float)Wo;
  for (I i = (I)blockIdx.x * (I)blockDim.x + (I)threadIdx.x; i < total; i += (I)gridDim.x * (I)blockDim.x) {
    const int c = (int)(i % (I)CV) * 8; const long long p = (long long)(i / (I)CV);
    const I pi = i / (I)CV;
    const int x = (int)(pi % (I)Wo), y = (int)((pi / (I)Wo) % (I)Ho); const long long n = (long long)(pi / ((I)Wo * (I)Ho));
    int y0, y1, x0, x1; float ly, lx;
    bl_src(y, ry, Hi, y0, y1, ly); bl_src(x, rx, Wi, x0, x1, lx);
    float a[8], b[8], cc[8], d[8], o[8];
    const T *base = src + n * Hi * Wi * C + c;
    ld8(base + ((long long)y0 * Wi + x0) * C, a); ld8(base + ((long long)y0 * Wi + x1) * C, b);
    ld8(base + ((long long)y1 * Wi + x0) * C, cc); ld8(base + ((long long)y1 * Wi + x1) * C, d);
    if (accumulate) ld8(dst + p * C + c, o); else {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] += (1.f - ly) * ((1.f - lx) * a[k] + lx * b[k]) + ly * ((1.f - lx) * cc[k] + lx * d[k]);
    st8(dst + p * C + c, o);
  }
}

template <typename T, typename I>
__global__ void __launch_bounds__(256)
bilinear_nhwc_bwd_kernel(int N, int Hi, int Wi, int Ho, int Wo, int C, const T *__restrict__ ddst, T *__restrict__ dsrc, int accumulate) {
  const int CV = C / 8;
  const I total = (I)N * (I)Hi * (I)Wi * (I)CV;
  const float ry = (float)Hi / (float)Ho, rx = (float)Wi / (float)Wo;
  for (I i = (I)blockIdx.x * (I)blockDim.x + (I)threadIdx.x; i < total; i += (I)gridDim.x * (I)blockDim.x) {
    const int c = (int)(i % (I)CV) * 8; const long long p = (long long)(i / (I)CV);
    const I pi = i / (I)CV;
    const int gx = (int)(pi % (I)Wi), gy = (int)((pi / (I)Wi) % (I)Hi); const long long n = (long long)(pi / ((I)Wi * (I)Hi));
    int ylo = (int)floorf(((float)gy - 0.5f) / ry - 0.5f) - 1, yhi = (int)ceilf(((float)gy + 1.5f) / ry - 0.5f) + 1;
    int xlo = (int)floorf(((float)gx - 0.5f) / rx - 0.5f) - 1, xhi = (int)ceilf(((float)gx + 1.5f) / rx - 0.5f) + 1;
    if (gy == 0) ylo = 0;
    if (gy == Hi - 1) yhi = Ho - 1;
    if (gx == 0) xlo = 0;
    if (gx == Wi - 1) xhi = Wo - 1;
    ylo = max(ylo, 0); yhi = min(yhi, Ho - 1); xlo = max(xlo, 0); xhi = min(xhi, Wo - 1);
    float acc[8];
    if (accumulate) ld8(dsrc + p * C + c, acc); else {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    }
    const T *gp = ddst + n * Ho * Wo * C + c;
    for (int y = ylo; y <= yhi; ++y) {
      int y0, y1; float ly;
      bl_src(y, ry, Hi, y0, y1, ly);
      const float wy = ((y0 == gy) ? (1.f - ly) : 0.f) + ((y1 == gy) ? ly : 0.f);
      if (wy == 0.f) continue;
      for (int x = xlo; x <= xhi; ++x) {
        int x0, x1; float lx;
        bl_src(x, rx, Wi, x0, x1, lx);
        const float wx = ((x0 == gx) ? (1.f - lx) : 0.f) + ((x1 == gx) ? lx : 0.f);
        if (wx == 0.f) continue;
        float g[8];
        ld8(gp + ((long long)y * Wo + x) * C, g);
        const float ww = wy * wx;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(ww, g[k], acc[k]);
      }
    }
    st8(dsrc + p * C + c, acc);
  }
}

// ---------------------------------------------------------------------------------------------------------
// ReLU outside a BatchNorm pass, Sigmoid head
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) relu_fwd_kernel(long long n8, const T *__restrict__ x, T *__restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    ld8(x + i * 8, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
    st8(y + i * 8, f);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) relu_bwd_kernel(long long n8, const T *__restrict__ r, const T *__restrict__ g, T *__restrict__ dx) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float f[8], gg[8];
    ld8(r + i * 8, f); ld8(g + i * 8, gg);
#pragma unroll
    for (int k = 0; k < 8; ++k) gg[k] = (f[k] > 0.f) ? gg[k] : 0.f;
    st8(dx + i * 8, gg);
  }
}

// I = unsigned (pixel count < 2^31, the launcher's default) or long long: 64-bit div / mod made these per-pixel kernels issue-bound
template <typename T, typename I>
__device__ __forceinline__ T *cvp(const View &v, I p, int H, int W, int c) {
  const int w = (int)(p % (I)W); const I r = p / (I)W; const int h = (int)(r % (I)H); const long long n = (long long)(r / (I)H);
  return reinterpret_cast<T *>(v.ptr) + (n * v.sn + (long long)h * v.sh + (long long)w * v.sw + c);
}
template <typename T, typename I>
__global__ void __launch_bounds__(256)
sigmoid_head_fwd_kernel(View z, int H, int W, long long NP, int K, float *__restrict__ out) {
  const I HW = (I)H * (I)W;
  for (I p = (I)blockIdx.x * (I)blockDim.x + (I)threadIdx.x; p < (I)NP; p += (I)gridDim.x * (I)blockDim.x) {
    const T *zp = cvp<T, I>(z, p, H, W, 0);
    const long long n = (long long)(p / HW), q = (long long)(p % HW);
    for (int k = 0; k < K; ++k) out[(n * K + k) * (long long)HW + q] = 1.f / (1.f + expf(-Cvt<T>::ld(zp + k)));
  }
}
// VEC: dz.C is a multiple of 8 and every pixel's channel vector is 16-byte aligned - one st8 per 8 channels instead of 8 scalar stores
template <typename T, typename I, bool VEC>
__global__ void __launch_bounds__(256)
sigmoid_head_bwd_kernel(View dz, int H, int W, long long NP, int K, const float *__restrict__ out, const float *__restrict__ dout) {
  const I HW = (I)H * (I)W;
  for (I p = (I)blockIdx.x * (I)blockDim.x + (I)threadIdx.x; p < (I)NP; p += (I)gridDim.x * (I)blockDim.x) {
    T *dp = cvp<T, I>(dz, p, H, W, 0);
    const long long n = (long long)(p / HW), q = (long long)(p % HW);
    if (VEC) {
      for (int c0 = 0; c0 < dz.C; c0 += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = c0 + j;
          v[j] = 0.f;
          if (c < K) { const float y = out[(n * K + c) * (long long)HW + q]; v[j] = dout[(n * K + c) * (long long)HW + q] * y * (1.f - y); }
        }
        st8(dp + c0, v);
      }
    } else {
      for (int c = 0; c < dz.C; ++c) {
        float v = 0.f;
        if (c < K) { const float y = out[(n * K + c) * (long long)HW + q]; v = dout[(n * K + c) * (long long)HW + q] * y * (1.f - y); }
        Cvt<T>::st(dp + c, v);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Stochastic regularisers of the encoder (changeformer.py:652-654: Dropout 0.1, attention dropout 0.1, DropPath 0.1).
// Masks are a pure function of (seed, *step_ptr, site, element index): nothing is stored, the backward regenerates them,
// and a captured CUDA graph draws fresh masks on every replay (the step counter lives in device memory).
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
dropout_apply_kernel(long long n, const T *__restrict__ x, T *__restrict__ y, float p, unsigned long long seed, const int *step_ptr, int site) {
  const unsigned long long key = cf_key(seed, step_ptr, site);
  const float ik = 1.f / (1.f - p);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    Cvt<T>::st(y + i, Cvt<T>::ld(x + i) * cf_keep(key, (unsigned long long)i, p, ik));
}

// MODE 0: x[i] += f(i) * t[i]   (x = x + drop_path(dropout(t)), Block.forward :246-247)
// MODE 1: t[i]  = f(i) * x[i]   (its backward: gradient of the branch output)
// f(i) = dp[sample(i)] * keep(i)/(1-p);  dp may be NULL (no DropPath)
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
branch_kernel(long long n, long long per_sample, T *__restrict__ x, T *__restrict__ t, float p, const float *__restrict__ dp,
              unsigned long long seed, const int *step_ptr, int site) {
  const unsigned long long key = cf_key(seed, step_ptr, site);
  const float ik = 1.f / (1.f - p);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float f = (p > 0.f) ? cf_keep(key, (unsigned long long)i, p, ik) : 1.f;
    if (dp) f *= dp[i / per_sample];
    if (MODE == 0) Cvt<T>::st(x + i, Cvt<T>::ld(x + i) + f * Cvt<T>::ld(t + i));
    else Cvt<T>::st(t + i, f * Cvt<T>::ld(x + i));
  }
}

// 16-byte variants (n, per_sample multiples of 8, n < 2^31): eight elements per thread, the same per-element RNG draw and the same
// arithmetic expressions as the scalar kernels above (bit-identical results), one sample lookup per vector instead of a 64-bit division per
// element.  The scalar kernels moved 2 bytes per thread and iteration and were issue-bound: 2.6 ms of the ChangeFormer bs=32 step.
template <typename T>
__global__ void __launch_bounds__(256)
dropout_apply_vec_kernel(unsigned nv, const T *x, T *y, float p, unsigned long long seed, const int *step_ptr, int site) {
  const unsigned long long key = cf_key(seed, step_ptr, site);
  const float ik = 1.f / (1.f - p);
  for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
    const unsigned long long i0 = (unsigned long long)v * 8;
    float f[8];
    ld8(x + i0, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = f[k] * cf_keep(key, i0 + k, p, ik);
    st8(y + i0, f);
  }
}

template <typename T, int MODE>
__global__ void __launch_bounds__(256)
branch_vec_kernel(unsigned nv, unsigned per_sample_v, T *x, T *t, float p, const float *__restrict__ dp, unsigned long long seed, const int *step_ptr,
                  int site) {
  const unsigned long long key = cf_key(seed, step_ptr, site);
  const float ik = 1.f / (1.f - p);
  for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
    const unsigned long long i0 = (unsigned long long)v * 8;
    const float d = dp ? dp[v / per_sample_v] : 1.f;
    float xv[8], tv[8];
    ld8(x + i0, xv);
    if (MODE == 0) ld8(t + i0, tv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float f = (p > 0.f) ? cf_keep(key, i0 + k, p, ik) : 1.f;
      if (dp) f *= d;
      if (MODE == 0) xv[k] = xv[k] + f * tv[k];
      else tv[k] = f * xv[k];
    }
    if (MODE == 0) st8(x + i0, xv); else st8(t + i0, tv);
  }
}

static inline int cgrid(long long work, int per_block, int cap_mult = 8) {
  long long g = (work + per_block - 1) / per_block;
  const long long cap = (long long)kNumSMs * cap_mult;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}


// ------------------------------------------------------------------------------------------------
// im2col / col2im: the strided convolutions (OverlapPatchEmbed.proj 7x7 s4|s2 p3, Attention.sr k = s) as GEMMs on the tensor-core
// conv engine:  col[(n,ho,wo)][(u*k+v)*C + c] = x[n, ho*s-p+u, wo*s-p+v, c]  (zero outside the image; columns >= k*k*C untouched),
// forward = col x W^T (1x1 ks_conv2d), weight gradient = 1x1 ks_conv2d_wgrad(col, dy), data gradient = col2im(dy x W).
// Pure data movement (HBM-bound): one 16-byte (bf16) / 32-byte (fp32) vector per thread.
// ------------------------------------------------------------------------------------------------
// I = unsigned (the launcher's choice whenever the item count is < 2^31) or long long: the six div / mod of the index decomposition in
// 64-bit arithmetic are ~400 instructions per 16-byte vector - the kernel was issue-bound, not HBM-bound
template <typename T, int VEC, typename I>
__global__ void __launch_bounds__(256)
im2col_kernel(int N, int Hi, int Wi, int Ho, int Wo, int k, int s, int p, View x, T *__restrict__ col, int Kp) {
  const int C = x.C, CV = C / VEC, taps = k * k;
  const I total = (I)N * (I)Ho * (I)Wo * (I)taps * (I)CV;
  const T *xp = reinterpret_cast<const T *>(x.ptr);
  for (I i = (I)blockIdx.x * (I)blockDim.x + (I)threadIdx.x; i < total; i += (I)gridDim.x * (I)blockDim.x) {
    const int cv = (int)(i % (I)CV); const int t = (int)((i / (I)CV) % (I)taps); const I row = i / ((I)CV * (I)taps);
    const int wo = (int)(row % (I)Wo), ho = (int)((row / (I)Wo) % (I)Ho); const long long n = (long long)(row / ((I)Wo * (I)Ho));
    const int h = ho * s - p + t / k, w = wo * s - p + t % k;
    const bool in = h >= 0 && h < Hi && w >= 0 && w < Wi;
    const T *src = xp + n * x.sn + (long long)h * x.sh + (long long)w * x.sw + cv * VEC;
    T *dst = col + (long long)row * Kp + (long long)t * C + cv * VEC;
    if (VEC == 1) dst[0] = in ? src[0] : T(0.f);
    else {
      constexpr int NV = (VEC * (int)sizeof(T)) / 16;
#pragma unroll
      for (int q = 0; q < NV; ++q)
        reinterpret_cast<uint4 *>(dst)[q] = in ? reinterpret_cast<const uint4 *>(src)[q] : make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

// dx[n,h,w,c] (+)= sum over the windows (ho,wo) that contain (h,w) of dcol[(n,ho,wo)][(u*k+v)*C + c], u = h+p-ho*s, v = w+p-wo*s
template <typename T, typename I>
__global__ void __launch_bounds__(256)
col2im_kernel(int N, int Hi, int Wi, int Ho, int Wo, int k, int s, int p, const T *__restrict__ dcol, int Kp, View dx, int accumulate) {
  const int C = dx.C, CV = C / 8;
  const I total = (I)N * (I)Hi * (I)Wi * (I)CV;
  T *dp = reinterpret_cast<T *>(dx.ptr);
  for (I i = (I)blockIdx.x * (I)blockDim.x + (I)threadIdx.x; i < total; i += (I)gridDim.x * (I)blockDim.x) {
    const int c = (int)(i % (I)CV) * 8; const I px = i / (I)CV;
    const int w = (int)(px % (I)Wi), h = (int)((px / (I)Wi) % (I)Hi); const long long n = (long long)(px / ((I)Wi * (I)Hi));
    float acc[8];
    T *out = dp + n * dx.sn + (long long)h * dx.sh + (long long)w * dx.sw + c;
    if (accumulate) ld8(out, acc); else {
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    }
    const int ho_lo = max(0, (h + p - k + s) / s), ho_hi = min(Ho - 1, (h + p) / s);      // ho*s <= h+p <= ho*s + k-1
    const int wo_lo = max(0, (w + p - k + s) / s), wo_hi = min(Wo - 1, (w + p) / s);
    for (int ho = ho_lo; ho <= ho_hi; ++ho) {
      const int u = h + p - ho * s;
      for (int wo = wo_lo; wo <= wo_hi; ++wo) {
        const int v = w + p - wo * s;
        float f[8];
        ld8(dcol + ((n * Ho + ho) * Wo + wo) * Kp + (long long)(u * k + v) * C + c, f);
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] += f[q];
      }
    }
    st8(out, acc);
  }
}

}  // namespace ks

using namespace ks;

#define KS_DISPATCH_T(dtype, CALL)                                  \
  do {                                                              \
    if ((dtype) == KS_F32) { CALL(float); }                         \
    else if ((dtype) == KS_BF16) { CALL(__nv_bfloat16); }           \
    else return KS_EINVAL;                                          \
  } while (0)

static inline bool a16(const void *p) { return ((uintptr_t)p % 16) == 0; }

extern "C" int ks_conv2d_strided(int dtype, int N, int Hi, int Wi, int Ho, int Wo, int ksize, int stride, int pad, const ks_view_t *src,
                                 const void *weight, const float *bias, const ks_view_t *dst, void *stream) {
  KS_CHECK_ARG(src && dst && weight && N > 0 && ksize >= 1 && ksize <= 8 && stride >= 1 && pad >= 0);
  KS_CHECK_ARG(Ho == (Hi + 2 * pad - ksize) / stride + 1 && Wo == (Wi + 2 * pad - ksize) / stride + 1);
  const long long M = (long long)N * Ho * Wo;
  dim3 grid((unsigned)((M + GBM - 1) / GBM), (unsigned)((dst->C + GBN - 1) / GBN));
#define CALL(T) conv_gen_kernel<T, 0><<<grid, 256, 0, (cudaStream_t)stream>>>(to_view(*src), to_view(*dst), N, Hi, Wi, Ho, Wo, ksize, stride, pad, \
                                                                             src->C, dst->C, (const T *)weight, bias, 0)
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_conv2d_strided_dgrad(int dtype, int N, int Hi, int Wi, int Ho, int Wo, int ksize, int stride, int pad, const ks_view_t *dy,
                                       const void *weight, const ks_view_t *dx, int accumulate, void *stream) {
  KS_CHECK_ARG(dy && dx && weight && N > 0 && ksize >= 1 && ksize <= 8 && stride >= 1 && pad >= 0);
  const long long M = (long long)N * ((Hi + stride - 1) / stride) * ((Wi + stride - 1) / stride);        // pixels of the largest phase
  dim3 grid((unsigned)((M + GBM - 1) / GBM), (unsigned)((dx->C + GBN - 1) / GBN), (unsigned)(stride * stride));
#define CALL(T) conv_gen_kernel<T, 1><<<grid, 256, 0, (cudaStream_t)stream>>>(to_view(*dy), to_view(*dx), N, Ho, Wo, Hi, Wi, ksize, stride, pad, \
                                                                             dx->C, dy->C, (const T *)weight, nullptr, accumulate)
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_conv2d_strided_wgrad(int dtype, int N, int Hi, int Wi, int Ho, int Wo, int ksize, int stride, int pad, const ks_view_t *x,
                                       const ks_view_t *dy, float *dw, int accumulate, void *stream) {
  KS_CHECK_ARG(x && dy && dw && N > 0 && ksize >= 1 && ksize <= 8 && stride >= 1 && pad >= 0);
  cudaStream_t st = (cudaStream_t)stream;
  const int taps = ksize * ksize, Cin = x->C, Cout = dy->C;
  if (!accumulate) { cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)taps * Cin * Cout, st); if (e != cudaSuccess) return (int)e; }
  const int bx = (Cin + 63) / 64, by = (Cout + 63) / 64;
  const long long M = (long long)N * Ho * Wo, tiles = (long long)bx * by * taps;
  long long splits = (kNumSMs * 6 + tiles - 1) / tiles;
  const long long maxs = (M + 255) / 256;
  if (splits > maxs) splits = maxs;
  if (splits < 1) splits = 1;
  if (splits * taps > 65535) splits = 65535 / taps;
  dim3 grid(bx, by, (unsigned)(taps * splits));
#define CALL(T) conv_gen_wgrad_kernel<T><<<grid, 256, 0, st>>>(to_view(*x), to_view(*dy), N, Hi, Wi, Ho, Wo, ksize, stride, pad, Cin, Cout, (int)splits, dw)
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

static int xa_cfg(int B, int Nq, int Nk, int heads, int dh, int nbuf, size_t &smem, int &rpc, int &nblk) {
  if (Nk < 1 || Nk > XA_KMAX || dh < 1 || dh > XA_DMAX) return KS_EUNSUPPORTED;
  smem = ((size_t)nbuf * Nk * (dh + 1) + (size_t)2 * XA_WARPS * XA_DMAX + (size_t)2 * XA_WARPS * XA_KMAX) * sizeof(float);
  nblk = 1;
  while ((long long)B * heads * nblk < 4 * kNumSMs && nblk * 64 < Nq) nblk <<= 1;
  rpc = ((Nq + nblk - 1) / nblk + XA_WARPS - 1) / XA_WARPS * XA_WARPS;
  nblk = (Nq + rpc - 1) / rpc;
  return KS_OK;
}

namespace ks {
int xattention_fwd_umma(int B, int Nq, int Nk, int heads, const void *q, long long ldq, const void *kv, long long ldkv, float scale,
                        void *out, long long ldo, void *probs, float pdrop, unsigned long long seed, const int *step_ptr, int site,
                        cudaStream_t st);
}

extern "C" int ks_xattention_fwd(int dtype, int B, int Nq, int Nk, int heads, int dh, const void *q, int64_t ldq, const void *kv, int64_t ldkv,
                                 float scale, void *out, int64_t ldo, void *probs, float pdrop, uint64_t seed, const int *step_ptr, int site,
                                 void *stream) {
  KS_CHECK_ARG(pdrop >= 0.f && pdrop < 1.f);
  KS_CHECK_ARG(B > 0 && Nq > 0 && heads > 0 && q && kv && out && probs);
  // tcgen05 variant (xattention_tc.cu): opt-in.  Measured on B200 (ChangeFormer bs=32, 13 calls per step): 0.915 vs 0.975 ms for the mma.sync
  // kernel below - a tile is 8 UMMAs next to a 49-wide softmax + dropout RNG per row, so the tensor-core path buys 6 % - and the two are
  // numerically equivalent (same error against an fp32 reference), but every rounding-level change of the attention output moves the
  // random-init bf16 gradients the end-to-end drift test pins (64 one-ulp flips: 3 % of the gradient norm), so the default stays put.
  if (dtype == KS_BF16 && dh == 64 && Nk >= 1 && Nk <= 64 && !g_opt.att_simt && g_opt.xatt_umma) {
    const int rc = xattention_fwd_umma(B, Nq, Nk, heads, q, ldq, kv, ldkv, scale, out, ldo, probs, pdrop, seed, step_ptr, site, (cudaStream_t)stream);
    if (rc != KS_EUNSUPPORTED) return rc;
  }
  if (dtype == KS_BF16 && dh == 64 && Nk >= 1 && Nk <= 64 && !g_opt.att_simt && a16(kv) && (ldkv * 2) % 16 == 0 && ((uintptr_t)q % 4) == 0 &&
      (ldq % 2) == 0 && ((uintptr_t)out % 4) == 0 && (ldo % 2) == 0) {
    int nb = 1;
    while ((long long)B * heads * nb < 4 * kNumSMs && nb * 128 < Nq) nb <<= 1;
    const int rpc_t = ((Nq + nb - 1) / nb + 15) / 16 * 16;
    nb = (Nq + rpc_t - 1) / rpc_t;
    xattention_fwd_tc_kernel<<<dim3((unsigned)nb, (unsigned)heads, (unsigned)B), XA_WARPS * 32, 0, (cudaStream_t)stream>>>(
        Nq, Nk, heads, (const __nv_bfloat16 *)q, ldq, (const __nv_bfloat16 *)kv, ldkv, scale, (__nv_bfloat16 *)out, ldo, (__nv_bfloat16 *)probs,
        rpc_t, pdrop, seed, step_ptr, site);
    KS_LAUNCH_RET();
  }
  size_t smem; int rpc, nblk;
  int rc = xa_cfg(B, Nq, Nk, heads, dh, 2, smem, rpc, nblk); if (rc) return rc;
  dim3 grid((unsigned)nblk, (unsigned)heads, (unsigned)B);
#define CALL(T) { cudaError_t e = cudaFuncSetAttribute(xattention_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); \
    if (e != cudaSuccess) return (int)e; \
    xattention_fwd_kernel<T><<<grid, XA_WARPS * 32, smem, (cudaStream_t)stream>>>(Nq, Nk, heads, dh, (const T *)q, ldq, (const T *)kv, ldkv, scale, \
                                                                                   (T *)out, ldo, (T *)probs, rpc, pdrop, seed, step_ptr, site); }
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_xattention_bwd(int dtype, int B, int Nq, int Nk, int heads, int dh, const void *q, int64_t ldq, const void *kv, int64_t ldkv,
                                 const void *probs, const void *dout, int64_t ldo, float scale, void *dq, int64_t lddq, float *dkv,
                                 float pdrop, uint64_t seed, const int *step_ptr, int site, void *stream) {
  KS_CHECK_ARG(pdrop >= 0.f && pdrop < 1.f);
  KS_CHECK_ARG(B > 0 && Nq > 0 && heads > 0 && q && kv && probs && dout && dq && dkv);
  size_t smem; int rpc, nblk;
  int rc = xa_cfg(B, Nq, Nk, heads, dh, 4, smem, rpc, nblk); if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e0 = cudaMemsetAsync(dkv, 0, sizeof(float) * (size_t)B * Nk * 2 * heads * dh, st); if (e0 != cudaSuccess) return (int)e0;
  dim3 grid((unsigned)nblk, (unsigned)heads, (unsigned)B);
  if (dtype == KS_BF16 && dh == 64 && Nk >= 1 && Nk <= 64 && !g_opt.att_simt && a16(kv) && (ldkv * 2) % 16 == 0 && a16(q) && (ldq * 2) % 16 == 0 &&
      a16(dout) && (ldo * 2) % 16 == 0 && ((uintptr_t)dq % 4) == 0 && (lddq % 2) == 0) {
    int nb = 1;
    while ((long long)B * heads * nb < 2 * kNumSMs && nb * XT_ROWS < Nq) nb <<= 1;
    int rpc_t = ((Nq + nb - 1) / nb + 15) / 16 * 16;
    if (rpc_t >= XT_ROWS) rpc_t = rpc_t / XT_ROWS * XT_ROWS;        // whole 128-row batches per CTA (392 rows would run a 4th batch of 8 rows)
    nb = (Nq + rpc_t - 1) / rpc_t;
    const size_t smem_t = (size_t)(2 * 64 + 4 * XT_ROWS) * XT_KP * sizeof(__nv_bfloat16);
    static bool attr = false;
    if (!attr) { cudaError_t e = cudaFuncSetAttribute(xattention_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
                 if (e != cudaSuccess) return (int)e; attr = true; }
    xattention_bwd_tc_kernel<<<dim3((unsigned)nb, (unsigned)heads, (unsigned)B), XA_WARPS * 32, smem_t, st>>>(
        Nq, Nk, heads, (const __nv_bfloat16 *)q, ldq, (const __nv_bfloat16 *)kv, ldkv, (const __nv_bfloat16 *)probs, (const __nv_bfloat16 *)dout, ldo,
        scale, (__nv_bfloat16 *)dq, lddq, dkv, rpc_t, pdrop, seed, step_ptr, site);
    KS_LAUNCH_RET();
  }
  if (dh == XB_D && XA_WARPS * 32 == 4 * XB_D) {
    const size_t smem64 = ((size_t)2 * XB_ROWS * XA_KMAX + (size_t)2 * XB_ROWS * XB_D + (size_t)2 * Nk * (XB_D + 1)) * sizeof(float);
#define CALL(T) { cudaError_t e = cudaFuncSetAttribute(xattention_bwd64_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); \
    if (e != cudaSuccess) return (int)e; \
    xattention_bwd64_kernel<T><<<grid, XA_WARPS * 32, smem64, st>>>(Nq, Nk, heads, (const T *)q, ldq, (const T *)kv, ldkv, (const T *)probs, \
                                                                    (const T *)dout, ldo, scale, (T *)dq, lddq, dkv, rpc, pdrop, seed, step_ptr, site); }
    KS_DISPATCH_T(dtype, CALL);
#undef CALL
    KS_LAUNCH_RET();
  }
#define CALL(T) { cudaError_t e = cudaFuncSetAttribute(xattention_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); \
    if (e != cudaSuccess) return (int)e; \
    xattention_bwd_kernel<T><<<grid, XA_WARPS * 32, smem, st>>>(Nq, Nk, heads, dh, (const T *)q, ldq, (const T *)kv, ldkv, (const T *)probs, \
                                                                 (const T *)dout, ldo, scale, (T *)dq, lddq, dkv, rpc, pdrop, seed, step_ptr, site); }
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_im2col(int dtype, int N, int Hi, int Wi, int Ho, int Wo, int ksize, int stride, int pad, const ks_view_t *x, void *col,
                         int Kp, void *stream) {
  KS_CHECK_ARG(N > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && ksize > 0 && stride > 0 && pad >= 0 && x && x->ptr && col);
  KS_CHECK_ARG(Kp >= ksize * ksize * x->C && (Ho - 1) * stride - pad + ksize - 1 < Hi + pad && (Wo - 1) * stride - pad + ksize - 1 < Wi + pad);
  const View xv = to_view(*x);
  const int es = (dtype == KS_F32) ? 4 : 2;
  const bool vec = xv.C % 8 == 0 && a16(xv.ptr) && a16(col) && (xv.sn * es) % 16 == 0 && (xv.sh * es) % 16 == 0 && (xv.sw * es) % 16 == 0 &&
                   ((long long)Kp * es) % 16 == 0;
  const long long items = (long long)N * Ho * Wo * ksize * ksize * (vec ? xv.C / 8 : xv.C);
  const int grid = cgrid(items, 256);
  const bool i32 = items < (1LL << 31) && !g_opt.cf_scalar;
#define CALL(T) { if (vec && i32) im2col_kernel<T, 8, unsigned><<<grid, 256, 0, (cudaStream_t)stream>>>(N, Hi, Wi, Ho, Wo, ksize, stride, pad, xv, (T *)col, Kp); \
                  else if (vec) im2col_kernel<T, 8, long long><<<grid, 256, 0, (cudaStream_t)stream>>>(N, Hi, Wi, Ho, Wo, ksize, stride, pad, xv, (T *)col, Kp); \
                  else if (i32) im2col_kernel<T, 1, unsigned><<<grid, 256, 0, (cudaStream_t)stream>>>(N, Hi, Wi, Ho, Wo, ksize, stride, pad, xv, (T *)col, Kp); \
                  else im2col_kernel<T, 1, long long><<<grid, 256, 0, (cudaStream_t)stream>>>(N, Hi, Wi, Ho, Wo, ksize, stride, pad, xv, (T *)col, Kp); }
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_col2im(int dtype, int N, int Hi, int Wi, int Ho, int Wo, int ksize, int stride, int pad, const void *dcol, int Kp,
                         const ks_view_t *dx, int accumulate, void *stream) {
  KS_CHECK_ARG(N > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && ksize > 0 && stride > 0 && pad >= 0 && dx && dx->ptr && dcol);
  KS_CHECK_ARG(Kp >= ksize * ksize * dx->C);
  const View dv = to_view(*dx);
  const int es = (dtype == KS_F32) ? 4 : 2;
  if (dv.C % 8 || !a16(dv.ptr) || !a16(dcol) || (dv.sn * es) % 16 || (dv.sh * es) % 16 || (dv.sw * es) % 16 || ((long long)Kp * es) % 16)
    return KS_EUNSUPPORTED;
  const long long items = (long long)N * Hi * Wi * (dv.C / 8);
  const int grid = cgrid(items, 256);
  const bool i32 = items < (1LL << 31) && !g_opt.cf_scalar;
#define CALL(T) { if (i32) col2im_kernel<T, unsigned><<<grid, 256, 0, (cudaStream_t)stream>>>(N, Hi, Wi, Ho, Wo, ksize, stride, pad, (const T *)dcol, Kp, dv, accumulate); \
                  else col2im_kernel<T, long long><<<grid, 256, 0, (cudaStream_t)stream>>>(N, Hi, Wi, Ho, Wo, ksize, stride, pad, (const T *)dcol, Kp, dv, accumulate); }
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_dwconv3x3_fwd(int dtype, int N, int H, int W, int C, const void *x, const float *w9, const float *bias, void *y, void *stream) {
  KS_CHECK_ARG(N > 0 && H > 0 && W > 0 && x && w9 && y);
  if (C % 8 || !a16(x) || !a16(y) || !a16(w9) || (bias && !a16(bias))) return KS_EUNSUPPORTED;
  if ((long long)N * H * W * (C / 8) >= (1LL << 31)) return KS_EUNSUPPORTED;
  const int grid = cgrid((long long)N * H * W * (C / 8), 256);
  const long long nblk = (long long)N * ((H + 1) / 2) * ((W + 1) / 2);
  const int gyb = (C / 8 + 31) / 32;
  int gxb = (kNumSMs * 2 + gyb - 1) / gyb; if (gxb > (nblk + 7) / 8) gxb = (int)((nblk + 7) / 8); if (gxb < 1) gxb = 1;
  if (dtype == KS_BF16 && (g_opt.dwconv_simple == 0 || g_opt.dwconv_simple == 3)) {            // shared-memory tile kernel (the default for bf16)
    static bool attr = false;
    if (!attr) { cudaError_t e = cudaFuncSetAttribute(dwconv_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * DW_IN_BYTES);
                 if (e != cudaSuccess) return (int)e; attr = true; }
    const long long ntile = (long long)N * ((H + DWT - 1) / DWT) * ((W + DWT - 1) / DWT);
    int gxt = (kNumSMs * 2 + gyb - 1) / gyb; if (gxt > ntile) gxt = (int)ntile;
    dwconv_tile_kernel<false><<<dim3(gxt, gyb), 256, 2 * DW_IN_BYTES, (cudaStream_t)stream>>>(N, H, W, C, (const __nv_bfloat16 *)x, w9, bias, (__nv_bfloat16 *)y);
    KS_LAUNCH_RET();
  }
#define CALL(T) { if (g_opt.dwconv_simple == 1) dwconv_kernel<T, false><<<grid, 256, 0, (cudaStream_t)stream>>>(N, H, W, C, (const T *)x, w9, bias, (T *)y); \
    else dwconv_block_kernel<T, false><<<dim3(gxb, gyb), 256, 0, (cudaStream_t)stream>>>(N, H, W, C, (const T *)x, w9, bias, (T *)y); }
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_dwconv3x3_bwd(int dtype, int N, int H, int W, int C, const void *x, const void *dy, const float *w9, void *dx, float *dw9,
                                float *dbias, void *stream) {
  KS_CHECK_ARG(N > 0 && H > 0 && W > 0 && x && dy && w9 && dx && dw9);
  if (C % 8 || !a16(x) || !a16(dy) || !a16(dx) || !a16(w9)) return KS_EUNSUPPORTED;
  if ((long long)N * H * W * (C / 8) >= (1LL << 31)) return KS_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = cgrid((long long)N * H * W * (C / 8), 256);
  const int gy = (C + 255) / 256;
  int gx = (kNumSMs * 4) / gy; if (gx < 1) gx = 1;
  const long long npix = (long long)N * H * W;
  if (gx > (npix + 7) / 8) gx = (int)((npix + 7) / 8);
  const long long nblk = (long long)N * ((H + 1) / 2) * ((W + 1) / 2);
  const int gyb = (C / 8 + 31) / 32;
  int gxb = (kNumSMs * 2 + gyb - 1) / gyb; if (gxb > (nblk + 7) / 8) gxb = (int)((nblk + 7) / 8); if (gxb < 1) gxb = 1;
  if (dtype == KS_BF16 && (g_opt.dwconv_simple == 0 || g_opt.dwconv_simple == 3)) {            // shared-memory tile kernels (the default for bf16)
    static bool attr = false;
    if (!attr) { cudaError_t e = cudaFuncSetAttribute(dwconv_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * DW_IN_BYTES);
                 if (e == cudaSuccess) e = cudaFuncSetAttribute(dwconv_wgrad_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (DW_IN_BYTES + DW_OUT_BYTES));
                 if (e != cudaSuccess) return (int)e; attr = true; }
    const long long ntile = (long long)N * ((H + DWT - 1) / DWT) * ((W + DWT - 1) / DWT);
    int gxt = (kNumSMs * 2 + gyb - 1) / gyb; if (gxt > ntile) gxt = (int)ntile;
    int gxw = (kNumSMs + gyb - 1) / gyb; if (gxw > ntile) gxw = (int)ntile;
    dwconv_tile_kernel<true><<<dim3(gxt, gyb), 256, 2 * DW_IN_BYTES, st>>>(N, H, W, C, (const __nv_bfloat16 *)dy, w9, nullptr, (__nv_bfloat16 *)dx);
    // one-tile images (7x7, the last encoder stage) pay the whole halo for nothing and give a CTA 3-4 tiles: measured 63 us vs 41 us for the
    // register-block kernel at 64 x 7x7 x 2048 (14x14 x 1280: 100 vs 97); option 3 keeps the tile kernel everywhere (tests)
    if (H * W >= 100 || g_opt.dwconv_simple == 3)
      dwconv_wgrad_tile_kernel<<<dim3(gxw, gyb), 512, 2 * (DW_IN_BYTES + DW_OUT_BYTES), st>>>(N, H, W, C, (const __nv_bfloat16 *)x, (const __nv_bfloat16 *)dy, dw9, dbias);
    else
      dwconv_wgrad_block_kernel<__nv_bfloat16><<<dim3(gxb, gyb), 256, 0, st>>>(N, H, W, C, (const __nv_bfloat16 *)x, (const __nv_bfloat16 *)dy, dw9, dbias);
    KS_LAUNCH_RET();
  }
#define CALL(T) { if (g_opt.dwconv_simple == 1) { dwconv_kernel<T, true><<<grid, 256, 0, st>>>(N, H, W, C, (const T *)dy, w9, nullptr, (T *)dx); \
      dwconv_wgrad_kernel<T><<<dim3(gx, gy), 256, 0, st>>>(N, H, W, C, (const T *)x, (const T *)dy, dw9, dbias); } \
    else { dwconv_block_kernel<T, true><<<dim3(gxb, gyb), 256, 0, st>>>(N, H, W, C, (const T *)dy, w9, nullptr, (T *)dx); \
      dwconv_wgrad_block_kernel<T><<<dim3(gxb, gyb), 256, 0, st>>>(N, H, W, C, (const T *)x, (const T *)dy, dw9, dbias); } }
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_bilinear_nhwc_fwd(int dtype, int N, int Hi, int Wi, int Ho, int Wo, int C, const void *src, void *dst, int accumulate, void *stream) {
  KS_CHECK_ARG(N > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && src && dst);
  if (C % 8 || !a16(src) || !a16(dst)) return KS_EUNSUPPORTED;
  const long long items = (long long)N * Ho * Wo * (C / 8);
  const int grid = cgrid(items, 256);
  const bool i32 = items < (1LL << 31) && (long long)N * Hi * Wi * (C / 8) < (1LL << 31) && !g_opt.cf_scalar;
#define CALL(T) { if (i32) bilinear_nhwc_fwd_kernel<T, unsigned><<<grid, 256, 0, (cudaStream_t)stream>>>(N, Hi, Wi, Ho, Wo, C, (const T *)src, (T *)dst, accumulate); \
                  else bilinear_nhwc_fwd_kernel<T, long long><<<grid, 256, 0, (cudaStream_t)stream>>>(N, Hi, Wi, Ho, Wo, C, (const T *)src, (T *)dst, accumulate); }
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_bilinear_nhwc_bwd(int dtype, int N, int Hi, int Wi, int Ho, int Wo, int C, const void *ddst, void *dsrc, int accumulate, void *stream) {
  KS_CHECK_ARG(N > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && ddst && dsrc);
  if (C % 8 || !a16(ddst) || !a16(dsrc)) return KS_EUNSUPPORTED;
  const long long items = (long long)N * Hi * Wi * (C / 8);
  const int grid = cgrid(items, 256, 16);
  const bool i32 = items < (1LL << 31) && !g_opt.cf_scalar;
#define CALL(T) { if (i32) bilinear_nhwc_bwd_kernel<T, unsigned><<<grid, 256, 0, (cudaStream_t)stream>>>(N, Hi, Wi, Ho, Wo, C, (const T *)ddst, (T *)dsrc, accumulate); \
                  else bilinear_nhwc_bwd_kernel<T, long long><<<grid, 256, 0, (cudaStream_t)stream>>>(N, Hi, Wi, Ho, Wo, C, (const T *)ddst, (T *)dsrc, accumulate); }
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_relu_fwd(int dtype, int64_t n, const void *x, void *y, void *stream) {
  KS_CHECK_ARG(n > 0 && x && y);
  if (n % 8 || !a16(x) || !a16(y)) return KS_EUNSUPPORTED;
  const int grid = cgrid(n / 8, 256 * 4);
#define CALL(T) relu_fwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(n / 8, (const T *)x, (T *)y)
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_relu_bwd(int dtype, int64_t n, const void *r, const void *g, void *dx, void *stream) {
  KS_CHECK_ARG(n > 0 && r && g && dx);
  if (n % 8 || !a16(r) || !a16(g) || !a16(dx)) return KS_EUNSUPPORTED;
  const int grid = cgrid(n / 8, 256 * 4);
#define CALL(T) relu_bwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(n / 8, (const T *)r, (const T *)g, (T *)dx)
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_sigmoid_head_fwd(int dtype, int N, int H, int W, const ks_view_t *z, int K, float *out, void *stream) {
  KS_CHECK_ARG(N > 0 && H > 0 && W > 0 && z && z->ptr && out && K > 0 && K <= z->C);
  const long long NP = (long long)N * H * W;
  const int grid = cgrid(NP, 256);
  const bool i32 = NP < (1LL << 31) && !g_opt.cf_scalar;
#define CALL(T) { if (i32) sigmoid_head_fwd_kernel<T, unsigned><<<grid, 256, 0, (cudaStream_t)stream>>>(to_view(*z), H, W, NP, K, out); \
                  else sigmoid_head_fwd_kernel<T, long long><<<grid, 256, 0, (cudaStream_t)stream>>>(to_view(*z), H, W, NP, K, out); }
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_sigmoid_head_bwd(int dtype, int N, int H, int W, const float *out, const float *dout, int K, const ks_view_t *dz, void *stream) {
  KS_CHECK_ARG(N > 0 && H > 0 && W > 0 && dz && dz->ptr && out && dout && K > 0 && K <= dz->C);
  const long long NP = (long long)N * H * W;
  const int grid = cgrid(NP, 256);
  const View dv = to_view(*dz);
  const int es = (dtype == KS_F32) ? 4 : 2;
  const bool i32 = NP < (1LL << 31) && !g_opt.cf_scalar;
  const bool vec = i32 && dv.C % 8 == 0 && a16(dv.ptr) && (dv.sn * es) % 16 == 0 && (dv.sh * es) % 16 == 0 && (dv.sw * es) % 16 == 0;
#define CALL(T) { if (vec) sigmoid_head_bwd_kernel<T, unsigned, true><<<grid, 256, 0, (cudaStream_t)stream>>>(dv, H, W, NP, K, out, dout); \
                  else if (i32) sigmoid_head_bwd_kernel<T, unsigned, false><<<grid, 256, 0, (cudaStream_t)stream>>>(dv, H, W, NP, K, out, dout); \
                  else sigmoid_head_bwd_kernel<T, long long, false><<<grid, 256, 0, (cudaStream_t)stream>>>(dv, H, W, NP, K, out, dout); }
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_dropout_apply(int dtype, int64_t n, const void *x, void *y, float p, uint64_t seed, const int *step_ptr, int site, void *stream) {
  KS_CHECK_ARG(n > 0 && x && y && p >= 0.f && p < 1.f);
  const int grid = cgrid(n, 256 * 8);
  const bool vec = n % 8 == 0 && n < (1LL << 31) && a16(x) && a16(y) && !g_opt.cf_scalar;
#define CALL(T) { if (vec) dropout_apply_vec_kernel<T><<<cgrid(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((unsigned)(n / 8), (const T *)x, (T *)y, p, seed, step_ptr, site); \
                  else dropout_apply_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(n, (const T *)x, (T *)y, p, seed, step_ptr, site); }
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_branch_add(int dtype, int64_t n, int64_t per_sample, void *x, const void *t, float p, const float *droppath, uint64_t seed,
                             const int *step_ptr, int site, void *stream) {
  KS_CHECK_ARG(n > 0 && per_sample > 0 && x && t && p >= 0.f && p < 1.f);
  const int grid = cgrid(n, 256 * 8);
  const bool vec = n % 8 == 0 && per_sample % 8 == 0 && n < (1LL << 31) && a16(x) && a16(t) && !g_opt.cf_scalar;
#define CALL(T) { if (vec) branch_vec_kernel<T, 0><<<cgrid(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((unsigned)(n / 8), (unsigned)(per_sample / 8), (T *)x, (T *)t, p, droppath, seed, step_ptr, site); \
                  else branch_kernel<T, 0><<<grid, 256, 0, (cudaStream_t)stream>>>(n, per_sample, (T *)x, (T *)t, p, droppath, seed, step_ptr, site); }
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}

extern "C" int ks_branch_scale(int dtype, int64_t n, int64_t per_sample, const void *dx, void *dt, float p, const float *droppath, uint64_t seed,
                               const int *step_ptr, int site, void *stream) {
  KS_CHECK_ARG(n > 0 && per_sample > 0 && dx && dt && p >= 0.f && p < 1.f);
  const int grid = cgrid(n, 256 * 8);
  const bool vec = n % 8 == 0 && per_sample % 8 == 0 && n < (1LL << 31) && a16(dx) && a16(dt) && !g_opt.cf_scalar;
#define CALL(T) { if (vec) branch_vec_kernel<T, 1><<<cgrid(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((unsigned)(n / 8), (unsigned)(per_sample / 8), (T *)dx, (T *)dt, p, droppath, seed, step_ptr, site); \
                  else branch_kernel<T, 1><<<grid, 256, 0, (cudaStream_t)stream>>>(n, per_sample, (T *)dx, (T *)dt, p, droppath, seed, step_ptr, site); }
  KS_DISPATCH_T(dtype, CALL);
#undef CALL
  KS_LAUNCH_RET();
}
