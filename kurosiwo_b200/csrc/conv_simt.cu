// CUDA-core implicit-GEMM convolution (forward / data-gradient) and weight-gradient.
// Exact fp32 FMA accumulation: this is the PARITY-mode engine (fp32 storage) and the
// fallback for shapes the tcgen05 engine (conv_tc.cu) does not take (e.g. the 2-channel stem).
//
// Reference call sites: models/snunet.py:15,17 (nn.Conv2d 3x3 p1), :41 (ConvTranspose2d k2 s2,
// expressed as four strided 1x1 phases), :132-144 (torch.cat -> K-dimension view list).
#include "common.cuh"

namespace ks {

constexpr int BM = 64, BN = 64, BK = 16;

__device__ __forceinline__ int find_view(const ViewList &vl, int c) {
  int d = 0;
#pragma unroll
  for (int i = 1; i < KS_MAX_VIEWS; ++i) if (i < vl.n && c >= vl.cstart[i]) d = i;
  return d;
}

template <typename T>
__global__ void __launch_bounds__(256)
conv_simt_kernel(ViewList srcs, ViewList dsts, int dst_acc_mask, int N, int H, int W, int ksize,
                 const T *__restrict__ weight, const float *__restrict__ bias, double *stats) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ float sstat[2][BN];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const long long M = (long long)N * H * W;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int Cin = srcs.cstart[srcs.n], Cout = dsts.cstart[dsts.n];
  const int taps = ksize * ksize, pad = ksize / 2;

  // load-role coordinates
  const int lrow = tid / 4, lk = (tid % 4) * 4;
  const long long lm = m0 + lrow;
  const bool lm_ok = lm < M;
  int ln = 0, lh = 0, lw = 0;
  if (lm_ok) { lw = (int)(lm % W); long long r = lm / W; lh = (int)(r % H); ln = (int)(r / H); }
  const int lco = n0 + lrow;  // weight row for B loads

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int tap = 0; tap < taps; ++tap) {
    const int dy = tap / ksize - pad, dx = tap % ksize - pad;
    const int hh = lh + dy, ww = lw + dx;
    const bool pix_ok = lm_ok && hh >= 0 && hh < H && ww >= 0 && ww < W;
    for (int s = 0; s < srcs.n; ++s) {
      const View &sv = srcs.v[s];
      const T *xp = reinterpret_cast<const T *>(sv.ptr) + ((long long)ln * sv.sn + (long long)hh * sv.sh + (long long)ww * sv.sw);
      const T *wp = weight + ((long long)tap * Cout + lco) * Cin + srcs.cstart[s];
      for (int c0 = 0; c0 < sv.C; c0 += BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = c0 + lk + i;
          float a = 0.f, b = 0.f;
          if (c < sv.C) {
            if (pix_ok) a = Cvt<T>::ld(xp + c);
            if (lco < Cout) b = Cvt<T>::ld(wp + c);
          }
          As[lk + i][lrow] = a;
          Bs[lk + i][lrow] = b;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
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

  // epilogue
  if (stats) { for (int i = tid; i < 2 * BN; i += 256) (&sstat[0][0])[i] = 0.f; __syncthreads(); }
  float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int w = (int)(m % W); long long r = m / W; const int h = (int)(r % H); const int n = (int)(r / H);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= Cout) continue;
      const int d = find_view(dsts, co);
      const View &dv = dsts.v[d];
      T *op = reinterpret_cast<T *>(dv.ptr) + ((long long)n * dv.sn + (long long)h * dv.sh + (long long)w * dv.sw + (co - dsts.cstart[d]));
      float v = acc[i][j] + (bias ? __ldg(bias + co) : 0.f);
      if ((dst_acc_mask >> d) & 1) v += Cvt<T>::ld(op);
      Cvt<T>::st(op, v);
      const float vs = round_as<T>(v);
      s1[j] += vs; s2[j] += vs * vs;
    }
  }
  if (stats) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { atomicAdd(&sstat[0][tx * 4 + j], s1[j]); atomicAdd(&sstat[1][tx * 4 + j], s2[j]); }
    __syncthreads();
    for (int i = tid; i < 2 * BN; i += 256) {
      const int r = i / BN, c = n0 + i % BN;
      if (c < Cout) atomicAdd(stats + (size_t)r * Cout + c, (double)sstat[r][i % BN]);
    }
  }
}

// ---- weight gradient ---------------------------------------------------------------------
// grid: x = 64-channel blocks of the X concat, y = 64-channel blocks of the dY concat,
//       z = taps * splits (split over pixels).  fp32 atomics into dw.
struct BlockTab { short view[64]; short c0[64]; int n; };

template <typename T>
__global__ void __launch_bounds__(256)
wgrad_simt_kernel(ViewList xs, ViewList dys, BlockTab xtab, BlockTab ytab, int N, int H, int W, int ksize,
                  int splits, float *dw) {
  __shared__ float As[BK][BM + 4];  // dY tile [px][co]
  __shared__ float Bs[BK][BN + 4];  // X tile  [px][ci]
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int tap = blockIdx.z / splits, split = blockIdx.z % splits;
  const int pad = ksize / 2, dyy = tap / ksize - pad, dxx = tap % ksize - pad;
  const View &xv = xs.v[xtab.view[blockIdx.x]]; const int xc0 = xtab.c0[blockIdx.x];
  const View &yv = dys.v[ytab.view[blockIdx.y]]; const int yc0 = ytab.c0[blockIdx.y];
  const int Cin = xs.cstart[xs.n], Cout = dys.cstart[dys.n];
  const long long M = (long long)N * H * W;
  const long long per = ((M + splits - 1) / splits + BK - 1) / BK * BK;
  const long long p0 = (long long)split * per, p1 = min(M, p0 + per);
  const int lp = tid / 16, lc = (tid % 16) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long pb = p0; pb < p1; pb += BK) {
    const long long p = pb + lp;
    float a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
    if (p < p1) {
      const int w = (int)(p % W); long long r = p / W; const int h = (int)(r % H); const int n = (int)(r / H);
      const T *yp = reinterpret_cast<const T *>(yv.ptr) + ((long long)n * yv.sn + (long long)h * yv.sh + (long long)w * yv.sw + yc0 + lc);
#pragma unroll
      for (int i = 0; i < 4; ++i) if (yc0 + lc + i < yv.C) a[i] = Cvt<T>::ld(yp + i);
      const int hh = h + dyy, ww = w + dxx;
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
        const T *xp = reinterpret_cast<const T *>(xv.ptr) + ((long long)n * xv.sn + (long long)hh * xv.sh + (long long)ww * xv.sw + xc0 + lc);
#pragma unroll
        for (int i = 0; i < 4; ++i) if (xc0 + lc + i < xv.C) b[i] = Cvt<T>::ld(xp + i);
      }
    }
    *reinterpret_cast<float4 *>(&As[lp][lc]) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4 *>(&Bs[lp][lc]) = make_float4(b[0], b[1], b[2], b[3]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
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
  const int cog0 = dys.cstart[ytab.view[blockIdx.y]] + yc0, cig0 = xs.cstart[xtab.view[blockIdx.x]] + xc0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int col = yc0 + ty * 4 + i;
    if (col >= yv.C) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cil = xc0 + tx * 4 + j;
      if (cil >= xv.C) continue;
      atomicAdd(dw + ((long long)tap * Cout + (cog0 + ty * 4 + i)) * Cin + (cig0 + tx * 4 + j), acc[i][j]);
    }
  }
}

static int make_block_tab(const ViewList &vl, BlockTab &tab) {
  tab.n = 0;
  for (int v = 0; v < vl.n; ++v)
    for (int c0 = 0; c0 < vl.v[v].C; c0 += 64) {
      if (tab.n >= 64) return KS_EUNSUPPORTED;
      tab.view[tab.n] = (short)v; tab.c0[tab.n] = (short)c0; ++tab.n;
    }
  return KS_OK;
}

int conv2d_simt(int dtype, int N, int H, int W, int ksize, const ViewList &srcs, const void *weight, const float *bias,
                const ViewList &dsts, int acc_mask, double *stats, cudaStream_t st) {
  const long long M = (long long)N * H * W;
  const int Cout = dsts.cstart[dsts.n];
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((Cout + BN - 1) / BN));
  if (dtype == KS_F32) conv_simt_kernel<float><<<grid, 256, 0, st>>>(srcs, dsts, acc_mask, N, H, W, ksize, (const float *)weight, bias, stats);
  else if (dtype == KS_BF16) conv_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(srcs, dsts, acc_mask, N, H, W, ksize, (const __nv_bfloat16 *)weight, bias, stats);
  else return KS_EINVAL;
  return (int)cudaGetLastError();
}

int wgrad_simt(int dtype, int N, int H, int W, int ksize, const ViewList &xs, const ViewList &dys, float *dw, cudaStream_t st) {
  BlockTab xt, yt;
  int rc = make_block_tab(xs, xt); if (rc) return rc;
  rc = make_block_tab(dys, yt); if (rc) return rc;
  const int taps = ksize * ksize;
  const long long M = (long long)N * H * W;
  long long tiles = (long long)xt.n * yt.n * taps;
  long long splits = (kNumSMs * 6 + tiles - 1) / tiles;
  const long long maxs = (M + 255) / 256;
  if (splits > maxs) splits = maxs;
  if (splits < 1) splits = 1;
  if (splits * taps > 65535) splits = 65535 / taps;
  dim3 grid(xt.n, yt.n, (unsigned)(taps * splits));
  if (dtype == KS_F32) wgrad_simt_kernel<float><<<grid, 256, 0, st>>>(xs, dys, xt, yt, N, H, W, ksize, (int)splits, dw);
  else if (dtype == KS_BF16) wgrad_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(xs, dys, xt, yt, N, H, W, ksize, (int)splits, dw);
  else return KS_EINVAL;
  return (int)cudaGetLastError();
}

}  // namespace ks
