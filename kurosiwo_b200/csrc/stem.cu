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

// =====================================================================================================================
// Tensor-core variants for the headline input (bf16 storage, Cin == 2: the VV / VH planes).  The CUDA-core kernels above put the
// output channel on the lane and walk 32 pixels per thread: ~30 issue slots per pixel and warp, i.e. instruction-bound at ~1 TB/s
// of a 6.4 TB/s memory system (ncu, round 2: stem_fwd 0.175 ms, stem_wgrad 0.228 ms per date against a 0.034 ms HBM floor).  Here
// the 3x3x2 window is the K dimension of warp-level MMAs (K = tap * 2 + ci = 18, padded to 32 = two k16 steps):
//   forward : D[16 px][8 co] += A[16 px][K] B[K][8 co];  the shared halo tile stores each pixel as ONE 32-bit word (ci0 | ci1 << 16),
//             so an A-fragment register (k = 2t, 2t+1 = both channels of tap t) is a single LDS.32; the weights are B fragments held
//             in 16 registers for the CTA's lifetime; BatchNorm Sum x / Sum x^2 of the stored values are accumulated per thread
//             over all tiles and reduced once at CTA exit (no separate bn_stats pass over the 205 MB output).
//   wgrad   : D[16 co][8 k] += dY^T[16 co][16 px] Xcol[16 px][8 k];  dY^T fragments come from ldmatrix.trans on a per-warp copy of the
//             NHWC dY row, Xcol fragments from the same halo tile (two LDS.32 + PRMT); 24 accumulators per thread for the CTA's lifetime.
// =====================================================================================================================
constexpr int SX_P = 36;    // words per halo row (34 used)
constexpr int SO_P = 20;    // words per staged pixel (16 used): 80-byte pitch keeps ldmatrix rows / 16-byte accesses conflict-free

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t *>(&h2);
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// Halo tile of the NCHW fp32 input as packed bf16 pairs.  Every thread owns up to two fixed tile elements, so the NEXT tile's
// global loads are issued (into registers) before the current tile is computed and land in shared memory at the top of the next
// iteration: the load latency overlaps the MMAs instead of sitting between two barriers.
struct XFetch {
  int r[2], c[2];
  bool own[2];
  float v[2][2];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int i = threadIdx.x + e * 256;
      own[e] = i < (ST_TH + 2) * (ST_TW + 2);
      r[e] = i / (ST_TW + 2); c[e] = i % (ST_TW + 2);
    }
  }
  __device__ __forceinline__ void fetch(const float *__restrict__ x, long long tile, int tiles_h, int tiles_w, int H, int W) {
    const int tw = (int)(tile % tiles_w); const long long q = tile / tiles_w;
    const int th = (int)(q % tiles_h), n = (int)(q / tiles_h);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int h = th * ST_TH + r[e] - 1, w = tw * ST_TW + c[e] - 1;
      v[e][0] = 0.f; v[e][1] = 0.f;
      if (own[e] && h >= 0 && h < H && w >= 0 && w < W) {
        const float *px = x + (((long long)n * 2) * H + h) * W + w;
        v[e][0] = __ldg(px); v[e][1] = __ldg(px + (long long)H * W);
      }
    }
  }
  __device__ __forceinline__ void stash(uint32_t (*xt)[SX_P]) const {
#pragma unroll
    for (int e = 0; e < 2; ++e)
      if (own[e]) xt[r[e]][c[e]] = pack_bf16x2(v[e][0], v[e][1]);
  }
};

__global__ void __launch_bounds__(256, 3)
stem_fwd_mma_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias, View dst,
                    int N, int H, int W, int tiles_h, int tiles_w, double *stats) {
  __shared__ uint32_t xt[ST_TH + 2][SX_P];
  __shared__ __align__(16) uint32_t ost[8][16][SO_P];     // per-warp output staging: 16 pixels x 32 channels bf16
  __shared__ float sst[2][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  uint32_t bw[2][4][2];
  float bs[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int co = nt * 8 + g;
    const float *wc = w + (long long)co * 18;              // [ci][tap]
    bw[0][nt][0] = pack_bf16x2(__ldg(wc + t), __ldg(wc + 9 + t));                    // k = 2t, 2t+1  <-> tap t, ci 0 / 1
    bw[0][nt][1] = pack_bf16x2(__ldg(wc + t + 4), __ldg(wc + 9 + t + 4));            // k + 8         <-> tap t + 4
    bw[1][nt][0] = (t == 0) ? pack_bf16x2(__ldg(wc + 8), __ldg(wc + 17)) : 0u;        // k = 16, 17    <-> tap 8
    bw[1][nt][1] = 0u;
    bs[nt][0] = bias ? __ldg(bias + nt * 8 + 2 * t) : 0.f;
    bs[nt][1] = bias ? __ldg(bias + nt * 8 + 2 * t + 1) : 0.f;
  }
  float s1[4][2], s2[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) { s1[nt][0] = s1[nt][1] = s2[nt][0] = s2[nt][1] = 0.f; }
  if (threadIdx.x < 64) sst[threadIdx.x >> 5][threadIdx.x & 31] = 0.f;
  const int dy0 = t / 3, dx0 = t % 3, dy1 = (t + 4) / 3, dx1 = (t + 4) % 3;
  const long long total = (long long)N * tiles_h * tiles_w;
  XFetch xf; xf.init();
  if ((long long)blockIdx.x < total) xf.fetch(x, blockIdx.x, tiles_h, tiles_w, H, W);
  for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int tw = (int)(tile % tiles_w); long long r = tile / tiles_w;
    const int th = (int)(r % tiles_h); const int n = (int)(r / tiles_h);
    const int h0 = th * ST_TH, w0 = tw * ST_TW;
    __syncthreads();
    xf.stash(xt);
    __syncthreads();
    if (tile + gridDim.x < total) xf.fetch(x, tile + gridDim.x, tiles_h, tiles_w, H, W);
    const int h = h0 + warp;
    if (h < H) {
      __nv_bfloat16 *orow = reinterpret_cast<__nv_bfloat16 *>(dst.ptr) + ((long long)n * dst.sn + (long long)h * dst.sh);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int c0 = half * 16;
        if (w0 + c0 < W) {                                   // warp-uniform
          uint32_t a0[4], a1[4];
          a0[0] = xt[warp + dy0][c0 + g + dx0];      a0[1] = xt[warp + dy0][c0 + g + 8 + dx0];
          a0[2] = xt[warp + dy1][c0 + g + dx1];      a0[3] = xt[warp + dy1][c0 + g + 8 + dx1];
          a1[0] = (t == 0) ? xt[warp + 2][c0 + g + 2] : 0u;
          a1[1] = (t == 0) ? xt[warp + 2][c0 + g + 10] : 0u;
          a1[2] = 0u; a1[3] = 0u;
          const bool ok0 = (w0 + c0 + g) < W, ok1 = (w0 + c0 + g + 8) < W;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            float d[4] = {bs[nt][0], bs[nt][1], bs[nt][0], bs[nt][1]};
            mma_16816(d, a0, bw[0][nt][0], bw[0][nt][1]);
            mma_16816(d, a1, bw[1][nt][0], bw[1][nt][1]);
            const __nv_bfloat162 lo = __floats2bfloat162_rn(d[0], d[1]), hi = __floats2bfloat162_rn(d[2], d[3]);
            const float2 ql = __bfloat1622float2(lo), qh = __bfloat1622float2(hi);     // statistics of the STORED values
            if (ok0) { s1[nt][0] += ql.x; s1[nt][1] += ql.y; s2[nt][0] = fmaf(ql.x, ql.x, s2[nt][0]); s2[nt][1] = fmaf(ql.y, ql.y, s2[nt][1]); }
            if (ok1) { s1[nt][0] += qh.x; s1[nt][1] += qh.y; s2[nt][0] = fmaf(qh.x, qh.x, s2[nt][0]); s2[nt][1] = fmaf(qh.y, qh.y, s2[nt][1]); }
            ost[warp][g][nt * 4 + t] = *reinterpret_cast<const uint32_t *>(&lo);
            ost[warp][g + 8][nt * 4 + t] = *reinterpret_cast<const uint32_t *>(&hi);
          }
          __syncwarp();
#pragma unroll
          for (int r2 = 0; r2 < 2; ++r2) {                   // 16-byte stores: a warp writes 8 whole 64-byte pixels per instruction
            const int px = (lane >> 2) + 8 * r2, wc = w0 + c0 + px;
            if (wc < W) {
              const uint4 v = *reinterpret_cast<const uint4 *>(&ost[warp][px][(lane & 3) * 4]);
              *reinterpret_cast<uint4 *>(orow + (long long)wc * dst.sw + (lane & 3) * 8) = v;
            }
          }
          __syncwarp();
        }
      }
    }
  }
  if (stats != nullptr) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float a = s1[nt][j], b = s2[nt][j];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
        if (g == 0) { atomicAdd(&sst[0][nt * 8 + 2 * t + j], a); atomicAdd(&sst[1][nt * 8 + 2 * t + j], b); }
      }
    __syncthreads();
    if (threadIdx.x < 64) atomicAdd(stats + threadIdx.x, (double)sst[threadIdx.x >> 5][threadIdx.x & 31]);     // [2][32]
  }
}

__global__ void __launch_bounds__(256, 3)
stem_wgrad_mma_kernel(const float *__restrict__ x, View dy, float *dw, int N, int H, int W, int tiles_h, int tiles_w) {
  __shared__ uint32_t xt[ST_TH + 2][SX_P];
  __shared__ __align__(16) uint32_t dyt[2][8][ST_TW][SO_P];   // double-buffered, per warp: one image row of dY, 32 pixels x 32 channels bf16
  __shared__ float sdw[32 * 18];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  float acc[2][3][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
  for (int i = threadIdx.x; i < 32 * 18; i += blockDim.x) sdw[i] = 0.f;
  // im2col column of this thread's B fragments: k = nt * 8 + g  (tap = k >> 1, ci = k & 1; k >= 18 is padding)
  int koff[3]; uint32_t ksel[3]; bool kok[3];
#pragma unroll
  for (int nt = 0; nt < 3; ++nt) {
    const int k = nt * 8 + g, tap = k >> 1;
    kok[nt] = k < 18;
    koff[nt] = kok[nt] ? (tap / 3) * SX_P + tap % 3 : 0;
    ksel[nt] = (k & 1) ? 0x7632u : 0x5410u;                // PRMT: the ci half of two neighbouring pixel words
  }
  const long long total = (long long)N * tiles_h * tiles_w;
  // dY row of this warp for a tile -> buffer `buf` (cp.async, 16 bytes per request, zero fill outside the image)
  auto fetch_dy = [&](long long tile, int buf) {
    const int tw = (int)(tile % tiles_w); const long long q = tile / tiles_w;
    const int th = (int)(q % tiles_h), n = (int)(q / tiles_h);
    const int h = th * ST_TH + warp, w0 = tw * ST_TW;
    const __nv_bfloat16 *grow = reinterpret_cast<const __nv_bfloat16 *>(dy.ptr) + ((long long)n * dy.sn + (long long)min(h, H - 1) * dy.sh);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = lane + 32 * i, px = idx >> 2, part = idx & 3, wc = w0 + px;
      const bool ok = (h < H) && (wc < W);
      const __nv_bfloat16 *src = grow + (long long)(ok ? wc : 0) * dy.sw + part * 8;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&dyt[buf][warp][px][part * 4]);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16u : 0u) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  XFetch xf; xf.init();
  if ((long long)blockIdx.x < total) { xf.fetch(x, blockIdx.x, tiles_h, tiles_w, H, W); fetch_dy(blockIdx.x, 0); }
  int buf = 0;
  for (long long tile = blockIdx.x; tile < total; tile += gridDim.x, buf ^= 1) {
    const int tw = (int)(tile % tiles_w); long long r = tile / tiles_w;
    const int th = (int)(r % tiles_h);
    const int h0 = th * ST_TH, w0 = tw * ST_TW;
    __syncthreads();                       // the previous tile's fragments have been read
    xf.stash(xt);
    const bool more = tile + gridDim.x < total;
    if (more) { xf.fetch(x, tile + gridDim.x, tiles_h, tiles_w, H, W); fetch_dy(tile + gridDim.x, buf ^ 1); }
    if (more) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const int h = h0 + warp;
    const uint32_t dyt_u = (uint32_t)__cvta_generic_to_shared(&dyt[buf][warp][0][0]);
    if (h < H) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int c0 = half * 16;
        if (w0 + c0 < W) {
          uint32_t a[2][4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const int id = lane >> 3, rr = lane & 7;
            const uint32_t addr = dyt_u + (uint32_t)(((c0 + (id >> 1) * 8 + rr) * SO_P + (mt * 16 + (id & 1) * 8) / 2) * 4);
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                         : "=r"(a[mt][0]), "=r"(a[mt][1]), "=r"(a[mt][2]), "=r"(a[mt][3]) : "r"(addr));
          }
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) {
            uint32_t b0 = 0u, b1 = 0u;
            if (kok[nt]) {
              const uint32_t *xr = &xt[warp][c0 + 2 * t] + koff[nt];
              b0 = __byte_perm(xr[0], xr[1], ksel[nt]);          // pixels 2t, 2t+1
              b1 = __byte_perm(xr[8], xr[9], ksel[nt]);          // pixels 2t+8, 2t+9
            }
            mma_16816(acc[0][nt], a[0], b0, b1);
            mma_16816(acc[1][nt], a[1], b0, b1);
          }
        }
      }
    }
  }
  // D fragment: rows co = mt*16 + g (+8), columns k = nt*8 + 2t (+1)
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int co = mt * 16 + g + ((i >> 1) ? 8 : 0), k = nt * 8 + 2 * t + (i & 1);
        if (k < 18) atomicAdd(&sdw[co * 18 + k], acc[mt][nt][i]);
      }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 18; i += blockDim.x) {
    const int co = i / 18, k = i % 18, tap = k >> 1, ci = k & 1;
    atomicAdd(dw + ((long long)co * 2 + ci) * 9 + tap, sdw[i]);
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

// 16-byte accesses on whole 64-byte pixels
static bool mma_view_ok(const ks_view_t &v) {
  return (((uintptr_t)v.ptr) % 16 == 0) && (v.sn % 8 == 0) && (v.sh % 8 == 0) && (v.sw % 8 == 0);
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
  if (dtype == KS_BF16 && Cin == 2 && !g_opt.stem_simt && mma_view_ok(*dst)) {
    const long long res = (long long)kNumSMs * 3;          // exactly the resident CTAs (80 registers x 256 threads: 3 per SM)
    grid = dim3((unsigned)(total < res ? total : res));
    stem_fwd_mma_kernel<<<grid, 256, 0, st>>>(x_nchw, w_oihw, bias, to_view(*dst), N, H, W, th, tw, stats);     // statistics fused
    return (int)cudaGetLastError();
  }
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
  if (dtype == KS_BF16 && Cin == 2 && !g_opt.stem_simt && mma_view_ok(*dy)) {
    const long long res = (long long)kNumSMs * 3;
    grid = dim3((unsigned)(total < res ? total : res));
    stem_wgrad_mma_kernel<<<grid, 256, 0, st>>>(x_nchw, to_view(*dy), dw_oihw, N, H, W, th, tw);
    return (int)cudaGetLastError();
  }
  if (dtype == KS_F32) return stem_wgrad_dispatch<float>(Cin, grid, st, x_nchw, to_view(*dy), dw_oihw, N, H, W, th, tw);
  if (dtype == KS_BF16) return stem_wgrad_dispatch<__nv_bfloat16>(Cin, grid, st, x_nchw, to_view(*dy), dw_oihw, N, H, W, th, tw);
  return KS_EINVAL;
}
