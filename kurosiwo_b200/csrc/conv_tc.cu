// TMA + tcgen05 implicit-GEMM convolution for sm_100a (bf16 NHWC, fp32 accumulate in TMEM).
//
//   out[pixel, co] = bias[co] + sum_{tap, src, ci} x_src[pixel + tap, ci] * w[tap][co][ci]
//
// Mapping: M = 128 output pixels (an 8-row x 16-column spatial window, of which 14 columns are
// real outputs for 3x3), N = Cout tile (<=256), K = (source, 64- or 32-channel chunk, tap).
//
// Halo reuse: for each K chunk ONE TMA box of (8+2) x 16 pixels x BK channels lands in shared
// memory (zero-filled outside the image = the conv padding).  Because the box pitch is exactly
// 16 pixels, the operand of tap (r,s) is the SAME smem tile viewed from row offset r*16+s: 128
// consecutive 128-byte rows.  So the nine taps are nine UMMA descriptors over one tile; L2->SMEM
// activation traffic is 1.43x the input instead of 9x (im2col / per-tap loads).
// Weights stream through a second mbarrier ring, one [BN x BK] K-major tile per (chunk, tap).
// MT output windows share every weight tile (MT accumulators of BN TMEM columns each).
//
// Warp roles (192 threads): warp0 = TMA producer, warp1 = TMEM alloc + MMA issuer,
// warps2-5 = epilogue (TMEM -> registers -> +bias (+=dst) -> bf16 -> global).
//
// Replaces: cuDNN fprop/dgrad behind nn.Conv2d / nn.ConvTranspose2d (models/snunet.py:15,17,41)
// and the torch.cat copies of models/snunet.py:132-144 (K loop walks the source views).
#include "common.cuh"
#include "tc_common.cuh"

namespace ks {

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

Options g_opt = {};

struct alignas(64) ConvTcParams {
  CUtensorMap src[KS_MAX_VIEWS];
  CUtensorMap wmap;
  ViewList dsts;
  const float *bias;
  int cstart[KS_MAX_VIEWS + 1];
  int n_src, acc_mask;
  int N, H, W, tiles_w, tiles_h, total_tiles;
  int BN, MT, SA, SB, TH, bo_mode;
  uint32_t a_tile_bytes, b_stage_bytes, tmem_cols, idesc;
  // v2 (persistent) fields
  int n_super, resident, n_acc, n_wtiles, TB;
  uint32_t b_tile_bytes;
  double *stats;
  int Cout, debug;
  int BNA;   // TMEM columns per accumulator: BN, or 3*BN when the three column taps are stacked along N
};

using namespace tc;

template <int BK, int KS>
__global__ void __launch_bounds__(192, 1) conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int TW = (KS == 3) ? 14 : 16;
  constexpr int PADK = KS / 2;
  constexpr int BOX_ROWS = (KS == 3) ? 10 : 8;
  constexpr uint32_t ROW_BYTES = BK * 2;
  constexpr uint32_t A_BOX_BYTES = BOX_ROWS * 16 * ROW_BYTES;
  constexpr uint32_t LAYOUT = (BK == 64) ? LAYOUT_SW128 : LAYOUT_SW64;
  constexpr uint32_t SBO = 8 * ROW_BYTES;
  constexpr int TAPS = KS * KS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int SA = p.SA, SB = p.SB, MT = p.MT, BN = p.BN;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *sm = smem_raw + (base - raw);
  const uint32_t a_base = base;
  const uint32_t b_base = a_base + (uint32_t)SA * MT * p.a_tile_bytes;
  const uint32_t bar_base = b_base + (uint32_t)SB * p.b_stage_bytes;
  auto a_full = [&](int i) { return bar_base + 8u * i; };
  auto a_empty = [&](int i) { return bar_base + 8u * (SA + i); };
  auto b_full = [&](int i) { return bar_base + 8u * (2 * SA + i); };
  auto b_empty = [&](int i) { return bar_base + 8u * (2 * SA + SB + i); };
  const uint32_t acc_full = bar_base + 8u * (2 * SA + 2 * SB);
  const uint32_t tmem_slot = acc_full + 8u;
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(sm + (tmem_slot - base));

  // this CTA's output windows
  const int t0 = blockIdx.x * MT;
  const int nvalid = min(MT, p.total_tiles - t0);
  const int n0 = blockIdx.y * BN;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.n_src; ++s) prefetch_tmap(&p.src[s]);
    prefetch_tmap(&p.wmap);
    for (int i = 0; i < SA; ++i) { mbar_init(a_full(i), 1); mbar_init(a_empty(i), 1); }
    for (int i = 0; i < SB; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int sa = 0, pa = 0, sb = 0, pb = 0;
      for (int s = 0; s < p.n_src; ++s) {
        const int C = p.cstart[s + 1] - p.cstart[s];
        for (int c0 = 0; c0 < C; c0 += BK) {
          mbar_wait(a_empty(sa), pa ^ 1);
          mbar_expect_tx(a_full(sa), (uint32_t)nvalid * A_BOX_BYTES);
          for (int mt = 0; mt < nvalid; ++mt) {
            const int t = t0 + mt;
            const int tw = t % p.tiles_w, th = (t / p.tiles_w) % p.tiles_h, n = t / (p.tiles_w * p.tiles_h);
            tma_load_4d(a_base + (uint32_t)(sa * MT + mt) * p.a_tile_bytes, &p.src[s], c0, tw * TW - PADK, th * p.TH - PADK, n, a_full(sa));
          }
          if (++sa == SA) { sa = 0; pa ^= 1; }
          for (int tap = 0; tap < TAPS; ++tap) {
            mbar_wait(b_empty(sb), pb ^ 1);
            mbar_expect_tx(b_full(sb), p.b_stage_bytes);
            tma_load_3d(b_base + (uint32_t)sb * p.b_stage_bytes, &p.wmap, p.cstart[s] + c0, n0, tap, b_full(sb));
            if (++sb == SB) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      int sa = 0, pa = 0, sb = 0, pb = 0;
      bool first = true;
      for (int s = 0; s < p.n_src; ++s) {
        const int C = p.cstart[s + 1] - p.cstart[s];
        for (int c0 = 0; c0 < C; c0 += BK) {
          mbar_wait(a_full(sa), pa);
          for (int tap = 0; tap < TAPS; ++tap) {
            mbar_wait(b_full(sb), pb);
            tc_fence_after();
            const uint32_t b_addr = b_base + (uint32_t)sb * p.b_stage_bytes;
            const uint32_t row_off = (KS == 3) ? (uint32_t)((tap / 3) * 16 + (tap % 3)) : 0u;
            for (int mt = 0; mt < nvalid; ++mt) {
              const uint32_t a_addr = a_base + (uint32_t)(sa * MT + mt) * p.a_tile_bytes + row_off * ROW_BYTES;
              const uint32_t bo = p.bo_mode ? ((a_addr >> 7) & 7u) : 0u;
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                const uint64_t ad = make_smem_desc(a_addr + k * 32, 16, SBO, LAYOUT, bo);
                const uint64_t bd = make_smem_desc(b_addr + k * 32, 16, SBO, LAYOUT, 0);
                umma_bf16(tmem_base + (uint32_t)(mt * BN), ad, bd, p.idesc, (first && k == 0) ? 0u : 1u);
              }
            }
            first = false;
            tc_commit(b_empty(sb));
            if (++sb == SB) { sb = 0; pb ^= 1; }
          }
          tc_commit(a_empty(sa));
          if (++sa == SA) { sa = 0; pa ^= 1; }
        }
      }
      tc_commit(acc_full);
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane group = warp % 4 =====
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    const int ty = row >> 4, tx = row & 15;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    for (int mt = 0; mt < nvalid; ++mt) {
      const int t = t0 + mt;
      const int tw = t % p.tiles_w, th = (t / p.tiles_w) % p.tiles_h, n = t / (p.tiles_w * p.tiles_h);
      const int h = th * p.TH + ty, w = tw * TW + tx;
      const bool ok = (tx < TW) && (ty < p.TH) && (h < p.H) && (w < p.W);
      for (int cc = 0; cc < BN; cc += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(mt * BN + cc), r);
        tmem_ld_wait();
        if (ok) {
          const int co = n0 + cc;
          int d = 0;
#pragma unroll
          for (int i = 1; i < KS_MAX_VIEWS; ++i) if (i < p.dsts.n && co >= p.dsts.cstart[i]) d = i;
          const View &dv = p.dsts.v[d];
          __nv_bfloat16 *op = reinterpret_cast<__nv_bfloat16 *>(dv.ptr) +
                              ((long long)n * dv.sn + (long long)h * dv.sh + (long long)w * dv.sw + (co - p.dsts.cstart[d]));
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + (p.bias ? __ldg(p.bias + co + i) : 0.f);
          if ((p.acc_mask >> d) & 1) {
            float o[8];
            ld8(op, o);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += o[i];
            ld8(op + 8, o);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[8 + i] += o[i];
          }
          float a[8], b[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { a[i] = v[i]; b[i] = v[8 + i]; }
          st8(op, a);
          st8(op + 8, b);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, p.tmem_cols); }
}


// ------------------------------------------------------------------------------------------------
// v2: persistent CTAs, optional smem-RESIDENT weights, double-buffered TMEM accumulators (the epilogue
// of super-tile i overlaps the MMAs of super-tile i+1) and BatchNorm statistics fused into the epilogue.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_reduce_scatter16(float (&v)[16], int lane) {
  // after the call v[0] holds the 32-lane sum of channel ((lane>>1)&15) (both lanes of a pair hold it)
  bool hi = lane & 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float send = hi ? v[i] : v[i + 8], keep = hi ? v[i + 8] : v[i]; v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16); }
  hi = lane & 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float send = hi ? v[i] : v[i + 4], keep = hi ? v[i + 4] : v[i]; v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8); }
  hi = lane & 4;
#pragma unroll
  for (int i = 0; i < 2; ++i) { const float send = hi ? v[i] : v[i + 2], keep = hi ? v[i + 2] : v[i]; v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4); }
  hi = lane & 2;
  { const float send = hi ? v[0] : v[1], keep = hi ? v[1] : v[0]; v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2); }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// Issue the MMAs of NT consecutive taps for up to 4 output windows.  Every descriptor is (per-stage base + compile-time
// immediate): no loop-carried dependency, so the single issuing lane is throughput- not latency-bound (the uniform
// datapath has ~10-cycle ALU latency; chained descriptor arithmetic cost ~55 cycles per MMA before this).
template <int NT, int BK, bool NS3>
__device__ __forceinline__ void issue_taps(uint32_t a_lo_row, uint32_t a_tile_step, uint32_t b_lo0, uint32_t b_tile_step,
                                           uint32_t desc_hi, uint32_t acc_col, int BNA, int nvalid, uint32_t idesc,
                                           uint32_t accum_or, bool skip) {
  constexpr uint32_t ROW16 = (BK * 2) >> 4;
  if constexpr (NS3) {
    // Column taps stacked along N: ONE UMMA per (kernel row, k step) multiplies the halo tile viewed from row offset
    // r*16 (column offset 0) with the [3*BN x BK] tile {W[r][0]; W[r][1]; W[r][2]} (three consecutive weight tiles), so
    // accumulator column block s holds sum_r x[q + 16 r] W[r][s]; the epilogue adds the blocks with a lane shift of s.
    // A 128 x 3BN x 16 UMMA reads 4 KB + 3 BN * 32 B of shared memory per 1.5 BN tensor cycles instead of per BN / 2.
    static_assert(NT % 3 == 0, "row groups");
#pragma unroll
    for (int rr = 0; rr < NT / 3; ++rr) {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        if (mt < nvalid) {
          const uint32_t a_lo = a_lo_row + (uint32_t)mt * a_tile_step + (uint32_t)(rr * 16) * ROW16;
          const uint32_t b_lo = b_lo0 + (uint32_t)(3 * rr) * b_tile_step;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + 2u * k);
            const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2u * k);
            if (!skip) umma_bf16(acc_col + (uint32_t)(mt * BNA), ad, bd, idesc, accum_or | (uint32_t)(rr | k));
          }
        }
      }
    }
  } else {
#pragma unroll
    for (int tl = 0; tl < NT; ++tl) {
      const uint32_t roff = (NT == 9) ? (uint32_t)((tl / 3) * 16 + tl % 3) : (uint32_t)tl;
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        if (mt < nvalid) {
          const uint32_t a_lo = a_lo_row + (uint32_t)mt * a_tile_step + roff * ROW16;
          const uint32_t b_lo = b_lo0 + (uint32_t)tl * b_tile_step;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + 2u * k);
            const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2u * k);
            if (!skip) umma_bf16(acc_col + (uint32_t)(mt * BNA), ad, bd, idesc, accum_or | (uint32_t)(tl | k));
          }
        }
      }
    }
  }
}

// EW = epilogue warps: 8 (320 threads, up to 2 CTAs per SM) or 16 (576 threads, ONE CTA per SM: configurations whose weights or
// TMEM columns already pin the CTA count to one and whose epilogue - 4 sequential 32-column items per warp - paced the kernel).
// STATR: BatchNorm statistics in per-thread registers over the CTA's tiles (2 x 32 floats; one CTA per SM), reduced once at CTA exit.
template <int BK, int KS, bool NS3, int EW, bool STATR>
__global__ void __launch_bounds__(64 + 32 * EW, (EW == 8 && !STATR) ? 2 : 1) conv_tc2_kernel(const __grid_constant__ ConvTcParams p) {
#define MBW(bar, par) do { if (p.debug & 32) mbar_wait_spin(bar, par); else mbar_wait(bar, par); } while (0)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int TW = (KS == 3) ? 14 : 16;
  constexpr int PADK = KS / 2;
  constexpr int BOX_ROWS = (KS == 3) ? 10 : 8;
  constexpr uint32_t ROW_BYTES = BK * 2;
  constexpr uint32_t A_BOX_BYTES = BOX_ROWS * 16 * ROW_BYTES;
  constexpr uint32_t LAYOUT = (BK == 64) ? LAYOUT_SW128 : LAYOUT_SW64;
  constexpr uint32_t SBO = 8 * ROW_BYTES;
  constexpr int TAPS = KS * KS;

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const int SA = p.SA, SB = p.SB, MT = p.MT, BN = p.BN, NACC = p.n_acc, BNA = p.BNA;
  const bool RES = p.resident != 0;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *sm = smem_raw + (base - raw);
  const uint32_t a_base = base;
  const uint32_t b_base = a_base + (uint32_t)SA * MT * p.a_tile_bytes;
  const int TB = p.TB;
  const uint32_t bar_base = b_base + (RES ? (uint32_t)p.n_wtiles * p.b_tile_bytes : (uint32_t)SB * p.b_stage_bytes);
  auto a_full = [&](int i) { return bar_base + 8u * i; };
  auto a_empty = [&](int i) { return bar_base + 8u * (SA + i); };
  auto b_full = [&](int i) { return bar_base + 8u * (2 * SA + i); };
  auto b_empty = [&](int i) { return bar_base + 8u * (2 * SA + SB + i); };
  const uint32_t misc = bar_base + 8u * (2 * SA + 2 * SB);
  auto acc_full = [&](int i) { return misc + 8u * i; };
  auto acc_empty = [&](int i) { return misc + 8u * (2 + i); };
  const uint32_t w_full = misc + 8u * 4;
  const uint32_t tmem_slot = misc + 8u * 5;
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(sm + (tmem_slot - base));
  float *sstat = reinterpret_cast<float *>(sm + (tmem_slot + 8u - base));   // [2][BN] BatchNorm partial sums
  float *sbias = sstat + 2 * BN;                                            // [BN] bias of this N tile (0 if none)
  int2 *sdst = reinterpret_cast<int2 *>(sbias + BN);                        // [BN/16] {dst view, channel offset in view}

  const int n0 = blockIdx.y * BN;

  if (warp == 0 && elect_one()) {
    for (int s = 0; s < p.n_src; ++s) prefetch_tmap(&p.src[s]);
    prefetch_tmap(&p.wmap);
    for (int i = 0; i < SA; ++i) { mbar_init(a_full(i), 1); mbar_init(a_empty(i), 1); }
    for (int i = 0; i < SB; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), EW); }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) sstat[i] = 0.f;
  for (int i = threadIdx.x; i < BN; i += blockDim.x) sbias[i] = p.bias ? __ldg(p.bias + n0 + i) : 0.f;
  for (int i = threadIdx.x; i < BN / 16; i += blockDim.x) {
    const int co = n0 + 16 * i;
    int d = 0;
    for (int q = 1; q < KS_MAX_VIEWS; ++q) if (q < p.dsts.n && co >= p.dsts.cstart[q]) d = q;
    sdst[i] = make_int2(d, co - p.dsts.cstart[d]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===== TMA producer (whole warp runs the loops; one elected lane issues) =====
    if (RES) {
      if (elect_one()) {
        mbar_expect_tx(w_full, (uint32_t)p.n_wtiles * p.b_tile_bytes);
        int idx = 0;
        for (int s = 0; s < p.n_src; ++s) {
          const int C = p.cstart[s + 1] - p.cstart[s];
          for (int c0 = 0; c0 < C; c0 += BK)
            for (int tap = 0; tap < TAPS; ++tap, ++idx)
              tma_load_3d(b_base + (uint32_t)idx * p.b_tile_bytes, &p.wmap, p.cstart[s] + c0, n0, tap, w_full);
        }
      }
      __syncwarp();
    }
    int sa = 0, pa = 0, sb = 0, pb = 0;
    for (int st = blockIdx.x; st < p.n_super; st += gridDim.x) {
      const int t0 = st * MT;
      const int nvalid = min(MT, p.total_tiles - t0);
      for (int s = 0; s < p.n_src; ++s) {
        const int C = p.cstart[s + 1] - p.cstart[s];
        for (int c0 = 0; c0 < C; c0 += BK) {
          MBW(a_empty(sa), pa ^ 1);
          if (elect_one()) {
            if (p.debug & 16) mbar_arrive(a_full(sa));
            else mbar_expect_tx(a_full(sa), (uint32_t)nvalid * A_BOX_BYTES);
            for (int mt = 0; mt < ((p.debug & 16) ? 0 : nvalid); ++mt) {
              const int t = t0 + mt;
              const int tw = t % p.tiles_w, th = (t / p.tiles_w) % p.tiles_h, n = t / (p.tiles_w * p.tiles_h);
              tma_load_4d(a_base + (uint32_t)(sa * MT + mt) * p.a_tile_bytes, &p.src[s], c0, tw * TW - PADK, th * p.TH - PADK, n, a_full(sa));
            }
          }
          __syncwarp();
          if (++sa == SA) { sa = 0; pa ^= 1; }
          if (!RES) {
            for (int tg = 0; tg < TAPS / TB; ++tg) {       // one weight stage = TB taps (a kernel row when TB == 3)
              MBW(b_empty(sb), pb ^ 1);
              if (elect_one()) {
                mbar_expect_tx(b_full(sb), p.b_stage_bytes);
                for (int tl = 0; tl < TB; ++tl)
                  tma_load_3d(b_base + (uint32_t)sb * p.b_stage_bytes + (uint32_t)tl * p.b_tile_bytes, &p.wmap, p.cstart[s] + c0, n0,
                              tg * TB + tl, b_full(sb));
              }
              __syncwarp();
              if (++sb == SB) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp waits; one elected lane issues tcgen05.mma / commit) =====
    if (RES) MBW(w_full, 0);
    // descriptor constants: hi word = SBO | version | layout, lo word = (addr >> 4) | LBO(=1) << 16
    const uint32_t desc_hi = (uint32_t)((SBO >> 4) & 0x3FFFu) | (1u << 14) | (LAYOUT << 29);
    const uint32_t a_tile_step = p.a_tile_bytes >> 4, b_tile_step = p.b_tile_bytes >> 4;
    const bool skip = (p.debug & 4) != 0;
    int sa = 0, pa = 0, sb = 0, pb = 0, as = 0, pacc = 0;
    for (int st = blockIdx.x; st < p.n_super; st += gridDim.x) {
      const int nvalid = min(MT, p.total_tiles - st * MT);
      MBW(acc_empty(as), pacc ^ 1);
      tc_fence_after();
      const uint32_t acc_col = tmem_base + (uint32_t)(as * MT * BNA);
      uint32_t accum = 0, widx = 0;
      for (int s = 0; s < p.n_src; ++s) {
        const int C = p.cstart[s + 1] - p.cstart[s];
        for (int c0 = 0; c0 < C; c0 += BK) {
          MBW(a_full(sa), pa);
          tc_fence_after();
          const uint32_t a_lo0 = (((a_base + (uint32_t)(sa * MT) * p.a_tile_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
          if (RES) {
            if (elect_one()) {
              const uint32_t b_lo0 = (((b_base + widx * p.b_tile_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
              issue_taps<TAPS, BK, NS3>(a_lo0, a_tile_step, b_lo0, b_tile_step, desc_hi, acc_col, BNA, nvalid, p.idesc, accum, skip);
              tc_commit(a_empty(sa));
            }
            __syncwarp();
            widx += TAPS;
          } else {
            for (int tg = 0; tg < TAPS / TB; ++tg) {
              MBW(b_full(sb), pb);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t b_lo0 = (((b_base + (uint32_t)sb * p.b_stage_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
                if (KS == 3 && TB == 3) {
                  issue_taps<3, BK, NS3>(a_lo0 + (uint32_t)(tg * 16) * (ROW_BYTES >> 4), a_tile_step, b_lo0, b_tile_step, desc_hi, acc_col, BNA,
                                         nvalid, p.idesc, accum | (uint32_t)tg, skip);
                } else {
                  const uint32_t roff = (KS == 3) ? (uint32_t)((tg / 3) * 16 + tg % 3) : 0u;
                  issue_taps<1, BK, false>(a_lo0 + roff * (ROW_BYTES >> 4), a_tile_step, b_lo0, b_tile_step, desc_hi, acc_col, BNA, nvalid,
                                           p.idesc, accum | (uint32_t)tg, skip);
                }
                tc_commit(b_empty(sb));
                if (tg == TAPS / TB - 1) tc_commit(a_empty(sa));
              }
              __syncwarp();
              if (++sb == SB) { sb = 0; pb ^= 1; }
            }
          }
          accum = 1;
          if (++sa == SA) { sa = 0; pa ^= 1; }
        }
      }
      if (elect_one()) tc_commit(acc_full(as));
      __syncwarp();
      if (++as == NACC) { as = 0; pacc ^= 1; }
    }
  } else {
    // ===== epilogue: warps 2..9.  TMEM lane group = warp % 4; the two warps of a lane group split the
    // (window, 32-column chunk) work items.  Per item: one 32-column TMEM load, the optional += loads are
    // issued before the TMEM wait, then bias, bf16 pack, four 16-byte stores and the BN-statistics butterfly.
    const int lg = warp & 3;
    const int half = (warp - 2) >> 2;      // 0 .. EW/4-1: the warps of a lane group split the work items
    const int row = lg * 32 + lane;
    const int ty = row >> 4, tx = row & 15;
    const int nchunk = (BN + 31) >> 5;
    float ra[STATR ? 32 : 1], rq[STATR ? 32 : 1];
    int r_cc = 0;
    if constexpr (STATR) {
#pragma unroll
      for (int i = 0; i < 32; ++i) { ra[i] = 0.f; rq[i] = 0.f; }
    }
    int as = 0, pacc = 0;
    for (int st = blockIdx.x; st < p.n_super; st += gridDim.x) {
      const int t0 = st * MT;
      const int nvalid = min(MT, p.total_tiles - t0);
      MBW(acc_full(as), pacc);
      tc_fence_after();
      int cur_mt = -1, n = 0, h = 0, w = 0;
      bool ok = false;
      for (int item = half; item < ((p.debug & 8) ? 0 : nvalid * nchunk); item += EW / 4) {
        const int mt = item / nchunk, cc = (item % nchunk) << 5;
        if (mt != cur_mt) {            // window coordinates change once per MT accumulator, not per 32-column item
          cur_mt = mt;
          const int t = t0 + mt;
          const int tw = t % p.tiles_w, th = (t / p.tiles_w) % p.tiles_h;
          n = t / (p.tiles_w * p.tiles_h);
          h = th * p.TH + ty; w = tw * TW + tx;
          ok = (tx < TW) && (ty < p.TH) && (h < p.H) && (w < p.W);
        }
        const bool wide = (BN - cc) >= 32;
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * MT * BNA + mt * BNA + cc);
        uint32_t r[32];
        if constexpr (!NS3) {
          if (p.debug & 2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = 0u;
          } else if (wide) tmem_ld32(taddr, r);
          else { uint32_t r16[16]; tmem_ld16(taddr, r16);
#pragma unroll
            for (int i = 0; i < 16; ++i) { r[i] = r16[i]; r[16 + i] = 0u; } }
        }
        // destination pointers of the two 16-column halves (view boundaries are multiples of 16 channels)
        __nv_bfloat16 *op[2]; bool accd[2];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int2 dd = sdst[min((cc >> 4) + hh, BN / 16 - 1)];
          const View &dv = p.dsts.v[dd.x];
          op[hh] = reinterpret_cast<__nv_bfloat16 *>(dv.ptr) + ((long long)n * dv.sn + (long long)h * dv.sh + (long long)w * dv.sw + dd.y);
          accd[hh] = ((p.acc_mask >> dd.x) & 1) != 0;
        }
        const int nh = wide ? 2 : 1;
        uint32_t old[2][8];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
          if (ok && hh < nh && accd[hh]) ld_global_v8(op[hh], old[hh]);
        float v[32];
        if constexpr (NS3) {
          // out[lane] = D0[lane] + D1[lane + 1] + D2[lane + 2]: column-tap block s sits BN columns further and belongs to the
          // pixel s columns to the right, i.e. lane + s of the same 16-pixel row (real outputs have tx <= 13, so tx + s <= 15).
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            if (hh < nh) {
              uint32_t d0[16], d1[16], d2[16];
              tmem_ld16(taddr + 16 * hh, d0);
              tmem_ld16(taddr + BN + 16 * hh, d1);
              tmem_ld16(taddr + 2 * BN + 16 * hh, d2);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i)
                v[16 * hh + i] = __uint_as_float(d0[i]) + __shfl_down_sync(0xffffffffu, __uint_as_float(d1[i]), 1) +
                                 __shfl_down_sync(0xffffffffu, __uint_as_float(d2[i]), 2);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[16 * hh + i] = 0.f;
            }
          }
        } else {
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        }
        if (p.bias) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (q < 4 * nh) {
              const float4 b4 = *reinterpret_cast<const float4 *>(sbias + cc + 4 * q);
              v[4 * q] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
            }
          }
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          if (ok && hh < nh) {
            if (accd[hh]) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&old[hh][i]));
                v[16 * hh + 2 * i] += f2.x; v[16 * hh + 2 * i + 1] += f2.y;
              }
            }
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[16 * hh + 2 * i], v[16 * hh + 2 * i + 1]);
              pk[i] = *reinterpret_cast<const uint32_t *>(&h2);
            }
            if (!(p.debug & 1)) st_global_v8(op[hh], pk);
          }
        }
        if constexpr (STATR) {
          if (p.stats) {                 // the host guarantees one fixed 32-channel chunk per warp (nchunk divides EW / 4)
            r_cc = cc;
            if (ok) {
#pragma unroll
              for (int i = 0; i < 32; ++i) { const float q = round_as<__nv_bfloat16>(v[i]); ra[i] += q; rq[i] = fmaf(q, q, rq[i]); }
            }
          }
        } else
        if (p.stats && !(p.debug & 128)) {      // tc_debug bit 7: skip the statistics work (ablation only)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            if (hh < nh) {
              float s1[16], s2[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) { const float q = ok ? round_as<__nv_bfloat16>(v[16 * hh + i]) : 0.f; s1[i] = q; s2[i] = q * q; }
              warp_reduce_scatter16(s1, lane);
              warp_reduce_scatter16(s2, lane);
              if ((lane & 1) == 0) {
                const int c = cc + 16 * hh + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                atomicAdd(&sstat[c], s1[0]);
                atomicAdd(&sstat[BN + c], s2[0]);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(as));
      if (++as == NACC) { as = 0; pacc ^= 1; }
    }
    if constexpr (STATR) {
      if (p.stats) {                   // one butterfly per warp and CTA instead of one per item
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float s1[16], s2[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) { s1[i] = ra[16 * hh + i]; s2[i] = rq[16 * hh + i]; }
          warp_reduce_scatter16(s1, lane);
          warp_reduce_scatter16(s2, lane);
          if ((lane & 1) == 0) {
            const int c = r_cc + 16 * hh + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            if (c < BN) { atomicAdd(&sstat[c], s1[0]); atomicAdd(&sstat[BN + c], s2[0]); }
          }
        }
      }
    }
  }
  __syncthreads();
  if (p.stats) {
    for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) {
      const int r = i / BN, c = i % BN;
      atomicAdd(p.stats + (size_t)r * p.Cout + n0 + c, (double)sstat[i]);
    }
  }
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, p.tmem_cols); }
}

#undef MBW
static bool tma_view_ok(const View &v) {
  return (((uintptr_t)v.ptr) % 16 == 0) && ((v.sn * 2) % 16 == 0) && ((v.sh * 2) % 16 == 0) && ((v.sw * 2) % 16 == 0) &&
         v.sw > 0 && v.sh > 0 && v.sn > 0;
}

int encode_act_map(CUtensorMap *m, const View &v, int N, int H, int W, int box_c, int box_w, int box_h, CUtensorMapSwizzle sw) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return KS_EDRIVER;
  cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)v.sw * 2, (cuuint64_t)v.sh * 2, (cuuint64_t)v.sn * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void *)v.ptr, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? KS_OK : KS_EDRIVER;
}

template <int BK, int KS>
static int launch_conv_tc(const ConvTcParams &p, dim3 grid, size_t smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BK, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  conv_tc_kernel<BK, KS><<<grid, 192, smem, st>>>(p);
  return (int)cudaGetLastError();
}

extern "C" int ks_bn_stats(int dtype, int N, int H, int W, const ks_view_t *x, double *sums, void *stream);

template <int BK, int KS, bool NS3, int EW, bool STATR>
static int launch_conv_tc2(const ConvTcParams &p, dim3 grid, size_t smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<BK, KS, NS3, EW, STATR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  conv_tc2_kernel<BK, KS, NS3, EW, STATR><<<grid, 64 + 32 * EW, smem, st>>>(p);
  return (int)cudaGetLastError();
}

int conv2d_tc(int N, int H, int W, int ksize, const ViewList &srcs, const void *weight, const float *bias,
              const ViewList &dsts, int acc_mask, double *stats, cudaStream_t st) {
  if (g_opt.tc_disable) return KS_EUNSUPPORTED;
  const int Cin = srcs.cstart[srcs.n], Cout = dsts.cstart[dsts.n];
  const bool v1_opt = g_opt.v1 != 0;
  bool all64 = true;
  for (int s = 0; s < srcs.n; ++s) {
    if (srcs.v[s].C % 32) return KS_EUNSUPPORTED;
    if (srcs.v[s].C % 64) all64 = false;
    if (!tma_view_ok(srcs.v[s])) return KS_EUNSUPPORTED;
  }
  for (int d = 0; d < dsts.n; ++d) {
    if (dsts.cstart[d] % 16 || dsts.v[d].C % 16) return KS_EUNSUPPORTED;
    if (!tma_view_ok(dsts.v[d])) return KS_EUNSUPPORTED;
    if (!v1_opt && ((((uintptr_t)dsts.v[d].ptr) % 32) || (dsts.v[d].sw * 2) % 32 || (dsts.v[d].sh * 2) % 32 || (dsts.v[d].sn * 2) % 32)) return KS_EUNSUPPORTED;
  }
  if (((uintptr_t)weight) % 16) return KS_EUNSUPPORTED;
  if (stats && dsts.n != 1) return KS_EUNSUPPORTED;
  if ((Cin * 2) % 16) return KS_EUNSUPPORTED;
  const int BK = all64 ? 64 : 32;
  const int taps = ksize * ksize;
  int nt = 1;
  for (;; ++nt) {
    if (nt > 64) return KS_EUNSUPPORTED;
    if (Cout % nt == 0 && (Cout / nt) <= 256 && (Cout / nt) % 16 == 0) break;
  }
  const bool v1 = g_opt.v1 != 0;
  ConvTcParams p;
  p.BN = Cout / nt;
  p.TH = (ksize == 3) ? ((H % 8 == 0) ? 8 : ((H % 7 == 0) ? 7 : 8)) : 8;
  const int TW = (ksize == 3) ? 14 : 16;
  p.tiles_w = (W + TW - 1) / TW; p.tiles_h = (H + p.TH - 1) / p.TH;
  const long long tt = (long long)p.tiles_w * p.tiles_h * N;
  if (tt > 0x7fffffffLL) return KS_EUNSUPPORTED;
  p.total_tiles = (int)tt;
  p.N = N; p.H = H; p.W = W;
  p.n_src = srcs.n; p.acc_mask = acc_mask; p.bias = bias; p.dsts = dsts;
  for (int i = 0; i <= KS_MAX_VIEWS; ++i) p.cstart[i] = srcs.cstart[i];
  p.bo_mode = g_opt.bo_mode;
  p.Cout = Cout;
  p.debug = g_opt.debug;
  p.stats = v1 ? nullptr : stats;
  const uint32_t row_bytes = BK * 2;
  const uint32_t a_rows = (ksize == 3) ? 162 : 128;
  p.a_tile_bytes = ((a_rows * row_bytes + 1023) / 1024) * 1024;
  p.b_tile_bytes = (uint32_t)p.BN * row_bytes;
  p.b_stage_bytes = p.b_tile_bytes;   // streamed: x TB below
  p.TB = 1;
  // column taps stacked along N (see issue_taps): 3x3 convs with narrow N tiles, where a 128 x BN x 16 UMMA is bound by its
  // 4 KB shared-memory read of the pixel operand (BN <= 64: at most 40 % / 67 % of the tensor rate)
  // Measured (scripts/bench_layers.py, KS_DEBUG ablations, B200): a 128 x N x 16 UMMA costs max(N / 2, (4096 + 32 N) / 128) cycles
  // (N = 32: 40, N = 96: 56), so stacking lifts the N = 32 layers from 40 % to 86 % of the tensor rate.  It also triples the
  // epilogue's TMEM reads and adds 2 shuffles per output, so it only pays when the K loop is long (Cin >= ~96); for N tiles of
  // 64 (N = 192) it measured no gain (977 vs 978 TFLOP/s on 384->64@112), so the default applies it to N tiles <= 32 only.
  const int ns3_min_cin = g_opt.ns3_min_cin > 0 ? g_opt.ns3_min_cin : 96;
  const int ns3_max_bn = g_opt.ns3_min_cin > 0 ? 80 : 32;
  bool ns3 = !v1 && ksize == 3 && p.BN <= ns3_max_bn && 3 * p.BN <= 256 && (3 * p.BN) % 16 == 0 && !g_opt.no_ns3 && Cin >= ns3_min_cin;
  int ns3_one_cta = 0;
  p.BNA = ns3 ? 3 * p.BN : p.BN;
  p.idesc = make_idesc_bf16(128, p.BNA, 0, 0);
  p.n_wtiles = (Cin / BK) * taps;
  const size_t budget = 222 * 1024;
  const size_t fixed = 1024 + 512 + (size_t)p.BN * 4 * 3 + (size_t)(p.BN / 16) * 8;
  // ---- resident weights? (all [BN x BK] tiles of this N tile stay in smem for the CTA's lifetime)
  const size_t w_bytes = (size_t)p.n_wtiles * p.b_tile_bytes;
  int MT, SA, SB;
  // BatchNorm statistics in per-thread registers over the CTA's tiles (conv_tc2_kernel<.., STATR>): for N tiles of <= 32 channels every
  // epilogue warp owns ONE 32-channel chunk, so 64 accumulators replace the per-item shuffle butterfly + shared atomics, which stall on the
  // same MIO pipe that feeds the UMMA operands (ncu: short-scoreboard stalls on the shuffles).  168 registers -> one CTA per SM, made up
  // for by four output windows per super-tile.  Measured (scripts/bench_layers.py): 32->32 @ 224 0.174 -> 0.151 ms, 128->32 0.298 -> 0.244,
  // 224->32 0.462 -> 0.443, 32->64 @ 112 0.107 -> 0.081; N tiles of 64 lose (384->64: 0.374 -> 0.469) and keep the butterfly.
  const bool statr = !v1 && stats != nullptr && p.BN <= 32 && g_opt.stat_mode != 1;
  p.resident = 0;
  // Measured on B200 (scripts/bench_layers.py): keeping the weights resident pays when the MMAs are short
  // (BN <= 32: one [BN x BK] tile is only 2-4 KB, so per-tap barrier round trips dominate) or when K is tiny
  // (Cin <= 32, the level-0 data-gradient convs; 1x1 transposed-conv phases); otherwise streamed weights with >1 CTA per SM win.
  const bool want_res = g_opt.no_resident ? false : (g_opt.mt < 0 ? true : (p.BN <= 32 || Cin <= 32 || ksize == 1));
  if (!v1 && want_res && w_bytes + 2 * p.a_tile_bytes + fixed <= budget) {
    p.resident = 1;
    MT = g_opt.mt > 0 ? g_opt.mt : (statr ? 4 : 2);
    while (MT > 1 && (2 * MT * p.BNA > 512 || w_bytes + 2 * (size_t)MT * p.a_tile_bytes + fixed > budget)) MT >>= 1;
    SA = g_opt.sa > 0 ? g_opt.sa : 4;
    while (SA > 2 && w_bytes + (size_t)SA * MT * p.a_tile_bytes + fixed > budget) --SA;
    SB = 1;
  } else {
    // streamed weights.  Measured (scripts/sweep_stream.py, B200): two co-resident CTAs per SM beat deeper pipelines, and one
    // weight stage per kernel ROW (3 taps) beats one per tap whenever it still leaves room for two CTAs.
    MT = g_opt.mt > 0 ? g_opt.mt : 1;
    while (MT > 1 && (v1 ? MT * p.BN > 512 : 2 * MT * p.BNA > 512)) MT >>= 1;
    auto bytes = [&](int mt, int sa, int sb, int tb) { return (size_t)sa * mt * p.a_tile_bytes + (size_t)sb * tb * p.b_tile_bytes + fixed; };
    const size_t half_sm = (227 * 1024) / 2 - 1024;
    const bool can3 = !v1 && ksize == 3 && g_opt.tb != 1;
    int TB = 1;
    if (ns3) {
      // a 512-column TMEM allocation allows ONE CTA per SM (two accumulators of 3*BN columns, deep rings); 256 columns allow two
      // co-resident CTAs (BN = 64: a single accumulator each, the other CTA's MMAs cover the epilogue)
      ns3_one_cta = g_opt.ns3_mode ? (g_opt.ns3_mode == 2) : 1;
      if (ns3_one_cta) {
        MT = (g_opt.mt > 0 ? g_opt.mt : 1); while (MT > 1 && 2 * MT * p.BNA > 512) MT >>= 1;
        TB = 3; SA = 4; SB = 4;
        while (bytes(MT, SA, SB, TB) > budget && (SA > 2 || SB > 2)) { if (SB >= SA && SB > 2) --SB; else --SA; }
      } else {
        MT = 1; TB = 3; SA = 2; SB = 2;
      }
    } else
    // 1x1 (plain GEMM: the ViT / ChangeFormer Linear layers): a stage holds only BK/16 = 4 UMMAs, so two-deep rings expose the
    // TMA round trip; measured (scripts/dbg_gemm2.py, 13312 x 768 x 3072): SA=SB=2 677, 3 924, 4 999, 5 1000 TFLOP/s
    if (!v1 && ksize == 1 && bytes(MT, 4, 4, 1) <= budget) { TB = 1; SA = 4; SB = 4; }
    else if (can3 && (g_opt.tb == 3 || bytes(MT, 2, 2, 3) <= half_sm)) { TB = 3; SA = 2; SB = 2; }
    else if (!v1 && bytes(MT, 2, 4, 1) <= half_sm) { TB = 1; SA = 2; SB = 4; }
    else if (can3 && bytes(MT, 3, 3, 3) <= budget) { TB = 3; SA = 3; SB = 3; }
    else { TB = 1; SA = 3; SB = 4; }
    if (g_opt.sa > 0) SA = g_opt.sa;
    if (g_opt.sb > 0) SB = g_opt.sb;
    while (bytes(MT, SA, SB, TB) > budget && SB > 2) --SB;
    while (bytes(MT, SA, SB, TB) > budget && SA > 2) --SA;
    while (bytes(MT, SA, SB, TB) > budget && MT > 1) MT >>= 1;
    if (bytes(MT, SA, SB, TB) > budget) { if (TB == 3) { TB = 1; } }
    if (bytes(MT, SA, SB, TB) > budget) return KS_EUNSUPPORTED;
    if (ns3 && TB != 3) { ns3 = false; p.BNA = p.BN; p.idesc = make_idesc_bf16(128, p.BN, 0, 0); }   // stacking needs a kernel row per weight stage
    p.TB = TB; p.b_stage_bytes = (uint32_t)TB * p.b_tile_bytes;
  }
  p.MT = MT; p.SA = SA; p.SB = SB;
  p.n_super = (p.total_tiles + MT - 1) / MT;
  p.n_acc = (2 * MT * p.BNA <= 512) ? 2 : 1;
  if (ns3 && !p.resident && !ns3_one_cta && 2 * MT * p.BNA > 256) p.n_acc = 1;
  if (g_opt.nacc > 0) p.n_acc = g_opt.nacc;
  if (v1) p.n_acc = 1;
  uint32_t cols = 32; while (cols < (uint32_t)(p.n_acc * MT * p.BNA)) cols <<= 1;
  if (cols > 512) return KS_EUNSUPPORTED;
  p.tmem_cols = cols;
  const size_t smem = (size_t)SA * MT * p.a_tile_bytes + (p.resident ? w_bytes : (size_t)SB * p.b_stage_bytes) + fixed;
  const CUtensorMapSwizzle sw = (BK == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  for (int s = 0; s < srcs.n; ++s) {
    int rc = encode_act_map(&p.src[s], srcs.v[s], N, H, W, BK, 16, (ksize == 3) ? 10 : 8, sw);
    if (rc) return rc;
  }
  {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return KS_EDRIVER;
    cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cin * Cout * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)p.BN, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&p.wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void *)weight, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return KS_EDRIVER;
  }
  int rc;
  if (v1) {
    dim3 grid((unsigned)p.n_super, (unsigned)nt);
    if (BK == 64 && ksize == 3) rc = launch_conv_tc<64, 3>(p, grid, smem, st);
    else if (BK == 64 && ksize == 1) rc = launch_conv_tc<64, 1>(p, grid, smem, st);
    else if (BK == 32 && ksize == 3) rc = launch_conv_tc<32, 3>(p, grid, smem, st);
    else rc = launch_conv_tc<32, 1>(p, grid, smem, st);
    if (rc) return rc;
    if (stats) {
      ks_view_t dv; dv.ptr = dsts.v[0].ptr; dv.sn = dsts.v[0].sn; dv.sh = dsts.v[0].sh; dv.sw = dsts.v[0].sw; dv.C = dsts.v[0].C; dv._pad = 0;
      return ks_bn_stats(KS_BF16, N, H, W, &dv, stats, (void *)st);
    }
    return KS_OK;
  }
  // persistent grid: CTAs per SM limited by shared memory; all N tiles of a super-tile run concurrently
  int per_sm = (int)((227 * 1024) / (smem + 1024)); if (per_sm < 1) per_sm = 1; if (per_sm > 4) per_sm = 4;
  if (per_sm > (int)(512 / cols)) per_sm = (int)(512 / cols);   // tcgen05.alloc of a further CTA would block until TMEM columns free up
  int gx = (kNumSMs * per_sm) / nt; if (gx < 1) gx = 1; if (gx > p.n_super) gx = p.n_super;
  dim3 grid((unsigned)gx, (unsigned)nt);
  // 16 epilogue warps when only one CTA fits per SM anyway and every lane group has >= 4 work items per super-tile
  const int items = MT * ((p.BN + 31) / 32);
  // measured (scripts/bench_layers.py, ew8 / ew16): 32->224 dgrad 0.532 -> 0.465 ms, 32->64 0.121 -> 0.102; with two N tiles (64->384) 4 % slower
  if (statr) { per_sm = 1; gx = kNumSMs / nt; if (gx < 1) gx = 1; if (gx > p.n_super) gx = p.n_super; grid = dim3((unsigned)gx, (unsigned)nt); }
  const bool ew16 = statr ? false : (g_opt.ew ? (g_opt.ew == 16) : (per_sm == 1 && items >= 4 && nt == 1));
#define KS_LAUNCH(BKv, KSv, NS3v) (statr ? launch_conv_tc2<BKv, KSv, NS3v, 8, true>(p, grid, smem, st) : \
    (ew16 ? launch_conv_tc2<BKv, KSv, NS3v, 16, false>(p, grid, smem, st) : launch_conv_tc2<BKv, KSv, NS3v, 8, false>(p, grid, smem, st)))
  if (BK == 64 && ksize == 3) rc = ns3 ? KS_LAUNCH(64, 3, true) : KS_LAUNCH(64, 3, false);
  else if (BK == 64 && ksize == 1) rc = KS_LAUNCH(64, 1, false);
  else if (BK == 32 && ksize == 3) rc = ns3 ? KS_LAUNCH(32, 3, true) : KS_LAUNCH(32, 3, false);
  else rc = KS_LAUNCH(32, 1, false);
#undef KS_LAUNCH
  return rc;
}

}  // namespace ks

extern "C" int ks_reset_options(void) {
  ks::g_opt = ks::Options{};
  return KS_OK;
}

extern "C" int ks_set_option(const char *name, int value) {
  if (!name) return KS_EINVAL;
  auto eq = [&](const char *s) { const char *a = name; while (*a && *a == *s) { ++a; ++s; } return *a == 0 && *s == 0; };
  if (eq("tc_mt")) ks::g_opt.mt = value;
  else if (eq("tc_bo_mode")) ks::g_opt.bo_mode = value;
  else if (eq("tc_disable")) ks::g_opt.tc_disable = value;
  else if (eq("wgrad_tc_disable")) ks::g_opt.wgrad_tc_disable = value;
  else if (eq("tc_sa")) ks::g_opt.sa = value;
  else if (eq("tc_sb")) ks::g_opt.sb = value;
  else if (eq("tc_v1")) ks::g_opt.v1 = value;
  else if (eq("tc_no_resident")) ks::g_opt.no_resident = value;
  else if (eq("tc_tb")) ks::g_opt.tb = value;
  else if (eq("tc_ns3_min_cin")) ks::g_opt.ns3_min_cin = value;   // > 0: stack column taps from this many input channels (default 96) and for N tiles up to 80
  else if (eq("tc_ns3_mode")) ks::g_opt.ns3_mode = value;         // streamed + stacked: 1 = two CTAs per SM (256 TMEM columns), 2 = one CTA per SM
  else if (eq("tc_nacc")) ks::g_opt.nacc = value;
  else if (eq("tc_ew")) ks::g_opt.ew = value;             // epilogue warps: 0 auto, 8, 16
  else if (eq("loss_chunks")) ks::g_opt.loss_chunks = value;   // perf experiments: CTAs per sample of the CE+Dice passes
  else if (eq("loss_no_bulk")) ks::g_opt.loss_no_bulk = value;   // 1 = register-staged CE+Dice passes (A/B comparisons)
  else if (eq("loss_variant")) ks::g_opt.loss_variant = value;   // 1 = two-pass CE+Dice kernels also when the resident single pass fits (A/B comparisons)
  else if (eq("loss_no_pdl")) ks::g_opt.loss_no_pdl = value;   // 1 = plain stream order between the two CE+Dice passes
  else if (eq("tc_no_ns3")) ks::g_opt.no_ns3 = value;   // 1 = one UMMA per tap also for narrow N tiles (A/B comparisons)
  else if (eq("wgrad_mode")) ks::g_opt.wgrad_mode = value;   // 0 auto (tap stacking along N), 2 = halo kernel v2
  else if (eq("ln_rows")) ks::g_opt.ln_rows = value;     // perf experiments: rows per thread of the LayerNorm-backward column pass
  else if (eq("cs_rows")) ks::g_opt.cs_rows = value;     // perf experiments: rows per row lane and CTA of ks_channel_sum (default 32)
  else if (eq("ew_cap")) ks::g_opt.ew_cap = value;       // perf experiments: CTAs per SM of the BatchNorm passes' grids (0 = default 8)
  else if (eq("stem_simt")) ks::g_opt.stem_simt = value;   // 1 = CUDA-core stem kernels also for bf16 / Cin == 2 (A/B comparisons)
  else if (eq("ecam_simt")) ks::g_opt.ecam_simt = value;   // 1 = CUDA-core ECAM final pass also for bf16 (A/B comparisons)
  else if (eq("xatt_umma")) ks::g_opt.xatt_umma = value;   // 1 = tcgen05 forward of the ChangeFormer spatial-reduction attention (default: mma.sync)
  else if (eq("tc_stat_mode")) ks::g_opt.stat_mode = value;   // 1 = shuffle butterfly for every BatchNorm-statistics epilogue (A/B comparisons)
  else if (eq("cf_scalar")) ks::g_opt.cf_scalar = value;   // 1 = scalar Dropout / DropPath kernels and 64-bit index arithmetic in im2col / col2im (A/B comparisons)
  else if (eq("dwconv_simple")) ks::g_opt.dwconv_simple = value;   // depth-wise conv kernels: 0 = shared-memory tiles (bf16) / 2x2 blocks (fp32), 1 = one output per thread, 2 = 2x2 register blocks, 3 = tiles also for the weight gradient of tiny images (A/B, tests)
  else if (eq("att_no_umma")) ks::g_opt.att_no_umma = value;   // 1 = mma.sync attention forward instead of the tcgen05 kernel (A/B comparisons)
  else if (eq("att_simt")) ks::g_opt.att_simt = value;   // 1 = CUDA-core attention kernels also for bf16 (A/B comparisons)
  else if (eq("tc_debug")) ks::g_opt.debug = value;   // perf experiments only: bit0 skip epilogue stores, bit1 skip TMEM loads, bit2 skip MMAs
  else return KS_EINVAL;
  return KS_OK;
}
