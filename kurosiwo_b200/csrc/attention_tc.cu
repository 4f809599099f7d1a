// ViT self-attention forward on the 5th-generation tensor cores: TMA -> shared memory -> tcgen05.mma -> TMEM.
//
//   S = Q K^T (per image and head: T <= 256 tokens, dh = 64) ; P = softmax(S * scale) ; O = P V
//   reference: models/vision_transformer.py:53-66 (two torch.matmul + nn.Softmax + einops rearranges)
//
// One CTA per (128-query tile, head, image).  All three operand tiles come straight out of the fused qkv matrix
// [B*Tp, 3*heads*64] through ONE tensor map (box = 64 features x 128 token rows, SWIZZLE_128B): Q and K are K-major operands
// (the contraction runs over the 64 features of a row), V is the MN-major B operand of the second product (the contraction
// runs over the token rows, the 64 features are the N dimension) - exactly the layout the rows already have, no transposes.
//   UMMA 1: S[128 x Tp]  = Q[128 x 64] K[Tp x 64]^T    -> TMEM columns [0, Tp)        (4 UMMAs of 128 x Tp x 16)
//   softmax: 4 warps, one thread per query row, three passes over the TMEM row (max | sum of exp | normalised bf16 P).
//            P goes to global memory (the backward reads it) and, in the canonical K-major SWIZZLE_128B layout, to the shared
//            memory that held Q and K (dead once UMMA 1 has completed) as the A operand of
//   UMMA 2: O[128 x 64]  = P[128 x Tp] V[Tp x 64]      -> TMEM columns [0, 64) (S is dead)  (Tp/16 UMMAs of 128 x 64 x 16)
// 96 KB of shared memory and 256 TMEM columns per CTA: two CTAs per SM overlap each other's load / MMA / softmax phases.
#include "common.cuh"
#include "tc_common.cuh"

namespace ks {

using namespace tc;

struct alignas(64) AttnTcParams {
  CUtensorMap qkv;               // dims (3*inner, Tp, B), box (64, 128, 1), SWIZZLE_128B
  __nv_bfloat16 *out, *probs;
  int T, Tp, heads, inner;
  float scale_log2e;             // scale * log2(e): exp((s - m) * scale) = exp2(s * c - m * c)
  uint32_t idesc_s, idesc_o;
};

__device__ __forceinline__ float att_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// shared memory map (bytes from the 1024-aligned base): P (A operand of UMMA 2, 4 chunks of [128 rows x 64 keys]) aliases Q | K | pad
constexpr uint32_t AT_Q = 0, AT_K = 16384, AT_V = 65536, AT_BAR = 98304, AT_SMEM = 98304 + 128 + 1024;

__global__ void __launch_bounds__(160, 2) attention_fwd_umma_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *sm = smem_raw + (base - raw);
  const uint32_t qk_full = base + AT_BAR, v_full = qk_full + 8, s_full = qk_full + 16, p_ready = qk_full + 24, o_full = qk_full + 32;
  const uint32_t tmem_slot = qk_full + 40;
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(sm + AT_BAR + 40);
  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const int mtile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int T = p.T, Tp = p.Tp;

  if (warp == 0) {
    if (elect_one()) {
      prefetch_tmap(&p.qkv);
      mbar_init(qk_full, 1); mbar_init(v_full, 1); mbar_init(s_full, 1); mbar_init(p_ready, 128); mbar_init(o_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===== TMA producer + MMA issuer (one elected lane; the warp stays converged around the waits) =====
    if (elect_one()) {
      const int cq = h * 64, ck = p.inner + h * 64, cv = 2 * p.inner + h * 64;
      mbar_expect_tx(qk_full, 3u * 16384u);
      tma_load_3d(base + AT_Q, &p.qkv, cq, mtile * 128, b, qk_full);            // rows >= Tp are out of bounds -> zero
      tma_load_3d(base + AT_K, &p.qkv, ck, 0, b, qk_full);
      tma_load_3d(base + AT_K + 16384, &p.qkv, ck, 128, b, qk_full);
      mbar_expect_tx(v_full, 2u * 16384u);
      tma_load_3d(base + AT_V, &p.qkv, cv, 0, b, v_full);
      tma_load_3d(base + AT_V + 16384, &p.qkv, cv, 128, b, v_full);
    }
    __syncwarp();
    mbar_wait(qk_full, 0);
    tc_fence_after();
    // K-major SWIZZLE_128B descriptors: SBO = 8 rows * 128 B, LBO unused (1), 32 bytes per 16-element K step
    const uint32_t kmaj_hi = (uint32_t)((1024u >> 4) & 0x3FFFu) | (1u << 14) | (LAYOUT_SW128 << 29);
    if (elect_one()) {
      const uint32_t q_lo = (((base + AT_Q) & 0x3FFFFu) >> 4) | (1u << 16), k_lo = (((base + AT_K) & 0x3FFFFu) >> 4) | (1u << 16);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_base, ((uint64_t)kmaj_hi << 32) | (uint64_t)(q_lo + 2u * k), ((uint64_t)kmaj_hi << 32) | (uint64_t)(k_lo + 2u * k),
                  p.idesc_s, (uint32_t)k);
      tc_commit(s_full);
    }
    __syncwarp();
    mbar_wait(v_full, 0);
    mbar_wait(p_ready, 0);
    tc_fence_after();
    if (elect_one()) {
      // A = P: K-major, chunk (ks >> 2) of 16 KB, 32 bytes per K step inside a chunk.  B = V: MN-major (64 features contiguous per
      // token row), SWIZZLE_128B, SBO = 8 rows * 128 B, one 64-wide N group (LBO unused), 16 token rows = 2048 bytes per K step.
      const uint32_t p_lo = (((base + AT_Q) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t v_lo = (((base + AT_V) & 0x3FFFFu) >> 4) | (((16384u >> 4) & 0x3FFFu) << 16);
      const int nks = Tp >> 4;
      for (int ks = 0; ks < nks; ++ks) {
        const uint64_t ad = ((uint64_t)kmaj_hi << 32) | (uint64_t)(p_lo + (uint32_t)(ks >> 2) * 1024u + 2u * (uint32_t)(ks & 3));
        const uint64_t bd = ((uint64_t)kmaj_hi << 32) | (uint64_t)(v_lo + (uint32_t)ks * 128u);
        umma_bf16(tmem_base, ad, bd, p.idesc_o, (uint32_t)ks);
      }
      tc_commit(o_full);
    }
    __syncwarp();
  } else {
    // ===== softmax + epilogue: warps 1..4, TMEM lane group = warp % 4, one thread per query row =====
    const int lg = warp & 3;
    const int r = lg * 32 + lane;                       // row inside the tile
    const int row = mtile * 128 + r;                    // token index of this query
    const uint32_t trow = tmem_base + ((uint32_t)(lg * 32) << 16);
    const float c = p.scale_log2e;
    const int nch = (Tp + 31) >> 5;                     // 32-column chunks (the last may be 16 wide: Tp % 32 == 16)
    mbar_wait(s_full, 0);
    tc_fence_after();
    // pass 1: row maximum over the valid keys
    float m = -INFINITY;
    for (int ch = 0; ch < nch; ++ch) {
      uint32_t v[32];
      const int c0 = ch << 5;
      if (Tp - c0 >= 32) tmem_ld32(trow + c0, v);
      else { uint32_t v16[16]; tmem_ld16(trow + c0, v16);
#pragma unroll
        for (int i = 0; i < 16; ++i) { v[i] = v16[i]; v[16 + i] = 0xff800000u; } }
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) if (c0 + i < T) m = fmaxf(m, __uint_as_float(v[i]));
    }
    const float mc = m * c;
    // pass 2: sum of exp
    float l = 0.f;
    for (int ch = 0; ch < nch; ++ch) {
      uint32_t v[32];
      const int c0 = ch << 5;
      if (Tp - c0 >= 32) tmem_ld32(trow + c0, v);
      else { uint32_t v16[16]; tmem_ld16(trow + c0, v16);
#pragma unroll
        for (int i = 0; i < 16; ++i) { v[i] = v16[i]; v[16 + i] = 0xff800000u; } }
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) if (c0 + i < T) l += att_ex2(fmaf(__uint_as_float(v[i]), c, -mc));
    }
    const bool row_ok = row < T;
    const float inv = row_ok ? 1.f / l : 0.f;           // rows >= T (padding tokens / out-of-bounds rows of the last tile): P = 0, O = 0
    // pass 3: normalised P as bf16 -> global (for the backward) and -> shared memory (A operand of UMMA 2)
    __nv_bfloat16 *prow = p.probs + (((long long)b * p.heads + h) * Tp + row) * Tp;
    for (int ch = 0; ch < nch; ++ch) {
      uint32_t v[32];
      const int c0 = ch << 5;
      const bool wide = (Tp - c0) >= 32;
      if (wide) tmem_ld32(trow + c0, v);
      else { uint32_t v16[16]; tmem_ld16(trow + c0, v16);
#pragma unroll
        for (int i = 0; i < 16; ++i) { v[i] = v16[i]; v[16 + i] = 0xff800000u; } }
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float e0 = (c0 + 2 * i < T) ? att_ex2(fmaf(__uint_as_float(v[2 * i]), c, -mc)) * inv : 0.f;
        const float e1 = (c0 + 2 * i + 1 < T) ? att_ex2(fmaf(__uint_as_float(v[2 * i + 1]), c, -mc)) * inv : 0.f;
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(e0, e1);
        pk[i] = *reinterpret_cast<const uint32_t *>(&h2);
      }
      const int nv = wide ? 4 : 2;                      // 16-byte vectors (8 keys each) of this chunk
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q < nv) {
          const uint4 u = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          const int j = (c0 >> 3) + q;                  // 16-byte chunk index along the key axis
          // canonical K-major SWIZZLE_128B tile: [64-key chunk][row][128 B], 16-byte units XOR-ed with (row & 7)
          *reinterpret_cast<uint4 *>(sm + AT_Q + (uint32_t)(j >> 3) * 16384u + (uint32_t)r * 128u + (uint32_t)(((j & 7) ^ (r & 7)) << 4)) = u;
          if (row < Tp) *reinterpret_cast<uint4 *>(prow + c0 + 8 * q) = u;
        }
      }
    }
    fence_proxy_async();                                // generic-proxy writes of P -> visible to the tensor core's async proxy
    tc_fence_before();
    mbar_arrive(p_ready);
    // epilogue: O (already normalised: P carries 1/l) -> bf16 -> out[row][h*64 .. h*64+63]
    mbar_wait(o_full, 0);
    tc_fence_after();
    uint32_t o0[32], o1[32];
    tmem_ld32(trow, o0);
    tmem_ld32(trow + 32, o1);
    tmem_ld_wait();
    if (row < Tp) {
      __nv_bfloat16 *orow = p.out + ((long long)b * Tp + row) * p.inner + h * 64;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t *src = (q < 4) ? (o0 + 8 * q) : (o1 + 8 * (q - 4));
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(src[2 * i]), __uint_as_float(src[2 * i + 1]));
          w[i] = *reinterpret_cast<const uint32_t *>(&h2);
        }
        *reinterpret_cast<uint4 *>(orow + 8 * q) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// ------------------------------------------------------------------------------------------------------------------------
// Backward, part 1 (one CTA per 128-query tile, head, image): the same pipeline with other operands.
//   UMMA 1: dP[128 x Tp] = dO[128 x 64] V[Tp x 64]^T                       (both K-major)            -> TMEM [0, Tp)
//   rows:   d = sum_j dP_j P_j ; dS = P (dP - d)  (P: the bf16 probabilities the forward wrote) -> bf16 -> global scratch (part 2
//           reads it) and -> shared memory over dO | V (dead), K-major SWIZZLE_128B
//   UMMA 2: dQ[128 x 64] = dS[128 x Tp] K[Tp x 64]   (K as MN-major B)      -> TMEM [0, 64) -> * scale -> dqkv[:, h*64 ..]
// ------------------------------------------------------------------------------------------------------------------------
struct alignas(64) AttnBwdRowsParams {
  CUtensorMap qkv;               // dims (3*inner, Tp, B), box (64, 128, 1)
  CUtensorMap dout;              // dims (inner, Tp, B), box (64, 128, 1)
  const __nv_bfloat16 *probs;
  __nv_bfloat16 *ds, *dqkv;
  int T, Tp, heads, inner;
  float scale;
  uint32_t idesc_s, idesc_o;
};

__global__ void __launch_bounds__(160, 2) attention_bwd_rows_umma_kernel(const __grid_constant__ AttnBwdRowsParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *sm = smem_raw + (base - raw);
  // smem: [dO 16 KB][V 32 KB][pad 16 KB] (= dS, 4 chunks, after UMMA 1) [K 32 KB]
  const uint32_t qk_full = base + AT_BAR, v_full = qk_full + 8, s_full = qk_full + 16, p_ready = qk_full + 24, o_full = qk_full + 32;
  const uint32_t tmem_slot = qk_full + 40;
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(sm + AT_BAR + 40);
  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const int mtile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int T = p.T, Tp = p.Tp;
  if (warp == 0) {
    if (elect_one()) {
      prefetch_tmap(&p.qkv); prefetch_tmap(&p.dout);
      mbar_init(qk_full, 1); mbar_init(v_full, 1); mbar_init(s_full, 1); mbar_init(p_ready, 128); mbar_init(o_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (warp == 0) {
    if (elect_one()) {
      const int ck = p.inner + h * 64, cv = 2 * p.inner + h * 64;
      mbar_expect_tx(qk_full, 3u * 16384u);
      tma_load_3d(base + AT_Q, &p.dout, h * 64, mtile * 128, b, qk_full);       // dO tile
      tma_load_3d(base + AT_K, &p.qkv, cv, 0, b, qk_full);                       // V (K-major B of UMMA 1)
      tma_load_3d(base + AT_K + 16384, &p.qkv, cv, 128, b, qk_full);
      mbar_expect_tx(v_full, 2u * 16384u);
      tma_load_3d(base + AT_V, &p.qkv, ck, 0, b, v_full);                        // K (MN-major B of UMMA 2)
      tma_load_3d(base + AT_V + 16384, &p.qkv, ck, 128, b, v_full);
    }
    __syncwarp();
    mbar_wait(qk_full, 0);
    tc_fence_after();
    const uint32_t kmaj_hi = (uint32_t)((1024u >> 4) & 0x3FFFu) | (1u << 14) | (LAYOUT_SW128 << 29);
    if (elect_one()) {
      const uint32_t a_lo = (((base + AT_Q) & 0x3FFFFu) >> 4) | (1u << 16), b_lo = (((base + AT_K) & 0x3FFFFu) >> 4) | (1u << 16);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_base, ((uint64_t)kmaj_hi << 32) | (uint64_t)(a_lo + 2u * k), ((uint64_t)kmaj_hi << 32) | (uint64_t)(b_lo + 2u * k),
                  p.idesc_s, (uint32_t)k);
      tc_commit(s_full);
    }
    __syncwarp();
    mbar_wait(v_full, 0);
    mbar_wait(p_ready, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t a_lo = (((base + AT_Q) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t b_lo = (((base + AT_V) & 0x3FFFFu) >> 4) | (((16384u >> 4) & 0x3FFFu) << 16);
      const int nks = Tp >> 4;
      for (int ks = 0; ks < nks; ++ks) {
        const uint64_t ad = ((uint64_t)kmaj_hi << 32) | (uint64_t)(a_lo + (uint32_t)(ks >> 2) * 1024u + 2u * (uint32_t)(ks & 3));
        const uint64_t bd = ((uint64_t)kmaj_hi << 32) | (uint64_t)(b_lo + (uint32_t)ks * 128u);
        umma_bf16(tmem_base, ad, bd, p.idesc_o, (uint32_t)ks);
      }
      tc_commit(o_full);
    }
    __syncwarp();
  } else {
    const int lg = warp & 3;
    const int r = lg * 32 + lane;
    const int row = mtile * 128 + r;
    const uint32_t trow = tmem_base + ((uint32_t)(lg * 32) << 16);
    const int nch = (Tp + 31) >> 5;
    const bool row_ok = row < T;                        // P rows >= T are zero: dS = 0 there
    const long long poff = (((long long)b * p.heads + h) * Tp + min(row, Tp - 1)) * Tp;
    const __nv_bfloat16 *prow = p.probs + poff;
    // this row of P (bf16, Tp <= 256 values = up to 32 x 16 bytes) is read ONCE, into registers, before UMMA 1 has even finished:
    // the loads overlap the TMA / MMA latency and both passes below run out of registers
    uint4 pq[8][4];
#pragma unroll
    for (int ch = 0; ch < 8; ++ch)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        pq[ch][q] = (ch < nch && row_ok && (ch * 32 + q * 8) < Tp) ? __ldg(reinterpret_cast<const uint4 *>(prow + ch * 32) + q) : make_uint4(0u, 0u, 0u, 0u);
    mbar_wait(s_full, 0);
    tc_fence_after();
    // pass 1: d = sum_j dP_j P_j
    float d = 0.f;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      if (ch < nch) {
        uint32_t v[32];
        const int c0 = ch << 5;
        const bool wide = (Tp - c0) >= 32;
        if (wide) tmem_ld32(trow + c0, v);
        else { uint32_t v16[16]; tmem_ld16(trow + c0, v16);
#pragma unroll
          for (int i = 0; i < 16; ++i) { v[i] = v16[i]; v[16 + i] = 0u; } }
        tmem_ld_wait();
        const uint32_t *pw = reinterpret_cast<const uint32_t *>(pq[ch]);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 pf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&pw[i]));
          d = fmaf(__uint_as_float(v[2 * i]), pf.x, d);
          d = fmaf(__uint_as_float(v[2 * i + 1]), pf.y, d);
        }
      }
    }
    // pass 2: dS = P (dP - d) -> bf16 -> global scratch + shared memory (A operand of UMMA 2)
    __nv_bfloat16 *dsrow = p.ds + poff;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      if (ch < nch) {
        uint32_t v[32];
        const int c0 = ch << 5;
        const bool wide = (Tp - c0) >= 32;
        if (wide) tmem_ld32(trow + c0, v);
        else { uint32_t v16[16]; tmem_ld16(trow + c0, v16);
#pragma unroll
          for (int i = 0; i < 16; ++i) { v[i] = v16[i]; v[16 + i] = 0u; } }
        tmem_ld_wait();
        const uint32_t *pw = reinterpret_cast<const uint32_t *>(pq[ch]);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 pf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&pw[i]));
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(pf.x * (__uint_as_float(v[2 * i]) - d), pf.y * (__uint_as_float(v[2 * i + 1]) - d));
          pk[i] = *reinterpret_cast<const uint32_t *>(&h2);
        }
        const int nv = wide ? 4 : 2;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q < nv) {
            const uint4 u = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            const int j = (c0 >> 3) + q;
            *reinterpret_cast<uint4 *>(sm + AT_Q + (uint32_t)(j >> 3) * 16384u + (uint32_t)r * 128u + (uint32_t)(((j & 7) ^ (r & 7)) << 4)) = u;
            if (row < Tp) *reinterpret_cast<uint4 *>(dsrow + c0 + 8 * q) = u;
          }
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(p_ready);
    mbar_wait(o_full, 0);
    tc_fence_after();
    uint32_t o0[32], o1[32];
    tmem_ld32(trow, o0);
    tmem_ld32(trow + 32, o1);
    tmem_ld_wait();
    if (row < Tp) {
      __nv_bfloat16 *orow = p.dqkv + ((long long)b * Tp + row) * (3LL * p.inner) + h * 64;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t *src = (q < 4) ? (o0 + 8 * q) : (o1 + 8 * (q - 4));
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(src[2 * i]) * p.scale, __uint_as_float(src[2 * i + 1]) * p.scale);
          w[i] = *reinterpret_cast<const uint32_t *>(&h2);
        }
        *reinterpret_cast<uint4 *>(orow + 8 * q) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// ------------------------------------------------------------------------------------------------------------------------
// Backward, part 2: the two products that contract over the QUERIES, out[128 keys x 64] = alpha * A[q x keys]^T B[q x 64]:
//   dK = scale * dS^T Q   and   dV = P^T dO.   A (dS or P, row-major [q][keys] in global memory) is the MN-major A operand (the
// 64 keys of a chunk are contiguous in a row, the contraction runs over the rows), B (Q or dO rows) the MN-major B operand: both tiles
// are plain TMA boxes of the arrays as they lie in memory.  One CTA per (key tile, {dK, dV}, head, image); 96 KB, 64 TMEM columns.
// ------------------------------------------------------------------------------------------------------------------------
struct alignas(64) AttnBwdColsParams {
  CUtensorMap a_ds, a_p;         // dims (Tp, Tp, B*heads), box (64, 128, 1)
  CUtensorMap qkv, dout;         // as above
  __nv_bfloat16 *dqkv;
  int T, Tp, heads, inner;
  float scale;
  uint32_t idesc;
};

__global__ void __launch_bounds__(160, 2) attention_bwd_cols_umma_kernel(const __grid_constant__ AttnBwdColsParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *sm = smem_raw + (base - raw);
  // smem: A = 2 key chunks x [256 q rows x 128 B] = 64 KB at 0 ; B = [256 q rows x 128 B] = 32 KB at 64 KB
  const uint32_t full = base + AT_BAR, o_full = full + 8, tmem_slot = full + 16;
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(sm + AT_BAR + 16);
  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const int kt = blockIdx.x >> 1, which = blockIdx.x & 1;     // which: 0 = dK, 1 = dV
  const int h = blockIdx.y, b = blockIdx.z;
  const int Tp = p.Tp;
  if (warp == 0) {
    if (elect_one()) {
      mbar_init(full, 1); mbar_init(o_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 64);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (warp == 0) {
    if (elect_one()) {
      const CUtensorMap *am = which ? &p.a_p : &p.a_ds;
      const CUtensorMap *bm = which ? &p.dout : &p.qkv;
      const int cb = h * 64;                                   // Q columns of qkv / dO columns of dout
      mbar_expect_tx(full, 6u * 16384u);
      for (int c = 0; c < 2; ++c)
        for (int rh = 0; rh < 2; ++rh)
          tma_load_3d(base + (uint32_t)c * 32768u + (uint32_t)rh * 16384u, am, kt * 128 + c * 64, rh * 128, b * p.heads + h, full);
      tma_load_3d(base + 65536u, bm, cb, 0, b, full);
      tma_load_3d(base + 65536u + 16384u, bm, cb, 128, b, full);
    }
    __syncwarp();
    mbar_wait(full, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t hi = (uint32_t)((1024u >> 4) & 0x3FFFu) | (1u << 14) | (LAYOUT_SW128 << 29);
      const uint32_t a_lo = ((base & 0x3FFFFu) >> 4) | (((32768u >> 4) & 0x3FFFu) << 16);            // LBO = stride between the two 64-key groups
      const uint32_t b_lo = (((base + 65536u) & 0x3FFFFu) >> 4) | (((16384u >> 4) & 0x3FFFu) << 16);
      const int nks = Tp >> 4;
      for (int ks = 0; ks < nks; ++ks)
        umma_bf16(tmem_base, ((uint64_t)hi << 32) | (uint64_t)(a_lo + (uint32_t)ks * 128u), ((uint64_t)hi << 32) | (uint64_t)(b_lo + (uint32_t)ks * 128u),
                  p.idesc, (uint32_t)ks);
      tc_commit(o_full);
    }
    __syncwarp();
  } else {
    const int lg = warp & 3;
    const int r = lg * 32 + lane;
    const int key = kt * 128 + r;
    const uint32_t trow = tmem_base + ((uint32_t)(lg * 32) << 16);
    const float alpha = which ? 1.f : p.scale;
    mbar_wait(o_full, 0);
    tc_fence_after();
    uint32_t o0[32], o1[32];
    tmem_ld32(trow, o0);
    tmem_ld32(trow + 32, o1);
    tmem_ld_wait();
    if (key < Tp) {
      __nv_bfloat16 *orow = p.dqkv + ((long long)b * Tp + key) * (3LL * p.inner) + (which ? 2 : 1) * p.inner + h * 64;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t *src = (q < 4) ? (o0 + 8 * q) : (o1 + 8 * (q - 4));
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(src[2 * i]) * alpha, __uint_as_float(src[2 * i + 1]) * alpha);
          w[i] = *reinterpret_cast<const uint32_t *>(&h2);
        }
        *reinterpret_cast<uint4 *>(orow + 8 * q) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

static int att_map3(CUtensorMap *m, const void *ptr, long long d0, long long d1, long long d2) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return KS_EDRIVER;
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)d0 * 2, (cuuint64_t)d0 * d1 * 2};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void *)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? KS_OK : KS_EDRIVER;
}

int attention_bwd_umma(int B, int T, int Tp, int heads, const void *qkv, const void *probs, const void *dout, float scale, void *dqkv,
                       void *ds, cudaStream_t st) {
  if (Tp % 16 || Tp > 256 || Tp < 16 || T > Tp || T < 1 || (Tp * 2) % 16) return KS_EUNSUPPORTED;
  for (const void *q : {qkv, probs, dout, (const void *)dqkv, (const void *)ds}) if (((uintptr_t)q) % 16) return KS_EUNSUPPORTED;
  const int inner = heads * 64;
  AttnBwdRowsParams pr;
  int rc = att_map3(&pr.qkv, qkv, 3LL * inner, Tp, B); if (rc) return rc;
  rc = att_map3(&pr.dout, dout, inner, Tp, B); if (rc) return rc;
  pr.probs = (const __nv_bfloat16 *)probs; pr.ds = (__nv_bfloat16 *)ds; pr.dqkv = (__nv_bfloat16 *)dqkv;
  pr.T = T; pr.Tp = Tp; pr.heads = heads; pr.inner = inner; pr.scale = scale;
  pr.idesc_s = make_idesc_bf16(128, Tp, 0, 0);
  pr.idesc_o = make_idesc_bf16(128, 64, 0, 1);
  AttnBwdColsParams pc;
  rc = att_map3(&pc.a_ds, ds, Tp, Tp, (long long)B * heads); if (rc) return rc;
  rc = att_map3(&pc.a_p, probs, Tp, Tp, (long long)B * heads); if (rc) return rc;
  pc.qkv = pr.qkv; pc.dout = pr.dout;
  pc.dqkv = (__nv_bfloat16 *)dqkv; pc.T = T; pc.Tp = Tp; pc.heads = heads; pc.inner = inner; pc.scale = scale;
  pc.idesc = make_idesc_bf16(128, 64, 1, 1);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_rows_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AT_SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(attention_bwd_cols_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AT_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  const unsigned mt = (unsigned)((Tp + 127) / 128);
  attention_bwd_rows_umma_kernel<<<dim3(mt, (unsigned)heads, (unsigned)B), 160, AT_SMEM, st>>>(pr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  attention_bwd_cols_umma_kernel<<<dim3(2 * mt, (unsigned)heads, (unsigned)B), 160, AT_SMEM, st>>>(pc);
  return (int)cudaGetLastError();
}

// Host side.  Returns KS_EUNSUPPORTED when the shape does not fit this kernel (the caller falls back to the mma.sync kernel).
int attention_fwd_umma(int B, int T, int Tp, int heads, const void *qkv, float scale, void *out, void *probs, cudaStream_t st) {
  if (Tp % 16 || Tp > 256 || Tp < 16 || T > Tp || T < 1) return KS_EUNSUPPORTED;
  if ((((uintptr_t)qkv) % 16) || (((uintptr_t)out) % 16) || (((uintptr_t)probs) % 16)) return KS_EUNSUPPORTED;
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return KS_EDRIVER;
  AttnTcParams p;
  const int inner = heads * 64;
  cuuint64_t dims[3] = {(cuuint64_t)(3 * inner), (cuuint64_t)Tp, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)(3 * inner) * 2, (cuuint64_t)Tp * 3 * inner * 2};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&p.qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void *)qkv, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return KS_EDRIVER;
  p.out = (__nv_bfloat16 *)out; p.probs = (__nv_bfloat16 *)probs;
  p.T = T; p.Tp = Tp; p.heads = heads; p.inner = inner;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.idesc_s = make_idesc_bf16(128, Tp, 0, 0);
  p.idesc_o = make_idesc_bf16(128, 64, 0, 1);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_fwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AT_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  dim3 grid((unsigned)((Tp + 127) / 128), (unsigned)heads, (unsigned)B);
  attention_fwd_umma_kernel<<<grid, 160, AT_SMEM, st>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace ks
