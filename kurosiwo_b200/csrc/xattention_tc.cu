// ChangeFormer spatial-reduction attention on the 5th-generation tensor cores (bf16, dh = 64, <= 64 reduced keys):
// TMA -> shared memory -> tcgen05.mma -> TMEM, the same pipeline as the ViT kernels in attention_tc.cu with other operands.
//
//   reference: models/changeformer.py:186-208   q = Linear(x); kv = Linear(LN(conv_{k=s=sr}(x))); attn = softmax(q k^T * scale);
//                                               attn = Dropout(attn); x = attn v
//
// Forward, one CTA per (128-query tile, head, image):
//   UMMA 1: S[128 x 64]  = Q[128 x 64] K[64 x 64]^T          (both K-major: the contraction runs over the 64 features of a row;
//                                                             key rows >= Nk are out of the tensor map's bounds -> zero)
//   softmax: 4 warps, one thread per query row: the 64 scores of the row from TMEM, max / exp2 / sum in registers, the bf16
//            probabilities (pre-dropout: what the backward reads) go through a linear staging buffer to ONE contiguous, coalesced
//            global block per tile ([Nq][Nk] rows are 98 bytes at Nk = 49: per-thread row stores would scatter 2-byte writes);
//            the dropped probabilities (stateless RNG of cformer.cu: a pure function of seed, step, site and the element index) go,
//            in the canonical K-major SWIZZLE_128B layout, over the dead Q tile as the A operand of
//   UMMA 2: O[128 x 64]  = Pdrop[128 x 64] V[64 x 64]         (V as the MN-major B operand: its rows as they lie in memory)
// 49 KB of shared memory and 64 TMEM columns per CTA: four CTAs per SM overlap each other's load / MMA / softmax phases.
#include "common.cuh"
#include "tc_common.cuh"

namespace ks {

using namespace tc;

// stateless dropout stream shared with cformer.cu (same constants: the oracle reproduces these masks)
__device__ __forceinline__ unsigned long long xq_mix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ unsigned long long xq_key(unsigned long long seed, const int *step_ptr, int site) {
  const unsigned long long step = step_ptr ? (unsigned long long)(unsigned int)*step_ptr : 0ull;
  return xq_mix64(seed ^ (step << 32) ^ ((unsigned long long)(unsigned int)site * 0x632BE59BD9B4E019ull));
}
__device__ __forceinline__ float xq_keep(unsigned long long key, unsigned long long idx, float p, float inv_keep) {
  const unsigned long long r = xq_mix64(key + idx);
  return ((float)(r >> 40) * (1.0f / 16777216.0f) >= p) ? inv_keep : 0.f;
}

struct alignas(64) XAttnFwdParams {
  CUtensorMap q;                 // dims (inner, Nq, B), box (64, 128, 1), SWIZZLE_128B
  CUtensorMap kv;                // dims (2*inner, Nk, B), box (64, 64, 1)
  __nv_bfloat16 *out, *probs;
  long long ldo;
  int Nq, Nk, heads, inner;
  float scale_log2e, pdrop, ikeep;
  unsigned long long seed;
  const int *step_ptr;
  int site;
  uint32_t idesc_s, idesc_o;
};

// shared memory map (bytes from the 1024-aligned base)
constexpr uint32_t XQ_Q = 0, XQ_K = 16384, XQ_V = 24576, XQ_PL = 32768, XQ_BAR = 49152 + 64, XQ_SMEM = XQ_BAR + 128 + 1024;

__device__ __forceinline__ float xq_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(160, 4) xattention_fwd_umma_kernel(const __grid_constant__ XAttnFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *sm = smem_raw + (base - raw);
  const uint32_t qk_full = base + XQ_BAR, v_full = qk_full + 8, s_full = qk_full + 16, p_ready = qk_full + 24, o_full = qk_full + 32;
  const uint32_t tmem_slot = qk_full + 40;
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(sm + XQ_BAR + 40);
  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const int mtile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int Nq = p.Nq, Nk = p.Nk;

  if (warp == 0) {
    if (elect_one()) {
      prefetch_tmap(&p.q); prefetch_tmap(&p.kv);
      mbar_init(qk_full, 1); mbar_init(v_full, 1); mbar_init(s_full, 1); mbar_init(p_ready, 128); mbar_init(o_full, 1);
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
    // ===== TMA producer + MMA issuer =====
    if (elect_one()) {
      mbar_expect_tx(qk_full, 16384u + 8192u);
      tma_load_3d(base + XQ_Q, &p.q, h * 64, mtile * 128, b, qk_full);           // query rows >= Nq: out of bounds -> zero
      tma_load_3d(base + XQ_K, &p.kv, h * 64, 0, b, qk_full);                     // key rows >= Nk: zero
      mbar_expect_tx(v_full, 8192u);
      tma_load_3d(base + XQ_V, &p.kv, p.inner + h * 64, 0, b, v_full);
    }
    __syncwarp();
    mbar_wait(qk_full, 0);
    tc_fence_after();
    const uint32_t kmaj_hi = (uint32_t)((1024u >> 4) & 0x3FFFu) | (1u << 14) | (LAYOUT_SW128 << 29);
    if (elect_one()) {
      const uint32_t q_lo = (((base + XQ_Q) & 0x3FFFFu) >> 4) | (1u << 16), k_lo = (((base + XQ_K) & 0x3FFFFu) >> 4) | (1u << 16);
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
      // A = Pdrop: K-major, 32 bytes per 16-key K step.  B = V: MN-major (64 features contiguous per key row), SWIZZLE_128B,
      // SBO = 8 rows * 128 B, one 64-wide N group, 16 key rows = 2048 bytes per K step.
      const uint32_t p_lo = (((base + XQ_Q) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t v_lo = (((base + XQ_V) & 0x3FFFFu) >> 4) | (((8192u >> 4) & 0x3FFFu) << 16);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_bf16(tmem_base, ((uint64_t)kmaj_hi << 32) | (uint64_t)(p_lo + 2u * (uint32_t)ks),
                  ((uint64_t)kmaj_hi << 32) | (uint64_t)(v_lo + (uint32_t)ks * 128u), p.idesc_o, (uint32_t)ks);
      tc_commit(o_full);
    }
    __syncwarp();
  } else {
    // ===== softmax + epilogue: warps 1..4, TMEM lane group = warp % 4, one thread per query row =====
    const int lg = warp & 3;
    const int r = lg * 32 + lane;                       // row inside the tile
    const int row = mtile * 128 + r;                    // query index inside the image
    const bool row_ok = row < Nq;
    const uint32_t trow = tmem_base + ((uint32_t)(lg * 32) << 16);
    const float c = p.scale_log2e;
    const unsigned long long dkey = xq_key(p.seed, p.step_ptr, p.site);
    const bool drop = p.pdrop > 0.f;
    // linear staging of the tile's probabilities: the tile is ONE contiguous block of global memory; the staging buffer starts at the
    // same offset modulo 16 bytes, so the copy below moves aligned 16-byte vectors on both sides
    const long long tile_e0 = (((long long)b * p.heads + h) * Nq + (long long)mtile * 128) * Nk;      // first element of the tile
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(p.probs + tile_e0)) & 15u);          // even: bf16 elements
    __nv_bfloat16 *stage = reinterpret_cast<__nv_bfloat16 *>(sm + XQ_PL + mis);
    mbar_wait(s_full, 0);
    tc_fence_after();
    uint32_t s0[32], s1[32];
    tmem_ld32(trow, s0);
    tmem_ld32(trow + 32, s1);
    tmem_ld_wait();
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (i < Nk) m = fmaxf(m, __uint_as_float(s0[i]));
      if (32 + i < Nk) m = fmaxf(m, __uint_as_float(s1[i]));
    }
    const float mc = m * c;
    float l = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float e0 = (i < Nk) ? xq_ex2(fmaf(__uint_as_float(s0[i]), c, -mc)) : 0.f;
      const float e1 = (32 + i < Nk) ? xq_ex2(fmaf(__uint_as_float(s1[i]), c, -mc)) : 0.f;
      s0[i] = __float_as_uint(e0); s1[i] = __float_as_uint(e1);
      l += e0 + e1;
    }
    const float inv = 1.f / l;
    const long long pbase = tile_e0 + (long long)r * Nk;              // this row's first element in probs
    uint32_t pk[32];                                                  // dropped probabilities, bf16 pairs
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float pq[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 2 * i + e;
        const float ev = __uint_as_float((j < 32) ? s0[j & 31] : s1[j & 31]);
        float v = 0.f;
        if (j < Nk) {
          const __nv_bfloat16 pb = __float2bfloat16_rn(ev * inv);     // the stored probability (pre-dropout)
          if (row_ok) stage[r * Nk + j] = pb;
          v = __bfloat162float(pb);
          if (drop && row_ok) v *= xq_keep(dkey, (unsigned long long)(pbase + j), p.pdrop, p.ikeep);
          if (!row_ok) v = 0.f;
        }
        pq[e] = v;
      }
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(pq[0], pq[1]);
      pk[i] = *reinterpret_cast<const uint32_t *>(&h2);
    }
    // canonical K-major SWIZZLE_128B tile over the dead Q tile: [row][128 B], 16-byte units XOR-ed with (row & 7)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 u = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      *reinterpret_cast<uint4 *>(sm + XQ_Q + (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4)) = u;
    }
    fence_proxy_async();                                // generic-proxy writes of P -> visible to the tensor core's async proxy
    tc_fence_before();
    mbar_arrive(p_ready);
    // coalesced copy of the staged probabilities (the four softmax warps only: named barrier 1)
    asm volatile("bar.sync 1, 128;" ::: "memory");
    {
      const int tid = threadIdx.x - 32;
      const int nrows = min(128, Nq - mtile * 128);
      const uint32_t nbytes = (uint32_t)nrows * (uint32_t)Nk * 2u;
      uint8_t *gdst = reinterpret_cast<uint8_t *>(p.probs + tile_e0);
      const uint8_t *ssrc = sm + XQ_PL + mis;
      const uint32_t head = (16u - mis) & 15u;                         // bytes up to the first 16-byte boundary
      const uint32_t hb = head < nbytes ? head : nbytes;
      for (uint32_t o = 2u * tid; o < hb; o += 256u) *reinterpret_cast<uint16_t *>(gdst + o) = *reinterpret_cast<const uint16_t *>(ssrc + o);
      const uint32_t body = (nbytes - hb) & ~15u;
      for (uint32_t o = 16u * tid; o < body; o += 2048u)
        *reinterpret_cast<uint4 *>(gdst + hb + o) = *reinterpret_cast<const uint4 *>(ssrc + hb + o);
      for (uint32_t o = hb + body + 2u * tid; o < nbytes; o += 256u)
        *reinterpret_cast<uint16_t *>(gdst + o) = *reinterpret_cast<const uint16_t *>(ssrc + o);
    }
    // epilogue: O (already normalised: P carries 1/l) -> bf16 -> out[row][h*64 .. h*64+63]
    mbar_wait(o_full, 0);
    tc_fence_after();
    uint32_t o0[32], o1[32];
    tmem_ld32(trow, o0);
    tmem_ld32(trow + 32, o1);
    tmem_ld_wait();
    if (row_ok) {
      __nv_bfloat16 *orow = p.out + ((long long)b * Nq + row) * p.ldo + h * 64;
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
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

static int xq_map3(CUtensorMap *m, const void *ptr, long long cols, long long ld, long long rows, long long batches, int box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return KS_EDRIVER;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batches};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)rows * ld * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void *)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? KS_OK : KS_EDRIVER;
}

// Host side.  Returns KS_EUNSUPPORTED when the shape does not fit this kernel (the caller falls back to the mma.sync kernel).
int xattention_fwd_umma(int B, int Nq, int Nk, int heads, const void *q, long long ldq, const void *kv, long long ldkv, float scale,
                        void *out, long long ldo, void *probs, float pdrop, unsigned long long seed, const int *step_ptr, int site,
                        cudaStream_t st) {
  const int inner = heads * 64;
  if (Nk < 1 || Nk > 64 || Nq < 1 || ldq < inner || ldkv < 2 * inner || ldo < inner) return KS_EUNSUPPORTED;
  if ((((uintptr_t)q) % 16) || (((uintptr_t)kv) % 16) || (((uintptr_t)out) % 16) || (((uintptr_t)probs) % 2)) return KS_EUNSUPPORTED;
  if ((ldq * 2) % 16 || (ldkv * 2) % 16 || (ldo * 2) % 16) return KS_EUNSUPPORTED;
  XAttnFwdParams p;
  int rc = xq_map3(&p.q, q, inner, ldq, Nq, B, 128); if (rc) return rc;
  rc = xq_map3(&p.kv, kv, 2LL * inner, ldkv, Nk, B, 64); if (rc) return rc;
  p.out = (__nv_bfloat16 *)out; p.probs = (__nv_bfloat16 *)probs; p.ldo = ldo;
  p.Nq = Nq; p.Nk = Nk; p.heads = heads; p.inner = inner;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.pdrop = pdrop; p.ikeep = 1.f / (1.f - pdrop); p.seed = seed; p.step_ptr = step_ptr; p.site = site;
  p.idesc_s = make_idesc_bf16(128, 64, 0, 0);
  p.idesc_o = make_idesc_bf16(128, 64, 0, 1);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(xattention_fwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XQ_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  dim3 grid((unsigned)((Nq + 127) / 128), (unsigned)heads, (unsigned)B);
  xattention_fwd_umma_kernel<<<grid, 160, XQ_SMEM, st>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace ks
