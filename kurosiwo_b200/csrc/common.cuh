// Shared device/host helpers for libkurosiwo_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/kurosiwo_b200.h"

#define KS_CHECK_ARG(cond) do { if (!(cond)) return KS_EINVAL; } while (0)
#define KS_LAUNCH_RET() do { cudaError_t e__ = cudaGetLastError(); return (int)e__; } while (0)

namespace ks {

constexpr int kNumSMs = 148;

struct Options { int mt, bo_mode, tc_disable, wgrad_tc_disable, sa, sb, v1, no_resident, debug, tb, wgrad_mode, att_simt, ew_cap, no_ns3, ns3_min_cin, ns3_mode, nacc, loss_chunks, loss_no_pdl, loss_no_bulk, loss_variant, ew, att_no_umma, cs_rows, ln_rows, stem_simt, ecam_simt, xatt_umma, stat_mode, dwconv_simple, cf_scalar; };
extern Options g_opt;

template <typename T> struct Cvt;
template <> struct Cvt<float> {
  static __device__ __forceinline__ float ld(const float *p) { return *p; }
  static __device__ __forceinline__ void st(float *p, float v) { *p = v; }
};
template <> struct Cvt<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }
};

template <typename T> __device__ __forceinline__ float round_as(float v);
template <> __device__ __forceinline__ float round_as<float>(float v) { return v; }
template <> __device__ __forceinline__ float round_as<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// Device-side copy of a view with typed pointer.
struct View {
  char *ptr;
  long long sn, sh, sw;
  int C;
};

static inline View to_view(const ks_view_t &v) {
  View r; r.ptr = (char *)v.ptr; r.sn = v.sn; r.sh = v.sh; r.sw = v.sw; r.C = v.C; return r;
}

struct ViewList {
  View v[KS_MAX_VIEWS];
  int cstart[KS_MAX_VIEWS + 1];  // cumulative channel offsets
  int n;
};

static inline int make_view_list(const ks_view_t *vs, int n, ViewList &out) {
  if (n < 1 || n > KS_MAX_VIEWS || vs == nullptr) return KS_EINVAL;
  out.n = n; out.cstart[0] = 0;
  for (int i = 0; i < n; ++i) {
    if (vs[i].ptr == nullptr || vs[i].C <= 0) return KS_EINVAL;
    out.v[i] = to_view(vs[i]);
    out.cstart[i + 1] = out.cstart[i] + vs[i].C;
  }
  for (int i = n; i < KS_MAX_VIEWS; ++i) { out.v[i] = out.v[0]; out.cstart[i + 1] = out.cstart[n]; }
  return KS_OK;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 8 bf16 <-> 8 floats via one 16-byte access
__device__ __forceinline__ void ld8(const __nv_bfloat16 *p, float (&f)[8]) {
  uint4 u = *reinterpret_cast<const uint4 *>(p);
  const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ void st8(__nv_bfloat16 *p, const float (&f)[8]) {
  uint4 u; __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4 *>(p) = u;
}
__device__ __forceinline__ void ld8(const float *p, float (&f)[8]) {
  float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void st8(float *p, const float (&f)[8]) {
  *reinterpret_cast<float4 *>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4 *>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

}  // namespace ks
