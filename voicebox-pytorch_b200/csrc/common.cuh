// Shared device helpers for libvbx_sm100a (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vbx.h"

#define VBX_DEVINL __device__ __forceinline__

namespace vbx {

constexpr int kNumSM = 148;  // B200: 2 dies x 74 SMs

struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};

VBX_DEVINL float2 bf2f(__nv_bfloat162 v) { return __bfloat1622float2(v); }
VBX_DEVINL __nv_bfloat162 f2bf(float a, float b) { return __floats2bfloat162_rn(a, b); }

// 16-byte streaming loads/stores: these tensors are touched once per kernel, keep them out of L1.
VBX_DEVINL uint4 ldg_nc_16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// same, without .nc: for buffers the same kernel may also write (in-place residual stream)
VBX_DEVINL uint4 ldg_16(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
VBX_DEVINL void stg_16(void* p, uint4 v) {
  // no "memory" clobber: volatile asm statements stay ordered among themselves, while the compiler remains free to hoist
  // independent plain loads (gamma/beta tables) above these streaming stores
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}

VBX_DEVINL void unpack8(uint4 u, float f[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = bf2f(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
VBX_DEVINL uint4 pack8(const float f[8]) {
  uint4 u;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = f2bf(f[2 * i], f[2 * i + 1]);
  return u;
}
// 8 consecutive floats = ONE 32-byte sector.  sm_100 has 256-bit global accesses (LDG.256 / STG.256): one request per sector
// instead of two 16-byte requests that each touch half of it (the half-sector pattern doubled the L2 <-> SM sector traffic of
// every fp32 stream in round 1).  Pointers that are only 16-byte aligned fall back to the two-request form.
VBX_DEVINL void ld8f(const float* p, float f[8]) {
  uint32_t r[8];
  if ((reinterpret_cast<uintptr_t>(p) & 31) == 0) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
  } else {
    const uint4 a = ldg_nc_16(p), b = ldg_nc_16(p + 4);
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[i]);
}
VBX_DEVINL void ld8f_rw(const float* p, float f[8]) {   // without .nc: the same kernel may also write the buffer
  uint32_t r[8];
  if ((reinterpret_cast<uintptr_t>(p) & 31) == 0) {
    asm volatile("ld.global.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
  } else {
    const uint4 a = ldg_16(p), b = ldg_16(p + 4);
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[i]);
}
VBX_DEVINL void st8f(float* p, const float f[8]) {
  if ((reinterpret_cast<uintptr_t>(p) & 31) == 0) {
    asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(__float_as_uint(f[0])),
                 "r"(__float_as_uint(f[1])), "r"(__float_as_uint(f[2])), "r"(__float_as_uint(f[3])), "r"(__float_as_uint(f[4])),
                 "r"(__float_as_uint(f[5])), "r"(__float_as_uint(f[6])), "r"(__float_as_uint(f[7])));
  } else {
    stg_16(p, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
    stg_16(p + 4, make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7])));
  }
}

// 4-element (8 / 16 byte) accesses
VBX_DEVINL uint2 ldg_nc_8(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
VBX_DEVINL void stg_8(void* p, uint2 v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y));
}
VBX_DEVINL void ld4f(const float* p, float* f, bool rw) {   // rw: the buffer may be written by this kernel (no .nc)
  const uint4 a = rw ? ldg_16(p) : ldg_nc_16(p);
  f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
}
VBX_DEVINL void st4f(float* p, const float* f) {
  stg_16(p, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
}
VBX_DEVINL void ld4h(const uint16_t* p, float* f) {
  const uint2 u = ldg_nc_8(p);
  const float2 a = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&u.x)), b = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
VBX_DEVINL void st4h(uint16_t* p, const float* f) {
  __nv_bfloat162 a = f2bf(f[0], f[1]), b = f2bf(f[2], f[3]);
  stg_8(p, make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b)));
}

VBX_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7): 1 MUFU.EX2 + 1 MUFU.RCP + 7 FMA.  The exact-erf GELU
// of the reference (nn.GELU(), F.gelu default) is reproduced to ~2e-7 absolute, far inside the bf16 output ulp, at a
// third of erff()'s instruction count, which keeps the GEGLU / conv passes HBM-bound instead of issue-bound.
// Returns Phi(x) = 0.5*(1+erf(x/sqrt2)) and e = exp(-x^2/2).
VBX_DEVINL float rcp_approx(float x) {  // MUFU.RCP, rel err 2^-23; __frcp_rn would add Newton steps + a slow-path CALL
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
VBX_DEVINL float ex2_approx(float x) {  // MUFU.EX2, rel err 2^-22; exp2f() adds denormal range handling
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Layout of the normed + rotated q / k and of their gradients between the rope kernels and the attention kernels.
// 1 (shipped): token-major [B, N, H, 64] -- the rope kernels then read and write whole-token contiguous runs (2-4 KB) and the
// attention kernels address tiles through their tensor maps' strides, exactly as they always did for V (read in place from the
// qkv GEMM output) and dO.  0: head-major [B, H, N, 64] (rounds 1-2 until the last change; one 128-256 byte piece per (token, head)
// on the rope side: 4.2-4.4 TB/s at best, profiles/r2_row_kernels_investigation.md).
#ifndef VBX_QK_TOKEN_MAJOR
#define VBX_QK_TOKEN_MAJOR 1
#endif
// element offset of head vector (b, h, n) in such a tensor
VBX_DEVINL int64_t qk_vec_off(int64_t b, int64_t h, int64_t n, int64_t H, int64_t N) {
#if VBX_QK_TOKEN_MAJOR
  return ((b * N + n) * H + h) * 64;
#else
  return ((b * H + h) * N + n) * 64;
#endif
}
// L2 prefetch of a contiguous, 16-byte aligned range (one instruction, no registers held): the row kernels are latency bound
// at 14-32 resident warps per SM (ncu: every stall is long-scoreboard, DRAM at 3.3-4.7 TB/s), so each warp announces the NEXT
// row it will touch while it works on the current one and its demand loads then hit L2.  VBX_ROW_PREFETCH=0 compiles it out.
#ifndef VBX_ROW_PREFETCH
#define VBX_ROW_PREFETCH 1
#endif
VBX_DEVINL void prefetch_l2_bulk(const void* p, uint32_t bytes) {
#if VBX_ROW_PREFETCH
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
#endif
}
VBX_DEVINL float normal_cdf(float x, float& e) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  e = ex2_approx(-1.4426950408889634f * z * z);
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float half_erfc = 0.5f * p * t * e;  // 0.5*erfc(z)
  return x >= 0.f ? 1.0f - half_erfc : half_erfc;
}
VBX_DEVINL float gelu_f(float x) {
  float e;
  return x * normal_cdf(x, e);
}
// d/dx gelu(x) = Phi(x) + x * phi(x)
VBX_DEVINL float gelu_grad_f(float x) {
  float e;
  const float c = normal_cdf(x, e);
  return fmaf(x * 0.3989422804014327f, e, c);
}

inline int grid_for(int64_t work_items, int per_block, int max_blocks_per_sm = 8) {
  int64_t g = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)kNumSM * max_blocks_per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace vbx

#define VBX_REQUIRE(cond, code) \
  do {                          \
    if (!(cond)) return (code); \
  } while (0)
#define VBX_ALIGNED16(p) ((reinterpret_cast<uintptr_t>(p) & 15) == 0)
#define VBX_LAUNCH_RC() ((int)cudaGetLastError())
