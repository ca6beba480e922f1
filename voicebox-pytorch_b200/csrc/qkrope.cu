// Attention prologue / epilogue-of-backward: per-head RMSNorm of q and k (MultiheadRMSNorm, vp.py:280-287), half-split
// rotary embedding (vp.py:193-199) and the 'b n (h d) -> b h n d' head split (vp.py:321), in ONE pass over the q and k
// blocks of the to_qkv GEMM output.  8 lanes own one 64-wide head vector.  cos/sin come from a torch-computed table so the
// -10000 register-token position (vp.py:440) gets a correctly range-reduced angle.
#include "common.cuh"

namespace vbx {

constexpr int kDh = 64;

// vector id -> (which: 0 = q, 1 = k ; token ; head).  Heads fastest, so a warp reads 512 contiguous bytes.
struct VecId {
  int64_t tok;
  int which, h;
};
VBX_DEVINL VecId decode(int64_t vid, int H) {
  VecId v;
  const int pair = (int)(vid % (2 * H));
  v.tok = vid / (2 * H);
  v.which = pair / H;
  v.h = pair - v.which * H;
  return v;
}

// Each of the 8 lanes of a head vector owns elements [4s, 4s+4) and [32+4s, 32+4s+4): the rotary partner of element d
// (d < 32) is d+32, held by the same thread, so the rotation needs no shuffles; cos/sin are one float4 each.
VBX_DEVINL void ld4bf(const uint16_t* p, float f[4]) {
  uint2 u;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(u.x), "=r"(u.y) : "l"(p));
  const float2 a = bf2f(*reinterpret_cast<__nv_bfloat162*>(&u.x)), c = bf2f(*reinterpret_cast<__nv_bfloat162*>(&u.y));
  f[0] = a.x; f[1] = a.y; f[2] = c.x; f[3] = c.y;
}
VBX_DEVINL void st4bf(uint16_t* p, const float f[4]) {
  __nv_bfloat162 a = f2bf(f[0], f[1]), c = f2bf(f[2], f[3]);
  uint2 u = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&c));
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(u.x), "r"(u.y) : "memory");
}
VBX_DEVINL float sum8(float v) {  // over the 8 lanes of one head vector
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

__global__ void __launch_bounds__(256) qkrope_fwd_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ cosv,
                                                          const float* __restrict__ sinv, const float* __restrict__ gq,
                                                          const float* __restrict__ gk, uint16_t* __restrict__ qh,
                                                          uint16_t* __restrict__ kh, int64_t B, int64_t N, int H) {
  const int sub = threadIdx.x & 7;
  const int64_t nvec = B * N * 2 * H;
  const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 3);
  for (int64_t vid = (int64_t)blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);; vid += stride) {
    const bool active = vid < nvec;  // keep all lanes in the shuffles
    const VecId id = decode(active ? vid : 0, H);
    const int64_t b = id.tok / N, n = id.tok - b * N;
    const uint16_t* src = qkv + id.tok * (3 * H * kDh) + id.which * (H * kDh) + id.h * kDh + sub * 4;
    float lo[4], hi[4];
    ld4bf(src, lo);
    ld4bf(src + 32, hi);
    const float* gam = id.which ? gk : gq;
    if (gam != nullptr) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) ss = fmaf(lo[i], lo[i], fmaf(hi[i], hi[i], ss));
      ss = sum8(ss);
      const float sc = 8.0f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize * sqrt(64)
      const float4 g0 = *reinterpret_cast<const float4*>(gam + id.h * kDh + sub * 4);
      const float4 g1 = *reinterpret_cast<const float4*>(gam + id.h * kDh + 32 + sub * 4);
      lo[0] *= sc * g0.x; lo[1] *= sc * g0.y; lo[2] *= sc * g0.z; lo[3] *= sc * g0.w;
      hi[0] *= sc * g1.x; hi[1] *= sc * g1.y; hi[2] *= sc * g1.z; hi[3] *= sc * g1.w;
    }
    const float4 c4 = *reinterpret_cast<const float4*>(cosv + n * 32 + sub * 4);
    const float4 s4 = *reinterpret_cast<const float4*>(sinv + n * 32 + sub * 4);
    const float cs[4] = {c4.x, c4.y, c4.z, c4.w}, sn[4] = {s4.x, s4.y, s4.z, s4.w};
    float olo[4], ohi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // rotate_half([a,b]) = [-b, a]
      olo[i] = fmaf(lo[i], cs[i], -hi[i] * sn[i]);
      ohi[i] = fmaf(hi[i], cs[i], lo[i] * sn[i]);
    }
    if (active) {
      uint16_t* dst = (id.which ? kh : qh) + ((b * H + id.h) * N + n) * kDh + sub * 4;
      st4bf(dst, olo);
      st4bf(dst + 32, ohi);
    }
    if (__all_sync(0xffffffffu, vid + stride >= nvec)) break;
  }
}

// backward: dy (f32 for q, bf16 for k) -> d qkv[q|k blocks] (bf16), dgamma_q / dgamma_k accumulated.
// The launch makes the grid stride a multiple of 2H so that each thread keeps the same (which, head, slice) for its
// whole loop and can hold its dgamma partial sums in registers.
__global__ void __launch_bounds__(256) qkrope_bwd_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ cosv,
                                                          const float* __restrict__ sinv, const float* __restrict__ gq,
                                                          const float* __restrict__ gk, const float* __restrict__ dqh,
                                                          const uint16_t* __restrict__ dkh, uint16_t* __restrict__ dqkv,
                                                          float* __restrict__ dgq, float* __restrict__ dgk, int64_t B,
                                                          int64_t N, int H) {
  const int sub = threadIdx.x & 7;
  const int64_t nvec = B * N * 2 * H;
  const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 3);
  const int64_t vid0 = (int64_t)blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);
  const VecId id0 = decode(vid0, H);
  const float* gam = id0.which ? gk : gq;
  float glo[4] = {1.f, 1.f, 1.f, 1.f}, ghi[4] = {1.f, 1.f, 1.f, 1.f}, dglo[4] = {0.f, 0.f, 0.f, 0.f}, dghi[4] = {0.f, 0.f, 0.f, 0.f};
  if (gam != nullptr) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      glo[i] = gam[id0.h * kDh + sub * 4 + i];
      ghi[i] = gam[id0.h * kDh + 32 + sub * 4 + i];
    }
  }
  for (int64_t vid = vid0;; vid += stride) {
    const bool active = vid < nvec;
    const VecId id = decode(active ? vid : 0, H);
    const int64_t b = id.tok / N, n = id.tok - b * N;
    const int64_t hoff = ((b * H + id.h) * N + n) * kDh + sub * 4;
    float dlo[4], dhi[4], xlo[4], xhi[4];
    if (id.which) {
      ld4bf(dkh + hoff, dlo);
      ld4bf(dkh + hoff + 32, dhi);
    } else {
      const float4 a = *reinterpret_cast<const float4*>(dqh + hoff), c = *reinterpret_cast<const float4*>(dqh + hoff + 32);
      dlo[0] = a.x; dlo[1] = a.y; dlo[2] = a.z; dlo[3] = a.w;
      dhi[0] = c.x; dhi[1] = c.y; dhi[2] = c.z; dhi[3] = c.w;
    }
    const int64_t goff = id.tok * (3 * H * kDh) + id.which * (H * kDh) + id.h * kDh + sub * 4;
    ld4bf(qkv + goff, xlo);
    ld4bf(qkv + goff + 32, xhi);
    const float4 c4 = *reinterpret_cast<const float4*>(cosv + n * 32 + sub * 4);
    const float4 s4 = *reinterpret_cast<const float4*>(sinv + n * 32 + sub * 4);
    const float cs[4] = {c4.x, c4.y, c4.z, c4.w}, sn[4] = {s4.x, s4.y, s4.z, s4.w};
    // undo the rotation (transpose): dz_lo = dy_lo c + dy_hi s ; dz_hi = dy_hi c - dy_lo s
    float zlo[4], zhi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      zlo[i] = fmaf(dlo[i], cs[i], dhi[i] * sn[i]);
      zhi[i] = fmaf(dhi[i], cs[i], -dlo[i] * sn[i]);
    }
    float olo[4], ohi[4];
    if (gam != nullptr) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) ss = fmaf(xlo[i], xlo[i], fmaf(xhi[i], xhi[i], ss));
      ss = sum8(ss);
      const float rinv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
      const float s1 = 8.0f * rinv;
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (active) {
          dglo[i] = fmaf(zlo[i] * xlo[i], s1, dglo[i]);  // dgamma += dz * xhat * 8
          dghi[i] = fmaf(zhi[i] * xhi[i], s1, dghi[i]);
        }
        zlo[i] *= glo[i];
        zhi[i] *= ghi[i];
        dot = fmaf(zlo[i], xlo[i], fmaf(zhi[i], xhi[i], dot));
      }
      dot = sum8(dot);
      const float s2 = s1 * rinv * rinv * dot;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        olo[i] = fmaf(zlo[i], s1, -xlo[i] * s2);
        ohi[i] = fmaf(zhi[i], s1, -xhi[i] * s2);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) olo[i] = zlo[i], ohi[i] = zhi[i];
    }
    if (active) {
      st4bf(dqkv + goff, olo);
      st4bf(dqkv + goff + 32, ohi);
    }
    if (__all_sync(0xffffffffu, vid + stride >= nvec)) break;
  }
  if (gam != nullptr && vid0 < nvec) {
    float* dg = (id0.which ? dgk : dgq) + id0.h * kDh + sub * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(dg + i, dglo[i]);
      atomicAdd(dg + 32 + i, dghi[i]);
    }
  }
}

}  // namespace vbx

using namespace vbx;

extern "C" int vbx_qkrope_fwd(const uint16_t* qkv, const float* cosv, const float* sinv, const float* gq, const float* gk,
                              uint16_t* qh, uint16_t* kh, int64_t B, int64_t N, int64_t H, void* stream) {
  VBX_REQUIRE(qkv && cosv && sinv && qh && kh, VBX_E_NULL);
  VBX_REQUIRE((gq == nullptr) == (gk == nullptr), VBX_E_NULL);
  VBX_REQUIRE(B > 0 && N > 0 && H > 0, VBX_E_SHAPE);
  VBX_REQUIRE(VBX_ALIGNED16(qkv) && VBX_ALIGNED16(qh) && VBX_ALIGNED16(kh), VBX_E_ALIGN);
  const int64_t nvec = B * N * 2 * H;
  qkrope_fwd_kernel<<<grid_for(nvec, 32, 8), 256, 0, (cudaStream_t)stream>>>(qkv, cosv, sinv, gq, gk, qh, kh, B, N, (int)H);
  return VBX_LAUNCH_RC();
}

extern "C" int vbx_qkrope_bwd(const uint16_t* qkv, const float* cosv, const float* sinv, const float* gq, const float* gk,
                              const float* dqh, const uint16_t* dkh, uint16_t* dqkv, float* dgq, float* dgk, int64_t B,
                              int64_t N, int64_t H, void* stream) {
  VBX_REQUIRE(qkv && cosv && sinv && dqh && dkh && dqkv, VBX_E_NULL);
  VBX_REQUIRE((gq == nullptr) == (gk == nullptr), VBX_E_NULL);
  VBX_REQUIRE(gq == nullptr || (dgq && dgk), VBX_E_NULL);
  VBX_REQUIRE(B > 0 && N > 0 && H > 0, VBX_E_SHAPE);
  VBX_REQUIRE(VBX_ALIGNED16(qkv) && VBX_ALIGNED16(dqh) && VBX_ALIGNED16(dkh) && VBX_ALIGNED16(dqkv), VBX_E_ALIGN);
  const int64_t nvec = B * N * 2 * H;
  // grid stride (grid*32 vectors) must be a multiple of 2H: round the grid to a multiple of m = 2H / gcd(32, 2H)
  int64_t a = 32, c = 2 * H;
  while (c) { int64_t t = a % c; a = c; c = t; }
  const int64_t m = (2 * H) / a;
  int64_t grid = grid_for(nvec, 32, 8);
  grid = ((grid + m - 1) / m) * m;
  qkrope_bwd_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(qkv, cosv, sinv, gq, gk, dqh, dkh, dqkv, dgq, dgk, B, N,
                                                                      (int)H);
  return VBX_LAUNCH_RC();
}
