// Attention prologue / epilogue-of-backward: per-head RMSNorm of q and k (MultiheadRMSNorm, vp.py:280-287), half-split
// rotary embedding (vp.py:193-199) and the 'b n (h d) -> b h n d' head split (vp.py:321), in ONE pass over the q and k
// blocks of the to_qkv GEMM output.  cos/sin come from a torch-computed table so the
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

// FOUR lanes own one 64-wide head vector: lane s holds elements [8s, 8s+8) and [32+8s, 32+8s+8) (two 16-byte accesses), so
// the rotary partner of element d (d < 32), d+32, sits in the same thread -- no shuffles for the rotation, two for the norm.
constexpr int kLpv = 4;                     // lanes per vector
constexpr int kVecPerBlock = 256 / kLpv;    // 64

VBX_DEVINL float sum4(float v) {  // over the 4 lanes of one head vector
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
VBX_DEVINL void ld8c(const float* p, float f[8]) {  // small tables (cos/sin/gamma): cached loads
  const float4 a = *reinterpret_cast<const float4*>(p), c = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = c.x; f[5] = c.y; f[6] = c.z; f[7] = c.w;
}

__global__ void __launch_bounds__(256) qkrope_fwd_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ cosv,
                                                          const float* __restrict__ sinv, const float* __restrict__ gq,
                                                          const float* __restrict__ gk, uint16_t* __restrict__ qh,
                                                          uint16_t* __restrict__ kh, int64_t B, int64_t N, int H) {
  const int sub = threadIdx.x & (kLpv - 1);
  const int64_t nvec = B * N * 2 * H;
  const int64_t stride = (int64_t)gridDim.x * kVecPerBlock;  // a multiple of 2H (host): (which, head) fixed per thread
  const int64_t vid0 = (int64_t)blockIdx.x * kVecPerBlock + (threadIdx.x / kLpv);
  VecId id = decode(vid0 < nvec ? vid0 : 0, H);
  const int64_t tok_step = stride / (2 * H);
  int64_t b = id.tok / N, n = id.tok - b * N;
  const float* gam = id.which ? gk : gq;
  float glo[8], ghi[8];
  if (gam != nullptr) {
    ld8c(gam + id.h * kDh + sub * 8, glo);
    ld8c(gam + id.h * kDh + 32 + sub * 8, ghi);
  }
  for (int64_t vid = vid0;; vid += stride) {
    const bool active = vid < nvec;  // keep all lanes in the shuffles
    if (!active) { id.tok = 0; b = 0; n = 0; }
    const uint16_t* src = qkv + id.tok * (3 * H * kDh) + id.which * (H * kDh) + id.h * kDh + sub * 8;
    float lo[8], hi[8], cs[8], sn[8];
    unpack8(ldg_nc_16(src), lo);
    unpack8(ldg_nc_16(src + 32), hi);
    ld8c(cosv + n * 32 + sub * 8, cs);
    ld8c(sinv + n * 32 + sub * 8, sn);
    if (gam != nullptr) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) ss = fmaf(lo[i], lo[i], fmaf(hi[i], hi[i], ss));
      ss = sum4(ss);
      const float sc = 8.0f * fminf(rsqrtf(ss), 1e12f);  // F.normalize (x / max(||x||, 1e-12)) * sqrt(64); MUFU.RSQ, rel err 2^-22
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        lo[i] *= sc * glo[i];
        hi[i] *= sc * ghi[i];
      }
    }
    float olo[8], ohi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {  // rotate_half([a,b]) = [-b, a]
      olo[i] = fmaf(lo[i], cs[i], -hi[i] * sn[i]);
      ohi[i] = fmaf(hi[i], cs[i], lo[i] * sn[i]);
    }
    if (active) {
      uint16_t* dst = (id.which ? kh : qh) + ((b * H + id.h) * N + n) * kDh + sub * 8;
      stg_16(dst, pack8(olo));
      stg_16(dst + 32, pack8(ohi));
    }
    if (__all_sync(0xffffffffu, vid + stride >= nvec)) break;
    id.tok += tok_step;  // advance (b, n) without dividing
    n += tok_step;
    while (n >= N) { n -= N; ++b; }
  }
}

// backward: dy (f32 for q, bf16 for k) -> d qkv[q|k blocks] (bf16), dgamma_q / dgamma_k accumulated in registers (each thread
// keeps the same (which, head, slice) for its whole loop) and flushed with one atomic per element per thread.
__global__ void __launch_bounds__(256) qkrope_bwd_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ cosv,
                                                          const float* __restrict__ sinv, const float* __restrict__ gq,
                                                          const float* __restrict__ gk, const float* __restrict__ dqh,
                                                          const uint16_t* __restrict__ dkh, uint16_t* __restrict__ dqkv,
                                                          float* __restrict__ dgq, float* __restrict__ dgk, int64_t B,
                                                          int64_t N, int H) {
  const int sub = threadIdx.x & (kLpv - 1);
  const int64_t nvec = B * N * 2 * H;
  const int64_t stride = (int64_t)gridDim.x * kVecPerBlock;
  const int64_t vid0 = (int64_t)blockIdx.x * kVecPerBlock + (threadIdx.x / kLpv);
  const VecId id0 = decode(vid0 < nvec ? vid0 : 0, H);
  const float* gam = id0.which ? gk : gq;
  float glo[8], ghi[8], dglo[8], dghi[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) glo[i] = ghi[i] = 1.f, dglo[i] = dghi[i] = 0.f;
  if (gam != nullptr) {
    ld8c(gam + id0.h * kDh + sub * 8, glo);
    ld8c(gam + id0.h * kDh + 32 + sub * 8, ghi);
  }
  VecId id = id0;
  const int64_t tok_step = stride / (2 * H);
  int64_t b = id.tok / N, n = id.tok - b * N;
  for (int64_t vid = vid0;; vid += stride) {
    const bool active = vid < nvec;
    if (!active) { id.tok = 0; b = 0; n = 0; }
    const int64_t hoff = ((b * H + id.h) * N + n) * kDh + sub * 8;
    float dlo[8], dhi[8], xlo[8], xhi[8], cs[8], sn[8];
    if (id.which) {
      unpack8(ldg_nc_16(dkh + hoff), dlo);
      unpack8(ldg_nc_16(dkh + hoff + 32), dhi);
    } else {
      ld8f(dqh + hoff, dlo);
      ld8f(dqh + hoff + 32, dhi);
    }
    const int64_t goff = id.tok * (3 * H * kDh) + id.which * (H * kDh) + id.h * kDh + sub * 8;
    unpack8(ldg_nc_16(qkv + goff), xlo);
    unpack8(ldg_nc_16(qkv + goff + 32), xhi);
    ld8c(cosv + n * 32 + sub * 8, cs);
    ld8c(sinv + n * 32 + sub * 8, sn);
    // undo the rotation (transpose): dz_lo = dy_lo c + dy_hi s ; dz_hi = dy_hi c - dy_lo s
    float zlo[8], zhi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      zlo[i] = fmaf(dlo[i], cs[i], dhi[i] * sn[i]);
      zhi[i] = fmaf(dhi[i], cs[i], -dlo[i] * sn[i]);
    }
    if (gam != nullptr) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) ss = fmaf(xlo[i], xlo[i], fmaf(xhi[i], xhi[i], ss));
      ss = sum4(ss);
      const float rinv = fminf(rsqrtf(ss), 1e12f);
      const float s1 = 8.0f * rinv;
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (active) {
          dglo[i] = fmaf(zlo[i] * xlo[i], s1, dglo[i]);  // dgamma += dz * xhat * 8
          dghi[i] = fmaf(zhi[i] * xhi[i], s1, dghi[i]);
        }
        zlo[i] *= glo[i];
        zhi[i] *= ghi[i];
        dot = fmaf(zlo[i], xlo[i], fmaf(zhi[i], xhi[i], dot));
      }
      dot = sum4(dot);
      const float s2 = s1 * rinv * rinv * dot;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        zlo[i] = fmaf(zlo[i], s1, -xlo[i] * s2);
        zhi[i] = fmaf(zhi[i], s1, -xhi[i] * s2);
      }
    }
    if (active) {
      stg_16(dqkv + goff, pack8(zlo));
      stg_16(dqkv + goff + 32, pack8(zhi));
    }
    if (__all_sync(0xffffffffu, vid + stride >= nvec)) break;
    id.tok += tok_step;
    n += tok_step;
    while (n >= N) { n -= N; ++b; }
  }
  if (gam != nullptr && vid0 < nvec) {
    float* dg = (id0.which ? dgk : dgq) + id0.h * kDh + sub * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(dg + i, dglo[i]);
      atomicAdd(dg + 32 + i, dghi[i]);
    }
  }
}

}  // namespace vbx

using namespace vbx;

// grid stride (grid*64 vectors) must be a multiple of 2H: round the grid to a multiple of m = 2H / gcd(64, 2H)
static int64_t rope_grid(int64_t nvec, int64_t H, int blocks_per_sm) {
  int64_t a = kVecPerBlock, c = 2 * H;
  while (c) { int64_t t = a % c; a = c; c = t; }
  const int64_t m = (2 * H) / a;
  int64_t grid = grid_for(nvec, kVecPerBlock, blocks_per_sm);
  return ((grid + m - 1) / m) * m;
}

extern "C" int vbx_qkrope_fwd(const uint16_t* qkv, const float* cosv, const float* sinv, const float* gq, const float* gk,
                              uint16_t* qh, uint16_t* kh, int64_t B, int64_t N, int64_t H, void* stream) {
  VBX_REQUIRE(qkv && cosv && sinv && qh && kh, VBX_E_NULL);
  VBX_REQUIRE((gq == nullptr) == (gk == nullptr), VBX_E_NULL);
  VBX_REQUIRE(B > 0 && N > 0 && H > 0, VBX_E_SHAPE);
  VBX_REQUIRE(VBX_ALIGNED16(qkv) && VBX_ALIGNED16(qh) && VBX_ALIGNED16(kh), VBX_E_ALIGN);
  const int64_t nvec = B * N * 2 * H;
  qkrope_fwd_kernel<<<(unsigned)rope_grid(nvec, H, 8), 256, 0, (cudaStream_t)stream>>>(qkv, cosv, sinv, gq, gk, qh, kh, B, N, (int)H);
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
  // 2 resident blocks per SM (122 registers; forcing 3 spills and is slower): more blocks would only multiply the dgamma atomics
  qkrope_bwd_kernel<<<(unsigned)rope_grid(nvec, H, 2), 256, 0, (cudaStream_t)stream>>>(qkv, cosv, sinv, gq, gk, dqh, dkh, dqkv, dgq, dgk, B, N,
                                                                      (int)H);
  return VBX_LAUNCH_RC();
}
