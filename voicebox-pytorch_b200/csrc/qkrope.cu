// Attention prologue / epilogue-of-backward: per-head RMSNorm of q and k (MultiheadRMSNorm, vp.py:280-287), half-split
// rotary embedding (vp.py:193-199) and the 'b n (h d) -> b h n d' head split (vp.py:321), in ONE pass over the q and k
// blocks of the to_qkv GEMM output.  8 lanes own one 64-wide head vector (8 elements = 16 B each); the rotary partner of
// element d is d +/- 32, i.e. the same register slot of lane ^ 4.  cos/sin come from a torch-computed table so the
// -10000 register-token position (vp.py:440) gets a correctly range-reduced angle.
#include "common.cuh"

namespace vbx {

constexpr int kDh = 64;

// vector id -> (which: 0 = q, 1 = k ; token ; head).  Heads fastest, so a warp reads 512 contiguous bytes.
struct VecId {
  int64_t tok;
  int which, h;
};
VBX_DEVINL void ld8_cached(const float* p, float f[8]) {  // small tables: let them live in L1
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
VBX_DEVINL VecId decode(int64_t vid, int H) {
  VecId v;
  const int pair = (int)(vid % (2 * H));
  v.tok = vid / (2 * H);
  v.which = pair / H;
  v.h = pair - v.which * H;
  return v;
}

__global__ void __launch_bounds__(256) qkrope_fwd_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ cosv,
                                                          const float* __restrict__ sinv, const float* __restrict__ gq,
                                                          const float* __restrict__ gk, uint16_t* __restrict__ qh,
                                                          uint16_t* __restrict__ kh, int64_t B, int64_t N, int H) {
  const int sub = threadIdx.x & 7;  // which 8-element slice of the head vector
  const int64_t nvec = B * N * 2 * H;
  const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 3);
  for (int64_t vid = (int64_t)blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);; vid += stride) {
    const bool active = vid < nvec;                        // keep all lanes in the shuffles
    const VecId id = decode(active ? vid : 0, H);
    const int64_t b = id.tok / N, n = id.tok - b * N;
    float v[8];
    unpack8(ldg_nc_16(qkv + id.tok * (3 * H * kDh) + id.which * (H * kDh) + id.h * kDh + sub * 8), v);
    const float* gam = id.which ? gk : gq;
    if (gam != nullptr) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) ss = fmaf(v[i], v[i], ss);
      ss += __shfl_xor_sync(0xffffffffu, ss, 1);
      ss += __shfl_xor_sync(0xffffffffu, ss, 2);
      ss += __shfl_xor_sync(0xffffffffu, ss, 4);
      const float sc = 8.0f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize * sqrt(64)
      float gl[8];
      ld8_cached(gam + id.h * kDh + sub * 8, gl);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = v[i] * sc * gl[i];
    }
    float o[8], cs[8], sn[8];
    const int fi = (sub & 3) * 8;  // frequency index base (d mod 32)
    ld8_cached(cosv + n * 32 + fi, cs);
    ld8_cached(sinv + n * 32 + fi, sn);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float partner = __shfl_xor_sync(0xffffffffu, v[i], 4);
      const float rot = (sub < 4) ? -partner : partner;  // rotate_half([a,b]) = [-b, a]
      o[i] = fmaf(v[i], cs[i], rot * sn[i]);
    }
    if (active) {
      uint16_t* dst = (id.which ? kh : qh) + ((b * H + id.h) * N + n) * kDh + sub * 8;
      stg_16(dst, pack8(o));
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
  float gl[8], dgl[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    gl[i] = gam ? gam[id0.h * kDh + sub * 8 + i] : 1.f;
    dgl[i] = 0.f;
  }
  for (int64_t vid = vid0;; vid += stride) {
    const bool active = vid < nvec;
    const VecId id = decode(active ? vid : 0, H);
    const int64_t b = id.tok / N, n = id.tok - b * N;
    const int64_t hoff = ((b * H + id.h) * N + n) * kDh + sub * 8;
    float dy[8], x[8];
    if (id.which) unpack8(ldg_nc_16(dkh + hoff), dy);
    else ld8f(dqh + hoff, dy);
    const int64_t goff = id.tok * (3 * H * kDh) + id.which * (H * kDh) + id.h * kDh + sub * 8;
    unpack8(ldg_nc_16(qkv + goff), x);
    // undo the rotation: dz[d] = dy[d] cos + (d<32 ? dy[d+32] : -dy[d-32]) sin
    float dz[8], cs[8], sn[8];
    const int fi = (sub & 3) * 8;
    ld8_cached(cosv + n * 32 + fi, cs);
    ld8_cached(sinv + n * 32 + fi, sn);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float partner = __shfl_xor_sync(0xffffffffu, dy[i], 4);
      dz[i] = fmaf(dy[i], cs[i], ((sub < 4) ? partner : -partner) * sn[i]);
    }
    float o[8];
    if (gam != nullptr) {
      float ss = 0.f, dot = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) ss = fmaf(x[i], x[i], ss);
      ss += __shfl_xor_sync(0xffffffffu, ss, 1);
      ss += __shfl_xor_sync(0xffffffffu, ss, 2);
      ss += __shfl_xor_sync(0xffffffffu, ss, 4);
      const float rinv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
      const float s1 = 8.0f * rinv;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (active) dgl[i] = fmaf(dz[i] * x[i], s1, dgl[i]);  // dgamma += dz * xhat * 8
        dz[i] *= gl[i];
        dot = fmaf(dz[i], x[i], dot);
      }
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      dot += __shfl_xor_sync(0xffffffffu, dot, 4);
      const float s2 = s1 * rinv * rinv * dot;
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(dz[i], s1, -x[i] * s2);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = dz[i];
    }
    if (active) stg_16(dqkv + goff, pack8(o));
    if (__all_sync(0xffffffffu, vid + stride >= nvec)) break;
  }
  if (gam != nullptr && vid0 < nvec) {
    float* dg = (id0.which ? dgk : dgq) + id0.h * kDh + sub * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(dg + i, dgl[i]);
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
