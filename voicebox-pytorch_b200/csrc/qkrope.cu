// Attention prologue / epilogue-of-backward: per-head RMSNorm of q and k (MultiheadRMSNorm, vp.py:280-287) and half-split
// rotary embedding (vp.py:193-199), in ONE pass over the q and k blocks of the to_qkv GEMM output.  The 'b n (h d) -> b h n d'
// head split (vp.py:321) costs nothing: q^ / k^ stay token-major [B, N, H, 64] and the attention kernels pick a head's tile through
// their tensor maps' strides (common.cuh: VBX_QK_TOKEN_MAJOR).  cos/sin come from a torch-computed table so the
// -10000 register-token position (vp.py:440) gets a correctly range-reduced angle.
#include "umma.cuh"

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

// Token order.  The q/k blocks of qkv are token-major, q_h / k_h / dq_h / dk_h head-major ([b, h, n, 64]): whichever way a
// block walks, one side is made of 128-256 byte pieces.  GROUPED order (used whenever the block's vector slots divide into whole
// tokens, i.e. VPB % 2H == 0): a block takes kTokGroup CONSECUTIVE tokens x all (which, head) pairs per outer step, so within a
// few microseconds it touches kTokGroup x 4 KB contiguous token-major bytes and, per head, kTokGroup consecutive 128-256 byte
// pieces = 1-2 KB contiguous head-major bytes (DRAM page and L2 write-combining locality; the strided order reached only
// 3.3-3.4 TB/s in the backward with every stall a DRAM wait, whatever the occupancy).  Otherwise the round-1 order: a thread
// group keeps (which, head) and jumps tok_step = grid * VPB / 2H tokens per iteration (grid rounded so that 2H | grid * VPB).
constexpr int kTokGroup = 8;
struct TokWalk {
  int which, h;          // fixed per thread
  int64_t tok0;          // legacy: first token; grouped: token offset inside a group
  int64_t step;          // tokens per inner iteration
  int passes;            // grouped: inner iterations per group
  int group;             // grouped: tokens per group (>= kTokGroup)
};
VBX_DEVINL TokWalk make_walk(int vslot, int vpb, int H, bool grouped) {
  TokWalk w;
  if (grouped) {
    const int pair = vslot % (2 * H);
    w.which = pair / H;
    w.h = pair - w.which * H;
    w.tok0 = vslot / (2 * H);
    w.step = vpb / (2 * H);
    w.group = w.step > kTokGroup ? (int)w.step : kTokGroup;
    w.passes = w.group / (int)w.step;
  } else {
    const int64_t vid0 = (int64_t)blockIdx.x * vpb + vslot;
    const int pair = (int)(vid0 % (2 * H));
    w.which = pair / H;
    w.h = pair - w.which * H;
    w.tok0 = vid0 / (2 * H);
    w.step = (int64_t)gridDim.x * vpb / (2 * H);
    w.group = 0;
    w.passes = 0x7fffffff;
  }
  return w;
}

// U vectors per thread in flight (see the backward below: bytes outstanding per SM, not occupancy, set these kernels' speed)
#ifndef VBX_QKROPE_FWD_U
#define VBX_QKROPE_FWD_U 4
#endif
#ifndef VBX_QKROPE_FWD_MINB
#define VBX_QKROPE_FWD_MINB 2
#endif
template <int U>
__global__ void __launch_bounds__(256, VBX_QKROPE_FWD_MINB) qkrope_fwd_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ cosv,
                                                          const float* __restrict__ sinv, const float* __restrict__ gq,
                                                          const float* __restrict__ gk, uint16_t* __restrict__ qh,
                                                          uint16_t* __restrict__ kh, int64_t B, int64_t N, int H, int grouped) {
  const int sub = threadIdx.x & (kLpv - 1);
  const int64_t total = B * N;
  const TokWalk w = make_walk(threadIdx.x / kLpv, kVecPerBlock, H, grouped != 0);
  const float* gam = w.which ? gk : gq;
  float glo[8], ghi[8];
  if (gam != nullptr) {
    ld8c(gam + w.h * kDh + sub * 8, glo);
    ld8c(gam + w.h * kDh + 32 + sub * 8, ghi);
  }
  for (int64_t g = blockIdx.x;; g += gridDim.x) {
    int64_t tok = grouped ? g * w.group + w.tok0 : w.tok0;
    if (grouped && g * w.group >= total) break;
    int64_t b = tok / N, n = tok - b * N;
    for (int ps = 0; ps < w.passes; ps += U) {
      uint4 xl[U], xh[U];
      const int64_t tok_c = tok, b_c = b, n_c = n;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t t_ = tok < total ? tok : 0;   // inactive lanes read token 0 and stay in the shuffles
        const uint16_t* src = qkv + t_ * (3 * H * kDh) + w.which * (H * kDh) + w.h * kDh + sub * 8;
        xl[u] = ldg_nc_16(src);
        xh[u] = ldg_nc_16(src + 32);
        tok += w.step;
        n += w.step;
        while (n >= N) { n -= N; ++b; }
      }
      int64_t tok2 = tok_c, b2 = b_c, n2 = n_c;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool active = tok2 < total;
        const int64_t b_ = active ? b2 : 0, n_ = active ? n2 : 0;
        tok2 += w.step;
        n2 += w.step;
        while (n2 >= N) { n2 -= N; ++b2; }
        float lo[8], hi[8], cs[8], sn[8];
        unpack8(xl[u], lo);
        unpack8(xh[u], hi);
        ld8c(cosv + n_ * 32 + sub * 8, cs);
        ld8c(sinv + n_ * 32 + sub * 8, sn);
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
          uint16_t* dst = (w.which ? kh : qh) + qk_vec_off(b_, w.h, n_, H, N) + sub * 8;
          stg_16(dst, pack8(olo));
          stg_16(dst + 32, pack8(ohi));
        }
      }
      if (!grouped && __all_sync(0xffffffffu, tok >= total)) break;
    }
    if (!grouped) break;
  }
}

// backward: dy (f32 for q, bf16 for k) -> d qkv[q|k blocks] (bf16), dgamma_q / dgamma_k accumulated in registers (each thread
// keeps the same (which, head, slice) for its whole loop) and flushed with one atomic per element per thread.
// EIGHT lanes per head vector here (lane s: elements [4s, 4s+4) and [32+4s, 32+4s+4)): with four lanes the kernel needed 123
// registers (2 blocks = 14 resident warps per SM) and ~300 dependent instructions per vector per thread.
constexpr int kLpvB = 8;                      // lanes per vector, backward
constexpr int kVecPerBlockB = 256 / kLpvB;    // 32
VBX_DEVINL float sum8(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
VBX_DEVINL void ld4c(const float* p, float f[4]) {  // small tables (cos/sin/gamma): cached loads
  const float4 a = *reinterpret_cast<const float4*>(p);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
}

// U tokens per thread are in flight at once: the loads of all U vectors are issued (into packed registers) before the first is
// processed.  Measured: at one vector per thread the backward sat at 3.3 TB/s whatever the occupancy (14 or 32 warps), the token
// order or an L2 prefetch -- a warp-iteration only has 384 bytes outstanding, ~12 KB per SM, a quarter of what 7 TB/s x ~1 us
// needs.
#ifndef VBX_QKROPE_BWD_U
#define VBX_QKROPE_BWD_U 4
#endif
#ifndef VBX_QKROPE_BWD_MINB
#define VBX_QKROPE_BWD_MINB 2
#endif
template <int U>
__global__ void __launch_bounds__(256, VBX_QKROPE_BWD_MINB) qkrope_bwd_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ cosv,
                                                             const float* __restrict__ sinv, const float* __restrict__ gq,
                                                             const float* __restrict__ gk, const float* __restrict__ dqh,
                                                             const uint16_t* __restrict__ dkh, uint16_t* __restrict__ dqkv,
                                                             float* __restrict__ dgq, float* __restrict__ dgk, int64_t B,
                                                             int64_t N, int H, int grouped) {
  const int sub = threadIdx.x & (kLpvB - 1);
  const int64_t total = B * N;
  const TokWalk w = make_walk(threadIdx.x / kLpvB, kVecPerBlockB, H, grouped != 0);
  const float* gam = w.which ? gk : gq;
  float glo[4], ghi[4], dglo[4], dghi[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) glo[i] = ghi[i] = 1.f, dglo[i] = dghi[i] = 0.f;
  if (gam != nullptr) {
    ld4c(gam + w.h * kDh + sub * 4, glo);
    ld4c(gam + w.h * kDh + 32 + sub * 4, ghi);
  }
  bool any = false;
  for (int64_t g = blockIdx.x;; g += gridDim.x) {
    int64_t tok = grouped ? g * w.group + w.tok0 : w.tok0;
    if (grouped && g * w.group >= total) break;
    int64_t b = tok / N, n = tok - b * N;
    for (int ps = 0; ps < w.passes; ps += U) {
      uint4 d0[U], d1[U];          // dy: f32 x 4 (q) or bf16 x 4 in .x/.y (k)
      uint2 x0[U], x1[U];          // pre-norm q / k, bf16 x 4
      const int64_t tok_c = tok;   // the compute phase walks (tok, n) again instead of keeping U offsets in registers
      const int64_t n_c = n;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool act = tok < total;
        any |= act;
        const int64_t t_ = act ? tok : 0, b_ = act ? b : 0, n_ = act ? n : 0;
        const int64_t hoff = qk_vec_off(b_, w.h, n_, H, N) + sub * 4;
        const int64_t goff = t_ * (3 * H * kDh) + w.which * (H * kDh) + w.h * kDh + sub * 4;
        if (w.which) {
          const uint2 a = ldg_nc_8(dkh + hoff), c = ldg_nc_8(dkh + hoff + 32);
          d0[u] = make_uint4(a.x, a.y, 0u, 0u);
          d1[u] = make_uint4(c.x, c.y, 0u, 0u);
        } else {
          d0[u] = ldg_nc_16(dqh + hoff);
          d1[u] = ldg_nc_16(dqh + hoff + 32);
        }
        x0[u] = ldg_nc_8(qkv + goff);
        x1[u] = ldg_nc_8(qkv + goff + 32);
        tok += w.step;
        n += w.step;
        while (n >= N) { n -= N; ++b; }
      }
      int64_t tok2 = tok_c, n2 = n_c;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool act = tok2 < total;
        const int64_t goff = (act ? tok2 : 0) * (3 * H * kDh) + w.which * (H * kDh) + w.h * kDh + sub * 4;
        const int nn = act ? (int)n2 : 0;
        tok2 += w.step;
        n2 += w.step;
        while (n2 >= N) n2 -= N;
        // compiler barrier: keeps the (L1-resident) cos / sin loads of later vectors from being hoisted above this one's math --
        // 32 more live registers, which at 3 blocks per SM spilled 370 bytes and cost more than the unroll gained
        asm volatile("" ::: "memory");
        float dlo[4], dhi[4], xlo[4], xhi[4], cs[4], sn[4];
        if (w.which) {
          const float2 a = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&d0[u].x)), c = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&d0[u].y));
          const float2 e = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&d1[u].x)), f = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&d1[u].y));
          dlo[0] = a.x; dlo[1] = a.y; dlo[2] = c.x; dlo[3] = c.y;
          dhi[0] = e.x; dhi[1] = e.y; dhi[2] = f.x; dhi[3] = f.y;
        } else {
          dlo[0] = __uint_as_float(d0[u].x); dlo[1] = __uint_as_float(d0[u].y); dlo[2] = __uint_as_float(d0[u].z); dlo[3] = __uint_as_float(d0[u].w);
          dhi[0] = __uint_as_float(d1[u].x); dhi[1] = __uint_as_float(d1[u].y); dhi[2] = __uint_as_float(d1[u].z); dhi[3] = __uint_as_float(d1[u].w);
        }
        {
          const float2 a = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&x0[u].x)), c = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&x0[u].y));
          const float2 e = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&x1[u].x)), f = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&x1[u].y));
          xlo[0] = a.x; xlo[1] = a.y; xlo[2] = c.x; xlo[3] = c.y;
          xhi[0] = e.x; xhi[1] = e.y; xhi[2] = f.x; xhi[3] = f.y;
        }
        ld4c(cosv + nn * 32 + sub * 4, cs);
        ld4c(sinv + nn * 32 + sub * 4, sn);
        // undo the rotation (transpose): dz_lo = dy_lo c + dy_hi s ; dz_hi = dy_hi c - dy_lo s
        float zlo[4], zhi[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          zlo[i] = fmaf(dlo[i], cs[i], dhi[i] * sn[i]);
          zhi[i] = fmaf(dhi[i], cs[i], -dlo[i] * sn[i]);
        }
        if (gam != nullptr) {
          float ss = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) ss = fmaf(xlo[i], xlo[i], fmaf(xhi[i], xhi[i], ss));
          ss = sum8(ss);
          const float rinv = fminf(rsqrtf(ss), 1e12f);
          const float s1 = 8.0f * rinv;
          float dot = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (act) {
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
            zlo[i] = fmaf(zlo[i], s1, -xlo[i] * s2);
            zhi[i] = fmaf(zhi[i], s1, -xhi[i] * s2);
          }
        }
        if (act) {
          st4h(dqkv + goff, zlo);
          st4h(dqkv + goff + 32, zhi);
        }
      }
      if (!grouped && __all_sync(0xffffffffu, tok >= total)) break;
    }
    if (!grouped) break;
  }
  if (gam != nullptr && any) {
    float* dg = (w.which ? dgk : dgq) + w.h * kDh + sub * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(dg + i, dglo[i]);
      atomicAdd(dg + 32 + i, dghi[i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward, staged variant for H = 16 (2H = 32 = the block's vector slots): the loads go through shared memory as 1-D bulk
// copies (cp.async.bulk + mbarrier), so the bytes in flight are set by the stage size (40 KB per block, two blocks per SM) and
// not by registers.  Stage = 4 consecutive tokens: per token one 4 KB copy of its q|k block of qkv, per (token, head) one 256 B
// copy of dq (fp32) and one 128 B copy of dk (bf16) -- 132 copies, one per thread, all completing on one mbarrier.  While the
// 256 threads work on stage i (same math, same thread <-> (which, head, slice) assignment as the register kernel above), stage
// i+1 is in flight.  Results leave through registers as before.
// ------------------------------------------------------------------------------------------------------------------
#ifndef VBX_ROPE_PIECE
#define VBX_ROPE_PIECE 2048
#endif
namespace ropes {
constexpr int kTok = 4, kH = 16, kStages = 2;
constexpr uint32_t kXBytes = kTok * 2 * kH * kDh * 2;      // 16 KB: [tok][q|k][h][64] bf16
constexpr uint32_t kDqBytes = kTok * kH * kDh * 4;          // 16 KB: [tok][h][64] f32
constexpr uint32_t kDkBytes = kTok * kH * kDh * 2;          //  8 KB: [tok][h][64] bf16
constexpr uint32_t kStageBytes = kXBytes + kDqBytes + kDkBytes;
constexpr uint32_t kOffBar = kStages * kStageBytes;
constexpr uint32_t kSmemBytes = kOffBar + 64;
}  // namespace ropes

VBX_DEVINL void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ptx::smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(256, 2) qkrope_bwd_staged_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ cosv,
                                                                    const float* __restrict__ sinv, const float* __restrict__ gq,
                                                                    const float* __restrict__ gk, const float* __restrict__ dqh,
                                                                    const uint16_t* __restrict__ dkh, uint16_t* __restrict__ dqkv,
                                                                    float* __restrict__ dgq, float* __restrict__ dgk, int64_t B,
                                                                    int64_t N) {
  using namespace ropes;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  const int tid = threadIdx.x, sub = tid & 7, pair = tid >> 3, which = pair >> 4, h = pair & 15;
  const int64_t total = B * N, groups = (total + kTok - 1) / kTok;
  constexpr int H = kH;
  if (tid == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const float* gam = which ? gk : gq;
  float glo[4], ghi[4], dglo[4], dghi[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) glo[i] = ghi[i] = 1.f, dglo[i] = dghi[i] = 0.f;
  if (gam != nullptr) {
    ld4c(gam + h * kDh + sub * 4, glo);
    ld4c(gam + h * kDh + 32 + sub * 4, ghi);
  }
  // one copy per thread: threads 0..3 the qkv token blocks, threads 4..131 the (token, which, head) gradient pieces
  auto issue = [&](int64_t g, int stage) {
    uint8_t* st = smem + stage * kStageBytes;
    const int64_t tok0 = g * kTok;
    if (tid == 0) {
      const int64_t left = total - tok0;
      const uint32_t ntok = left < kTok ? (uint32_t)left : (uint32_t)kTok;
      ptx::mbar_arrive_expect_tx(&bars[stage], ntok * (kStageBytes / kTok));
    }
#if VBX_QK_TOKEN_MAJOR
    // a token's 10 KB = [4 KB q|k block of qkv | 4 KB dq (fp32) | 2 KB dk (bf16)], each contiguous in memory, fetched as
    // kPiece-byte bulk copies (one per thread)
    constexpr int kPiece = VBX_ROPE_PIECE, kPer = 10240 / kPiece;
    static_assert(2048 % kPiece == 0 && kPiece % 16 == 0 && kTok * kPer <= 256, "pieces must tile the 4 KB / 4 KB / 2 KB runs exactly");
    if (tid < kTok * kPer) {
      const int t = tid / kPer, off = (tid - t * kPer) * kPiece;
      const int64_t tok = tok0 + t;
      if (tok < total) {
        if (off < 4096) bulk_load_1d(st + t * 4096 + off, reinterpret_cast<const uint8_t*>(qkv + tok * (3 * H * kDh)) + off, kPiece, &bars[stage]);
        else if (off < 8192) bulk_load_1d(st + kXBytes + t * 4096 + (off - 4096), reinterpret_cast<const uint8_t*>(dqh + tok * (H * kDh)) + (off - 4096), kPiece, &bars[stage]);
        else bulk_load_1d(st + kXBytes + kDqBytes + t * 2048 + (off - 8192), reinterpret_cast<const uint8_t*>(dkh + tok * (H * kDh)) + (off - 8192), kPiece, &bars[stage]);
      }
    }
#else
    if (tid < kTok) {
      const int64_t tok = tok0 + tid;
      if (tok < total) bulk_load_1d(st + tid * (kXBytes / kTok), qkv + tok * (3 * H * kDh), kXBytes / kTok, &bars[stage]);
    } else if (tid < kTok + kTok * 2 * H) {
      const int j = tid - kTok, t = j >> 5, p = j & 31, wh = p >> 4, hh = p & 15;
      const int64_t tok = tok0 + t;
      if (tok < total) {
        const int64_t b = tok / N, n = tok - b * N;
        const int64_t hoff = ((b * H + hh) * N + n) * kDh;
        if (wh == 0) bulk_load_1d(st + kXBytes + (t * H + hh) * 256, dqh + hoff, 256, &bars[stage]);
        else bulk_load_1d(st + kXBytes + kDqBytes + (t * H + hh) * 128, dkh + hoff, 128, &bars[stage]);
      }
    }
#endif
  };
  int it = 0;
  if ((int64_t)blockIdx.x < groups) issue(blockIdx.x, 0);
  for (int64_t g = blockIdx.x; g < groups; g += gridDim.x, ++it) {
    const int stage = it & 1;
    if (g + gridDim.x < groups) issue(g + gridDim.x, stage ^ 1);   // that buffer was released by the __syncthreads below
    ptx::mbar_wait(&bars[stage], (it >> 1) & 1);
    const uint8_t* st = smem + stage * kStageBytes;
    int64_t tok = g * kTok;
    int64_t b = tok / N, n = tok - b * N;
#pragma unroll
    for (int t = 0; t < kTok; ++t) {
      const bool act = tok < total;
      float dlo[4], dhi[4], xlo[4], xhi[4], cs[4], sn[4];
      if (act) {
        const uint8_t* xp = st + t * (kXBytes / kTok) + which * (H * kDh * 2) + h * (kDh * 2) + sub * 8;
        const uint2 x0 = *reinterpret_cast<const uint2*>(xp), x1 = *reinterpret_cast<const uint2*>(xp + 64);
        const float2 a = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&x0.x)), c = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&x0.y));
        const float2 e = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&x1.x)), f = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&x1.y));
        xlo[0] = a.x; xlo[1] = a.y; xlo[2] = c.x; xlo[3] = c.y;
        xhi[0] = e.x; xhi[1] = e.y; xhi[2] = f.x; xhi[3] = f.y;
        if (which) {
          const uint8_t* dp = st + kXBytes + kDqBytes + (t * H + h) * 128 + sub * 8;
          const uint2 d0 = *reinterpret_cast<const uint2*>(dp), d1 = *reinterpret_cast<const uint2*>(dp + 64);
          const float2 a2 = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&d0.x)), c2 = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&d0.y));
          const float2 e2 = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&d1.x)), f2 = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&d1.y));
          dlo[0] = a2.x; dlo[1] = a2.y; dlo[2] = c2.x; dlo[3] = c2.y;
          dhi[0] = e2.x; dhi[1] = e2.y; dhi[2] = f2.x; dhi[3] = f2.y;
        } else {
          const uint8_t* dp = st + kXBytes + (t * H + h) * 256 + sub * 16;
          const float4 d0 = *reinterpret_cast<const float4*>(dp), d1 = *reinterpret_cast<const float4*>(dp + 128);
          dlo[0] = d0.x; dlo[1] = d0.y; dlo[2] = d0.z; dlo[3] = d0.w;
          dhi[0] = d1.x; dhi[1] = d1.y; dhi[2] = d1.z; dhi[3] = d1.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) dlo[i] = dhi[i] = 0.f, xlo[i] = xhi[i] = 1.f;   // keeps the lanes in the shuffles, results unused
      }
      const int nn = act ? (int)n : 0;
      ld4c(cosv + nn * 32 + sub * 4, cs);
      ld4c(sinv + nn * 32 + sub * 4, sn);
      // undo the rotation (transpose): dz_lo = dy_lo c + dy_hi s ; dz_hi = dy_hi c - dy_lo s
      float zlo[4], zhi[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        zlo[i] = fmaf(dlo[i], cs[i], dhi[i] * sn[i]);
        zhi[i] = fmaf(dhi[i], cs[i], -dlo[i] * sn[i]);
      }
      if (gam != nullptr) {
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) ss = fmaf(xlo[i], xlo[i], fmaf(xhi[i], xhi[i], ss));
        ss = sum8(ss);
        const float rinv = fminf(rsqrtf(ss), 1e12f);
        const float s1 = 8.0f * rinv;
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (act) {
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
          zlo[i] = fmaf(zlo[i], s1, -xlo[i] * s2);
          zhi[i] = fmaf(zhi[i], s1, -xhi[i] * s2);
        }
      }
      if (act) {
        uint16_t* dst = dqkv + tok * (3 * H * kDh) + which * (H * kDh) + h * kDh + sub * 4;
        st4h(dst, zlo);
        st4h(dst + 32, zhi);
      }
      ++tok;
      if (++n >= N) { n = 0; ++b; }
    }
    __syncthreads();   // every thread has read this stage: the next issue() may overwrite it
  }
  if (gam != nullptr) {
    float* dg = (which ? dgk : dgq) + h * kDh + sub * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(dg + i, dglo[i]);
      atomicAdd(dg + 32 + i, dghi[i]);
    }
  }
}

}  // namespace vbx

using namespace vbx;

// grid stride (grid * vectors per block) must be a multiple of 2H: round the grid to a multiple of m = 2H / gcd(vpb, 2H)
static int64_t rope_grid(int64_t nvec, int64_t H, int blocks_per_sm, int vec_per_block = kVecPerBlock) {
  int64_t a = vec_per_block, c = 2 * H;
  while (c) { int64_t t = a % c; a = c; c = t; }
  const int64_t m = (2 * H) / a;
  int64_t grid = grid_for(nvec, vec_per_block, blocks_per_sm);
  return ((grid + m - 1) / m) * m;
}

// grouped token order when a block's vector slots divide into whole tokens (see TokWalk); VBX_QKROPE_ORDER=strided forces the
// round-1 order (A/B runs)
struct RopeLaunch {
  int64_t grid;
  int grouped;
};
static RopeLaunch rope_launch(int64_t tokens, int64_t H, int blocks_per_sm, int vec_per_block) {
  static const bool strided = getenv("VBX_QKROPE_ORDER") != nullptr && getenv("VBX_QKROPE_ORDER")[0] == 's';
  RopeLaunch r;
  r.grouped = (!strided && vec_per_block % (2 * H) == 0) ? 1 : 0;
  if (r.grouped) {
    const int64_t step = vec_per_block / (2 * H), group = step > kTokGroup ? step : kTokGroup;
    const int64_t groups = (tokens + group - 1) / group, cap = (int64_t)kNumSM * blocks_per_sm;
    r.grid = groups < cap ? groups : cap;
  } else {
    r.grid = rope_grid(tokens * 2 * H, H, blocks_per_sm, vec_per_block);
  }
  return r;
}

extern "C" int vbx_qkrope_fwd(const uint16_t* qkv, const float* cosv, const float* sinv, const float* gq, const float* gk,
                              uint16_t* qh, uint16_t* kh, int64_t B, int64_t N, int64_t H, void* stream) {
  VBX_REQUIRE(qkv && cosv && sinv && qh && kh, VBX_E_NULL);
  VBX_REQUIRE((gq == nullptr) == (gk == nullptr), VBX_E_NULL);
  VBX_REQUIRE(B > 0 && N > 0 && H > 0, VBX_E_SHAPE);
  VBX_REQUIRE(VBX_ALIGNED16(qkv) && VBX_ALIGNED16(qh) && VBX_ALIGNED16(kh), VBX_E_ALIGN);
  const int64_t nvec = B * N * 2 * H;
  const RopeLaunch rl = rope_launch(B * N, H, VBX_QKROPE_FWD_MINB, kVecPerBlock);
  const int64_t step = kVecPerBlock / (2 * H);
  const bool un = rl.grouped && step >= 1 && ((step > kTokGroup ? 1 : kTokGroup / step) % VBX_QKROPE_FWD_U == 0);
  if (un)
    qkrope_fwd_kernel<VBX_QKROPE_FWD_U><<<(unsigned)rl.grid, 256, 0, (cudaStream_t)stream>>>(qkv, cosv, sinv, gq, gk, qh, kh, B, N, (int)H, rl.grouped);
  else
    qkrope_fwd_kernel<1><<<(unsigned)rl.grid, 256, 0, (cudaStream_t)stream>>>(qkv, cosv, sinv, gq, gk, qh, kh, B, N, (int)H, rl.grouped);
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
  // H = 16: the shared-memory staged kernel (VBX_QKROPE_BWD=regs selects the register kernel for A/B runs)
  static const bool regs_only = getenv("VBX_QKROPE_BWD") != nullptr && getenv("VBX_QKROPE_BWD")[0] == 'r';
  if (H == ropes::kH && !regs_only && VBX_ALIGNED16(qkv) && (N * 64 * 2) % 16 == 0) {
    const int64_t groups = (B * N + ropes::kTok - 1) / ropes::kTok;
    const int64_t grid = groups < 2 * kNumSM ? groups : 2 * kNumSM;
    cudaError_t ce = cudaFuncSetAttribute(qkrope_bwd_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ropes::kSmemBytes);
    if (ce != cudaSuccess) return (int)ce;
    qkrope_bwd_staged_kernel<<<(unsigned)grid, 256, ropes::kSmemBytes, (cudaStream_t)stream>>>(qkv, cosv, sinv, gq, gk, dqh, dkh, dqkv, dgq,
                                                                                               dgk, B, N);
    return VBX_LAUNCH_RC();
  }
  const RopeLaunch rl = rope_launch(B * N, H, VBX_QKROPE_BWD_MINB, kVecPerBlockB);
  // VBX_QKROPE_BWD_U (4) vectors in flight per thread when the grouped walk has a multiple of that many passes per group, else one
  const int64_t step = kVecPerBlockB / (2 * H > 0 ? 2 * H : 1);
  const bool u4 = rl.grouped && step >= 1 && ((step > kTokGroup ? 1 : kTokGroup / step) % VBX_QKROPE_BWD_U == 0);
  if (u4)
    qkrope_bwd_kernel<VBX_QKROPE_BWD_U><<<(unsigned)rl.grid, 256, 0, (cudaStream_t)stream>>>(qkv, cosv, sinv, gq, gk, dqh, dkh, dqkv, dgq, dgk, B, N, (int)H,
                                                                              rl.grouped);
  else
    qkrope_bwd_kernel<1><<<(unsigned)rl.grid, 256, 0, (cudaStream_t)stream>>>(qkv, cosv, sinv, gq, gk, dqh, dkh, dqkv, dgq, dgk, B, N, (int)H,
                                                                              rl.grouped);
  return VBX_LAUNCH_RC();
}
