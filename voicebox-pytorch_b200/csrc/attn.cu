// Flash attention for the Voicebox trunk on tcgen05 tensor cores (sm_100a): softmax(scale * Q K^T + key mask) V, dim_head 64.
//
// Operands are staged by TMA (SWIZZLE_128B boxes of 128 rows x 64 bf16) into the canonical UMMA shared-memory layouts,
// the two GEMMs of every tile run as tcgen05.mma (M = 128) with fp32 accumulators in TMEM, and the softmax runs with ONE
// THREAD PER QUERY ROW straight out of TMEM (tcgen05.ld 32x32b): no shuffles, no shared-memory round trip for S.
// Warp roles:  warps 0-3 softmax / correction / epilogue,  warp 4 TMA producer,  warp 5 TMEM allocator + MMA issuer.
//
// Replaces attend.py:100-137 (math path: einsum, scale, masked_fill(-finfo.max), softmax, einsum) and the SDPA
// delegation attend.py:71-98; the head merge 'b h n d -> b n (h d)' (vp.py:332) is folded into the epilogue store.
#include <cfloat>
#include <mutex>

#include "umma.cuh"

namespace vbx {

using namespace ptx;

constexpr int kBM = 128;  // query rows per CTA / per tile
constexpr int kBN = 128;  // keys per tile
constexpr int kDh = 64;
constexpr uint32_t kTileBytes = kBN * kDh * 2;  // 16 KB

VBX_DEVINL float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
VBX_DEVINL float bf16_bits_to_float(uint16_t b) { return __uint_as_float((uint32_t)b << 16); }

// =====================================================================================================================
// forward
// =====================================================================================================================
namespace fwd {
constexpr uint32_t kOffQ = 0, kOffK = 16384, kOffV = 49152, kOffP = 81920, kOffBar = 114688, kOffBias = kOffBar + 128;
constexpr uint32_t kSmemBytes = kOffBias + 2 * kBN * 2;  // 115,328 B -> two CTAs per SM
enum { Q_FULL = 0, KV_FULL = 1, KV_EMPTY = 3, S_FULL = 5, S_FREE = 6, P_FULL = 7, O_FULL = 8, NUM_BARS = 9 };
constexpr uint32_t kTmemCols = 256;  // S: [0,128)  O_tile: [128,192)
}  // namespace fwd

__global__ void __launch_bounds__(192, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                const __grid_constant__ CUtensorMap mv, const uint8_t* __restrict__ key_mask, float scale_log2,
                uint16_t* __restrict__ o, float* __restrict__ lse, int N, int H) {
  using namespace fwd;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + NUM_BARS * 8);
  uint16_t* s_bias = reinterpret_cast<uint16_t*>(smem + kOffBias);  // [2][128] bf16: 0 / -FLT_MAX-ish / -inf

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kBM, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (N + kBN - 1) / kBN;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(&bars[Q_FULL], 1);
    mbar_init(&bars[KV_FULL], 1);
    mbar_init(&bars[KV_FULL + 1], 1);
    mbar_init(&bars[KV_EMPTY], 1);
    mbar_init(&bars[KV_EMPTY + 1], 1);
    mbar_init(&bars[S_FULL], 1);
    mbar_init(&bars[S_FREE], 128);
    mbar_init(&bars[P_FULL], 128);
    mbar_init(&bars[O_FULL], 1);
    fence_barrier_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&mq);
    tma_prefetch_desc(&mk);
    tma_prefetch_desc(&mv);
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------ TMA producer ------------------------------------------------
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars[Q_FULL], kTileBytes);
      tma_load_4d(smem + kOffQ, &mq, &bars[Q_FULL], 0, q0, h, b);
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        mbar_wait(&bars[KV_EMPTY + st], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars[KV_FULL + st], 2 * kTileBytes);
        tma_load_4d(smem + kOffK + st * kTileBytes, &mk, &bars[KV_FULL + st], 0, j * kBN, h, b);
        tma_load_4d(smem + kOffV + st * kTileBytes, &mv, &bars[KV_FULL + st], 0, j * kBN, h, b);
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------ MMA issuer --------------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(kBM, kBN, false, false);  // S = Q K^T      (both K-major)
      constexpr uint32_t idesc_o = make_idesc(kBM, kDh, false, true);   // O = P V        (V is MN-major: [keys][d])
      const uint32_t aQ = smem_u32(smem + kOffQ), aP = smem_u32(smem + kOffP);
      mbar_wait(&bars[Q_FULL], 0);
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        const uint32_t aK = smem_u32(smem + kOffK + st * kTileBytes), aV = smem_u32(smem + kOffV + st * kTileBytes);
        mbar_wait(&bars[KV_FULL + st], (j >> 1) & 1);
        mbar_wait(&bars[S_FREE], (j & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k) umma_bf16(tmem_base, desc_kmajor(aQ, k), desc_kmajor(aK, k), idesc_s, k > 0);
        umma_commit(&bars[S_FULL]);
        mbar_wait(&bars[P_FULL], j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kBN / 16; ++k)
          umma_bf16(tmem_base + 128, desc_kmajor(aP, k), desc_mnmajor(aV, k), idesc_o, k > 0);
        umma_commit(&bars[O_FULL]);
        umma_commit(&bars[KV_EMPTY + st]);
      }
    }
  } else {
    // ------------------------------------------------ softmax: one thread per query row ---------------------------
    const int r = threadIdx.x;  // 0..127 == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    float m = -FLT_MAX, l = 0.f;
    float acc[kDh];
#pragma unroll
    for (int i = 0; i < kDh; ++i) acc[i] = 0.f;

    for (int j = 0; j < nkv; ++j) {
      const int k0 = j * kBN;
      const bool masked_tile = (key_mask != nullptr) || (k0 + kBN > N);  // CTA-uniform
      const uint16_t* bias = s_bias + (j & 1) * kBN;
      if (masked_tile) {
        const int key = k0 + r;
        uint16_t v = 0;
        if (key >= N) v = 0xFF80;                                             // -inf: the key does not exist
        else if (key_mask != nullptr && !key_mask[(int64_t)b * N + key]) v = 0xFF7F;  // -3.39e38 ~ -finfo.max fill
        s_bias[(j & 1) * kBN + r] = v;
        named_bar_sync(1, 128);
      }
      mbar_wait(&bars[S_FULL], j & 1);
      tc_fence_after();
      // pass 1: running max of the scaled, masked logits (log2 domain)
      float mx = m;
#pragma unroll 1
      for (int c = 0; c < kBN / 32; ++c) {
        float s[32];
        tmem_ld32(t_lane + c * 32, s);
        if (masked_tile) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaf(s[i], scale_log2, bf16_bits_to_float(bias[c * 32 + i])));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, s[i] * scale_log2);
        }
      }
      const float alpha = ex2(m - mx);
      m = mx;
      // pass 2: p = 2^(t - m), row sum, P -> shared memory (bf16, K-major SW128, two 64-key sub-tiles)
      float rowsum = 0.f;
      const float neg_m = -m;
#pragma unroll 1
      for (int c = 0; c < kBN / 32; ++c) {
        float s[32];
        tmem_ld32(t_lane + c * 32, s);
        if (c == kBN / 32 - 1) {  // S is in registers: the MMA warp may overwrite it
          tc_fence_before();
          mbar_arrive(&bars[S_FREE]);
        }
        if (masked_tile) {
#pragma unroll
          for (int i = 0; i < 32; ++i) s[i] = ex2(fmaf(s[i], scale_log2, bf16_bits_to_float(bias[c * 32 + i])) + neg_m);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) s[i] = ex2(fmaf(s[i], scale_log2, neg_m));
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) rowsum += s[i];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c16 = c * 4 + q;
          uint8_t* dst = smem + kOffP + (c16 >> 3) * kSubTileBytes + r * 128 + (((c16 & 7) ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(dst) = pack8(&s[q * 8]);
        }
      }
      l = fmaf(l, alpha, rowsum);
      fence_proxy_async();
      mbar_arrive(&bars[P_FULL]);
      // O_acc = O_acc * alpha + P V
      mbar_wait(&bars[O_FULL], j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < kDh / 32; ++c) {
        float v[32];
        tmem_ld32(t_lane + 128 + c * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[c * 32 + i] = fmaf(acc[c * 32 + i], alpha, v[i]);
      }
      tc_fence_before();
    }
    // epilogue: normalise, merge heads ('b h n d -> b n (h d)'), log-sum-exp for the backward
    const int q = q0 + r;
    if (q < N) {
      const float inv_l = 1.0f / l;
      uint16_t* dst = o + (((int64_t)b * N + q) * H + h) * kDh;
#pragma unroll
      for (int c = 0; c < kDh / 8; ++c) {
        float t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = acc[c * 8 + i] * inv_l;
        stg_16(dst + c * 8, pack8(t));
      }
      if (lse != nullptr) lse[((int64_t)b * H + h) * N + q] = m + log2f(l);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, kTmemCols);
}

// =====================================================================================================================
// backward
// =====================================================================================================================
// delta[b,h,n] = sum_d O * dO   (8 lanes per head vector; o / dout are [B,N,H*64])
__global__ void __launch_bounds__(256) attn_delta_kernel(const uint16_t* __restrict__ o, const uint16_t* __restrict__ dout,
                                                          float* __restrict__ delta, int64_t B, int64_t N, int H) {
  const int sub = threadIdx.x & 7;
  const int64_t nvec = B * N * H;
  const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 3);
  for (int64_t vid = (int64_t)blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);; vid += stride) {
    const bool active = vid < nvec;
    const int64_t v = active ? vid : 0;
    float a[8], d[8];
    unpack8(ldg_nc_16(o + v * kDh + sub * 8), a);
    unpack8(ldg_nc_16(dout + v * kDh + sub * 8), d);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(a[i], d[i], s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (active && sub == 0) {
      const int64_t tok = v / H, hh = v - tok * H, bb = tok / N, n = tok - bb * N;
      delta[(bb * H + hh) * N + n] = s;
    }
    if (__all_sync(0xffffffffu, vid + stride >= nvec)) break;
  }
}

namespace bwd {
constexpr uint32_t kOffK = 0, kOffV = 16384, kOffQ = 32768, kOffdO = 65536, kOffPT = 98304, kOffdST = 131072,
                   kOffBar = 163840, kOffLse = kOffBar + 128, kOffDelta = kOffLse + 2 * kBM * 4;
constexpr uint32_t kSmemBytes = kOffDelta + 2 * kBM * 4;  // 166,016 B: one CTA per SM
enum { KV_FULL = 0, QD_FULL = 1, QD_EMPTY = 3, ST_FULL = 5, ST_FREE = 6, DS_FULL = 7, DQ_FULL = 8, NUM_BARS = 9 };
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColST = 0, kColDPT = 128, kColDV = 256, kColDK = 320, kColDQ = 384;
}  // namespace bwd

VBX_DEVINL void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// One CTA per (key tile, head, batch); loops over the query tiles.  Everything is computed TRANSPOSED (keys on the
// TMEM lanes) so that P^T and dS^T come out K-major for the dV / dK GEMMs, which accumulate in TMEM across the loop:
//   S^T = K Q^T, dP^T = V dO^T  ->  P^T = 2^(c S^T - lse), dS^T = P^T (dP^T - delta)
//   dV += P^T dO,  dK += dS^T Q,  dQ_i = dS K  (fp32 red.global.add into dq; scaled by `scale`)
__global__ void __launch_bounds__(192, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                const __grid_constant__ CUtensorMap mv, const __grid_constant__ CUtensorMap mdo,
                const uint8_t* __restrict__ key_mask, float scale, float scale_log2, const float* __restrict__ lse,
                const float* __restrict__ delta, float* __restrict__ dq, uint16_t* __restrict__ dk, uint16_t* __restrict__ dv,
                int64_t dv_bs, int64_t dv_ns, int N, int H) {
  using namespace bwd;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + NUM_BARS * 8);
  float* s_lse = reinterpret_cast<float*>(smem + kOffLse);      // [2][128]
  float* s_delta = reinterpret_cast<float*>(smem + kOffDelta);  // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * kBN, h = blockIdx.y, b = blockIdx.z;
  const int nq = (N + kBM - 1) / kBM;
  const int64_t bh = (int64_t)b * H + h;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(&bars[KV_FULL], 1);
    mbar_init(&bars[QD_FULL], 1);
    mbar_init(&bars[QD_FULL + 1], 1);
    mbar_init(&bars[QD_EMPTY], 1);
    mbar_init(&bars[QD_EMPTY + 1], 1);
    mbar_init(&bars[ST_FULL], 1);
    mbar_init(&bars[ST_FREE], 128);
    mbar_init(&bars[DS_FULL], 128);
    mbar_init(&bars[DQ_FULL], 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars[KV_FULL], 2 * kTileBytes);
      tma_load_4d(smem + kOffK, &mk, &bars[KV_FULL], 0, k0, h, b);
      tma_load_4d(smem + kOffV, &mv, &bars[KV_FULL], 0, k0, h, b);
      for (int i = 0; i < nq; ++i) {
        const int st = i & 1;
        mbar_wait(&bars[QD_EMPTY + st], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars[QD_FULL + st], 2 * kTileBytes);
        tma_load_4d(smem + kOffQ + st * kTileBytes, &mq, &bars[QD_FULL + st], 0, i * kBM, h, b);
        tma_load_4d(smem + kOffdO + st * kTileBytes, &mdo, &bars[QD_FULL + st], 0, i * kBM, h, b);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc_kk = make_idesc(128, 128, false, false);
      constexpr uint32_t idesc_kmn = make_idesc(128, kDh, false, true);
      constexpr uint32_t idesc_mnmn = make_idesc(128, kDh, true, true);
      const uint32_t aK = smem_u32(smem + kOffK), aV = smem_u32(smem + kOffV), aPT = smem_u32(smem + kOffPT),
                     aDST = smem_u32(smem + kOffdST);
      mbar_wait(&bars[KV_FULL], 0);
      for (int i = 0; i < nq; ++i) {
        const int st = i & 1;
        const uint32_t aQ = smem_u32(smem + kOffQ + st * kTileBytes), aDO = smem_u32(smem + kOffdO + st * kTileBytes);
        mbar_wait(&bars[QD_FULL + st], (i >> 1) & 1);
        mbar_wait(&bars[ST_FREE], (i & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k)  // S^T = K Q^T
          umma_bf16(tmem_base + kColST, desc_kmajor(aK, k), desc_kmajor(aQ, k), idesc_kk, k > 0);
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k)  // dP^T = V dO^T
          umma_bf16(tmem_base + kColDPT, desc_kmajor(aV, k), desc_kmajor(aDO, k), idesc_kk, k > 0);
        umma_commit(&bars[ST_FULL]);
        mbar_wait(&bars[DS_FULL], i & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kBM / 16; ++k)  // dV += P^T dO
          umma_bf16(tmem_base + kColDV, desc_kmajor(aPT, k), desc_mnmajor(aDO, k), idesc_kmn, (i > 0) || (k > 0));
#pragma unroll
        for (int k = 0; k < kBM / 16; ++k)  // dK += dS^T Q
          umma_bf16(tmem_base + kColDK, desc_kmajor(aDST, k), desc_mnmajor(aQ, k), idesc_kmn, (i > 0) || (k > 0));
#pragma unroll
        for (int k = 0; k < kBN / 16; ++k)  // dQ_i = dS K    (A = dS^T read MN-major)
          umma_bf16(tmem_base + kColDQ, desc_mnmajor(aDST, k), desc_mnmajor(aK, k), idesc_mnmn, k > 0);
        umma_commit(&bars[DQ_FULL]);
        umma_commit(&bars[QD_EMPTY + st]);
      }
    }
  } else {
    const int r = threadIdx.x;  // key row of this tile == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int key = k0 + r;
    float bias = 0.f;
    if (key >= N) bias = -INFINITY;
    else if (key_mask != nullptr && !key_mask[(int64_t)b * N + key]) bias = -FLT_MAX;

    for (int i = 0; i < nq; ++i) {
      const int st = i & 1, q0 = i * kBM;
      {
        const int q = q0 + r;
        s_lse[st * kBM + r] = (q < N) ? lse[bh * N + q] : INFINITY;  // +inf -> p = 0 for padded queries
        s_delta[st * kBM + r] = (q < N) ? delta[bh * N + q] : 0.f;
        named_bar_sync(1, 128);
      }
      mbar_wait(&bars[ST_FULL], i & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < kBM / 32; ++c) {
        float s[32], dp[32];
        tmem_ld32(t_lane + kColST + c * 32, s);
        tmem_ld32(t_lane + kColDPT + c * 32, dp);
        if (c == kBM / 32 - 1) {
          tc_fence_before();
          mbar_arrive(&bars[ST_FREE]);
        }
#pragma unroll
        for (int x = 0; x < 32; ++x) {
          const float p = ex2(fmaf(s[x], scale_log2, bias) - s_lse[st * kBM + c * 32 + x]);
          s[x] = p;
          dp[x] = p * (dp[x] - s_delta[st * kBM + c * 32 + x]);
        }
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          const int c16 = c * 4 + qd;
          const uint32_t off = (c16 >> 3) * kSubTileBytes + r * 128 + (((c16 & 7) ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(smem + kOffPT + off) = pack8(&s[qd * 8]);
          *reinterpret_cast<uint4*>(smem + kOffdST + off) = pack8(&dp[qd * 8]);
        }
      }
      fence_proxy_async();
      mbar_arrive(&bars[DS_FULL]);
      // dQ tile: TMEM lane = query row
      mbar_wait(&bars[DQ_FULL], i & 1);
      tc_fence_after();
      const int q = q0 + r;
#pragma unroll
      for (int c = 0; c < kDh / 32; ++c) {
        float v[32];
        tmem_ld32(t_lane + kColDQ + c * 32, v);
        if (q < N) {
          float* dst = dq + (bh * N + q) * kDh + c * 32;
#pragma unroll
          for (int x = 0; x < 32; x += 4) red_add_v4(dst + x, v[x] * scale, v[x + 1] * scale, v[x + 2] * scale, v[x + 3] * scale);
        }
      }
      tc_fence_before();
    }
    // all MMAs are complete (DQ_FULL of the last tile): write dV, dK for this key row.  The TMEM loads are
    // .sync.aligned (whole warp, converged); only the global stores are predicated on the key being real.
    {
      const bool live = key < N;
      uint16_t* dvp = dv + (int64_t)b * dv_bs + (int64_t)(live ? key : 0) * dv_ns + h * kDh;
      uint16_t* dkp = dk + (bh * N + (live ? key : 0)) * kDh;
#pragma unroll
      for (int c = 0; c < kDh / 32; ++c) {
        float v[32];
        tmem_ld32(t_lane + kColDV + c * 32, v);
        if (live) {
#pragma unroll
          for (int x = 0; x < 32; x += 8) stg_16(dvp + c * 32 + x, pack8(&v[x]));
        }
        tmem_ld32(t_lane + kColDK + c * 32, v);
#pragma unroll
        for (int x = 0; x < 32; ++x) v[x] *= scale;
        if (live) {
#pragma unroll
          for (int x = 0; x < 32; x += 8) stg_16(dkp + c * 32 + x, pack8(&v[x]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, kTmemCols);
}

// =====================================================================================================================
// tcgen05 / TMA self-test: C[128x128] = A B^T with K = 128, every operand-major combination, manual or TMA staging
// =====================================================================================================================
template <bool kTma>
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mb,
                     const uint16_t* __restrict__ A, const uint16_t* __restrict__ Bm, float* __restrict__ C, int a_mn, int b_mn) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + 2 * kSubTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * kSubTileBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (kTma) {
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(&bars[0], 4 * kSubTileBytes);
      tma_load_4d(sA, &ma, &bars[0], 0, 0, 0, 0);
      tma_load_4d(sA + kSubTileBytes, &ma, &bars[0], 64, 0, 0, 0);
      tma_load_4d(sB, &mb, &bars[0], 0, 0, 0, 0);
      tma_load_4d(sB + kSubTileBytes, &mb, &bars[0], 64, 0, 0, 0);
    }
    mbar_wait(&bars[0], 0);
  } else {
    // global [128][128] (second index contiguous) -> two SW128 sub-tiles, written with generic-proxy stores
    for (int i = threadIdx.x; i < 128 * 128; i += blockDim.x) {
      const int row = i >> 7, col = i & 127;
      *reinterpret_cast<uint16_t*>(sA + sw128_offset(row, col)) = A[i];
      *reinterpret_cast<uint16_t*>(sB + sw128_offset(row, col)) = Bm[i];
    }
    fence_proxy_async();
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc(128, 128, a_mn != 0, b_mn != 0);
    const uint32_t aA = smem_u32(sA), aB = smem_u32(sB);
    for (int k = 0; k < 8; ++k)
      umma_bf16(tmem_base, a_mn ? desc_mnmajor(aA, k) : desc_kmajor(aA, k), b_mn ? desc_mnmajor(aB, k) : desc_kmajor(aB, k),
                idesc, k > 0);
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < 4; ++c) {
    float v[32];
    tmem_ld32(t_lane + c * 32, v);
    for (int x = 0; x < 32; ++x) C[threadIdx.x * 128 + c * 32 + x] = v[x];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int make_tmap(CUtensorMap* out, const void* base, int64_t inner, int64_t N, int64_t H, int64_t B, int64_t n_stride,
                     int64_t h_stride, int64_t b_stride, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return VBX_E_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (n_stride * 2) % 16 || (h_stride * 2) % 16 || (b_stride * 2) % 16)
    return VBX_E_ALIGN;
  cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)N, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)n_stride * 2, (cuuint64_t)h_stride * 2, (cuuint64_t)b_stride * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult rc = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? VBX_OK : VBX_E_DRIVER;
}

int make_tmap_bf16_4d(CUtensorMap* out, const void* base, int64_t N, int64_t H, int64_t B, int64_t n_stride, int64_t h_stride,
                      int64_t b_stride, int box_rows) {
  return make_tmap(out, base, kDh, N, H, B, n_stride, h_stride, b_stride, box_rows);
}

}  // namespace vbx

using namespace vbx;

static const float kLog2e = 1.4426950408889634f;

extern "C" int vbx_attn_fwd(const uint16_t* q, const uint16_t* k, const uint16_t* v, int64_t v_bs, int64_t v_ns,
                            const uint8_t* key_mask, float scale, uint16_t* o, float* lse, int64_t B, int64_t H, int64_t N,
                            void* stream) {
  VBX_REQUIRE(q && k && v && o, VBX_E_NULL);
  VBX_REQUIRE(B > 0 && H > 0 && N > 0 && B < 65536 && H < 65536 && N < (1 << 24), VBX_E_SHAPE);
  VBX_REQUIRE(VBX_ALIGNED16(o), VBX_E_ALIGN);
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = make_tmap_bf16_4d(&mq, q, N, H, B, kDh, N * kDh, H * N * kDh, kBM)) != VBX_OK) return rc;
  if ((rc = make_tmap_bf16_4d(&mk, k, N, H, B, kDh, N * kDh, H * N * kDh, kBN)) != VBX_OK) return rc;
  if ((rc = make_tmap_bf16_4d(&mv, v, N, H, B, v_ns, kDh, v_bs, kBN)) != VBX_OK) return rc;
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd::kSmemBytes);
  });
  dim3 grid((unsigned)((N + kBM - 1) / kBM), (unsigned)H, (unsigned)B);
  attn_fwd_kernel<<<grid, 192, fwd::kSmemBytes, (cudaStream_t)stream>>>(mq, mk, mv, key_mask, scale * kLog2e, o, lse, (int)N,
                                                                       (int)H);
  return VBX_LAUNCH_RC();
}

extern "C" int vbx_attn_bwd(const uint16_t* q, const uint16_t* k, const uint16_t* v, int64_t v_bs, int64_t v_ns,
                            const uint8_t* key_mask, float scale, const uint16_t* o, const uint16_t* dout, const float* lse,
                            float* delta, float* dq, uint16_t* dk, uint16_t* dv, int64_t dv_bs, int64_t dv_ns, int64_t B,
                            int64_t H, int64_t N, void* stream) {
  VBX_REQUIRE(q && k && v && o && dout && lse && delta && dq && dk && dv, VBX_E_NULL);
  VBX_REQUIRE(B > 0 && H > 0 && N > 0 && B < 65536 && H < 65536 && N < (1 << 24), VBX_E_SHAPE);
  VBX_REQUIRE(VBX_ALIGNED16(o) && VBX_ALIGNED16(dout) && VBX_ALIGNED16(dq) && VBX_ALIGNED16(dk) && VBX_ALIGNED16(dv) &&
                  (dv_bs % 8 == 0) && (dv_ns % 8 == 0),
              VBX_E_ALIGN);
  CUtensorMap mq, mk, mv, mdo;
  int rc;
  if ((rc = make_tmap_bf16_4d(&mq, q, N, H, B, kDh, N * kDh, H * N * kDh, kBM)) != VBX_OK) return rc;
  if ((rc = make_tmap_bf16_4d(&mk, k, N, H, B, kDh, N * kDh, H * N * kDh, kBN)) != VBX_OK) return rc;
  if ((rc = make_tmap_bf16_4d(&mv, v, N, H, B, v_ns, kDh, v_bs, kBN)) != VBX_OK) return rc;
  if ((rc = make_tmap_bf16_4d(&mdo, dout, N, H, B, H * kDh, kDh, N * H * kDh, kBM)) != VBX_OK) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  attn_delta_kernel<<<grid_for(B * N * H, 32, 8), 256, 0, s>>>(o, dout, delta, B, N, (int)H);
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd::kSmemBytes);
  });
  dim3 grid((unsigned)((N + kBN - 1) / kBN), (unsigned)H, (unsigned)B);
  attn_bwd_kernel<<<grid, 192, bwd::kSmemBytes, s>>>(mq, mk, mv, mdo, key_mask, scale, scale * kLog2e, lse, delta, dq, dk, dv,
                                                     dv_bs, dv_ns, (int)N, (int)H);
  return VBX_LAUNCH_RC();
}

extern "C" int vbx_umma_selftest(const uint16_t* a, const uint16_t* b, float* c, int variant, void* stream) {
  VBX_REQUIRE(a && b && c, VBX_E_NULL);
  VBX_REQUIRE(variant >= 0 && variant < 8, VBX_E_SHAPE);
  const int b_mn = variant & 1, a_mn = (variant >> 1) & 1, tma = (variant >> 2) & 1;
  CUtensorMap ma, mb;
  int rc;
  if ((rc = make_tmap(&ma, a, 128, 128, 1, 1, 128, 128 * 128, 128 * 128, 128)) != VBX_OK) return rc;
  if ((rc = make_tmap(&mb, b, 128, 128, 1, 1, 128, 128 * 128, 128 * 128, 128)) != VBX_OK) return rc;
  const int smem = 4 * ptx::kSubTileBytes + 64;
  cudaStream_t s = (cudaStream_t)stream;
  if (tma) {
    cudaFuncSetAttribute(umma_selftest_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    umma_selftest_kernel<true><<<1, 128, smem, s>>>(ma, mb, a, b, c, a_mn, b_mn);
  } else {
    cudaFuncSetAttribute(umma_selftest_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    umma_selftest_kernel<false><<<1, 128, smem, s>>>(ma, mb, a, b, c, a_mn, b_mn);
  }
  return VBX_LAUNCH_RC();
}
