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
#include <cstdlib>
#include <mutex>

#include "umma.cuh"

namespace vbx {

using namespace ptx;

constexpr int kBM = 128;  // query rows per CTA / per tile
constexpr int kBN = 128;  // keys per tile
constexpr int kDh = 64;
constexpr uint32_t kTileBytes = kBN * kDh * 2;  // 16 KB

// EXPERIMENT, off in libvbx_sm100a.so (csrc/build_exp.sh builds lib/libvbx_exp.so with it on; select with VBX_LIB=...):
// tail-aware tiles.  N' = N + 16 register tokens = 8*128 + 16, so the last key tile (forward) / last query tile (backward)
// is 7/8 padding.  With the flag on, its S / S^T / dP^T GEMMs run with N = roundup(valid, 16) instead of 128, its P V / dV /
// dK GEMMs with valid/16 K-steps instead of 8, and the softmax / exp code skips (warp-uniformly) chunks that are all padding.
#ifndef VBX_EXP_TAIL
#define VBX_EXP_TAIL 1   // on since round 2: parity-green on every tail geometry, -5 % backward time (profiles/README.md)
#endif
// EXPERIMENT (backward), off in libvbx_sm100a.so: VBX_EXP_DSBUF=1 double-buffers dS^T in shared memory (paid for by
// staging dQ one 32-column half at a time) and releases P^T (TMEM) with its own barrier right after the dV GEMM.  In the
// shipped kernel the compute warps may only store P^T / dS^T of tile i+1 once ALL of tile i's dV/dK/dQ GEMMs have retired,
// so per tile the GEMM group (~1900 clk in the trace) and the store phase (~800 clk) run back to back; with the flag the
// stores of tile i+1 overlap the dK/dQ GEMMs of tile i.
#ifndef VBX_EXP_DSBUF
#define VBX_EXP_DSBUF 1  // on since round 2: parity-green, -4 % backward time; required by VBX_BWD_ISSUERS == 3
#endif

// Backward: number of tcgen05.mma ISSUING warps.  One thread sustains one tcgen05.mma per ~55-70 clk whatever its shape (ptxas
// wraps every MMA in an ELECT / BRA.U.ANY loop; tools/umma_bench.py: N = 64 TS-mode MMAs retire at 66 clk each from one
// issuing warp, 33 clk from two, 32 clk from three = the nominal rate).  The backward issues 40 MMAs per tile pair: from one
// warp that is ~2500 clk of pure issue against ~1540 clk of tensor-pipe time, and the trace showed the MMA warp issuing
// back to back for the whole tile.  With 3: warp A = S^T / dP^T of the next tile, warp B = dV + dK, warp C = dQ.
#ifndef VBX_BWD_ISSUERS
#define VBX_BWD_ISSUERS 3
#endif
#if VBX_BWD_ISSUERS != 1 && VBX_BWD_ISSUERS != 3
#error "VBX_BWD_ISSUERS must be 1 or 3"
#endif
#if VBX_BWD_ISSUERS == 3 && !VBX_EXP_DSBUF
#error "three issuing warps need the double-buffered dS^T tile (VBX_EXP_DSBUF=1)"
#endif

// ---- optional pipeline trace (built only with -DVBX_TRACE into lib/libvbx_trace.so; tools/trace_attn.py reads it) ----------
#ifdef VBX_TRACE
__device__ long long* g_trace = nullptr;
#define TRACE(role, tile, point)                                                                            \
  do {                                                                                                      \
    if (g_trace != nullptr && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && (tile) < 16)         \
      g_trace[((role) * 16 + (tile)) * 8 + (point)] = clock64();                                            \
  } while (0)
#else
#define TRACE(role, tile, point) \
  do {                           \
  } while (0)
#endif

VBX_DEVINL float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// EXPERIMENT, off in libvbx_sm100a.so (VBX_EXP_DEFS=-DVBX_EXP_POLY=k csrc/build_exp.sh): both attention kernels are bound by
// the 16 ex2/clk/SM of the MUFU pipe (1024 clk per 128x128 tile) while the FMA pipe idles.  With k in 1..4, k of every 4
// consecutive exponentials are evaluated on the FMA/ALU pipes instead: round-to-nearest range reduction with the 1.5*2^23
// magic constant, a degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (max relative error 7.5e-5, 50x below the bf16
// rounding P gets anyway; fit + fp32 emulation in tools/fit_exp2_poly.py), exponent inserted with one shift-add.
// Inputs are clamped at -125, so -inf gives 2^-125 (2.4e-38) instead of 0: only used where that is harmless (see call sites).
#ifndef VBX_EXP_POLY
#define VBX_EXP_POLY 0
#endif
VBX_DEVINL float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;       // low mantissa bits of t = round(x)
  const float f = x - (t - 12582912.0f);  // in [-0.5, 0.5]
  float p = fmaf(f, 0.0551716685295105f, 0.2426111251115799f);
  p = fmaf(p, f, 0.6932609677314758f);
  p = fmaf(p, f, 0.9999280571937561f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// element i of an unrolled loop: MUFU or polynomial, decided at compile time
#define VBX_EX2_AT(i, x) ((((i) & 3) < VBX_EXP_POLY) ? ex2_poly(x) : ex2(x))
VBX_DEVINL float bf16_bits_to_float(uint16_t b) { return __uint_as_float((uint32_t)b << 16); }

// =====================================================================================================================
// forward
// =====================================================================================================================
// CTA = 128 query rows of one (batch, head); 2 CTAs per SM.  10 warps:
//   warps 0-7  softmax: row r = 32*(w%4)+lane (the TMEM lane quarter a warp may touch), column half w/4 -- two threads per
//              row, each owning 64 of the 128 key columns of S and 32 of the 64 columns of O, so every SM sub-partition
//              holds 4 softmax warps (latency hiding) instead of 1; the two halves agree on the running row max through
//              a 1 KB shared-memory exchange + one named barrier per tile.
//   warp 8     TMA producer (Q once; K and V double-buffered, L2 prefetch two tiles ahead)
//   warp 9     TMEM allocator + tcgen05.mma issuer
// P never touches shared memory: the softmax threads write it (bf16, two keys per 32-bit column) straight into TMEM with
// tcgen05.st and O = P V runs as a TS-mode MMA (A from TMEM, only V read from shared memory).  With SS-mode MMAs the SM's
// 128 B/clk of shared-memory bandwidth -- not the tensor pipe -- was the limiter (8 KB of operands per 64-clk MMA).
namespace fwd {
constexpr uint32_t kOffQ = 0, kOffK = 16384, kOffV = 49152, kOffBar = 81920, kOffBias = kOffBar + 128,
                   kOffMax = kOffBias + 2 * kBN * 4;
constexpr uint32_t kSmemBytes = kOffMax + 2 * 2 * kBM * 4;  // 85,120 B -> two CTAs per SM
enum { Q_FULL = 0, K_FULL = 1, K_EMPTY = 3, V_FULL = 5, V_EMPTY = 7, S_FULL = 9, S_FREE = 10, P_FULL = 11, O_FULL = 12, NUM_BARS = 13 };
constexpr uint32_t kTmemCols = 256;  // S: [0,128)  O_tile: [128,192)  P (bf16 pairs): [192,256)
constexpr uint32_t kColO = 128, kColP = 192;
constexpr int kThreads = 320;
}  // namespace fwd

// packed-fp32 helpers of the forward softmax (VBX_FWD_F32X2=0 restores the scalar instruction streams for A/B runs)
#ifndef VBX_FWD_F32X2
#define VBX_FWD_F32X2 1
#endif
VBX_DEVINL float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
VBX_DEVINL void fma_scale32(float (&acc)[32], float alpha, const float (&v)[32]) {   // acc = acc * alpha + v
#if VBX_FWD_F32X2
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    const float2 t = __ffma2_rn(make_float2(acc[i], acc[i + 1]), make_float2(alpha, alpha), make_float2(v[i], v[i + 1]));
    acc[i] = t.x;
    acc[i + 1] = t.y;
  }
#else
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = fmaf(acc[i], alpha, v[i]);
#endif
}

__global__ void __launch_bounds__(fwd::kThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                const __grid_constant__ CUtensorMap mv, const uint8_t* __restrict__ key_mask, float scale_log2,
                uint16_t* __restrict__ o, float* __restrict__ lse, int N, int H) {
  using namespace fwd;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + NUM_BARS * 8);
  float* s_bias = reinterpret_cast<float*>(smem + kOffBias);  // [2][128]: 0 / -FLT_MAX / -inf per key of the tile
  float* s_max = reinterpret_cast<float*>(smem + kOffMax);    // [2][2][128]: per-tile partial row max of each column half

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kBM, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (N + kBN - 1) / kBN;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(&bars[Q_FULL], 1);
    mbar_init(&bars[K_FULL], 1);
    mbar_init(&bars[K_FULL + 1], 1);
    mbar_init(&bars[K_EMPTY], 1);
    mbar_init(&bars[K_EMPTY + 1], 1);
    mbar_init(&bars[V_FULL], 1);
    mbar_init(&bars[V_FULL + 1], 1);
    mbar_init(&bars[V_EMPTY], 1);
    mbar_init(&bars[V_EMPTY + 1], 1);
    mbar_init(&bars[S_FULL], 1);
    mbar_init(&bars[S_FREE], 256);
    mbar_init(&bars[P_FULL], 256);
    mbar_init(&bars[O_FULL], 1);
    fence_barrier_init();
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&mq);
    tma_prefetch_desc(&mk);
    tma_prefetch_desc(&mv);
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------ TMA producer ------------------------------------------------
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars[Q_FULL], kTileBytes);
      tma_load_4d(smem + kOffQ, &mq, &bars[Q_FULL], 0, q0, h, b);
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        if (j + 2 < nkv) {  // warm L2 two tiles ahead
          tma_prefetch_l2_4d(&mk, 0, (j + 2) * kBN, h, b);
          tma_prefetch_l2_4d(&mv, 0, (j + 2) * kBN, h, b);
        }
        mbar_wait(&bars[K_EMPTY + st], ((j >> 1) & 1) ^ 1);  // S(j-2) has retired
        mbar_arrive_expect_tx(&bars[K_FULL + st], kTileBytes);
        tma_load_4d(smem + kOffK + st * kTileBytes, &mk, &bars[K_FULL + st], 0, j * kBN, h, b);
        mbar_wait(&bars[V_EMPTY + st], ((j >> 1) & 1) ^ 1);  // PV(j-2) has retired
        mbar_arrive_expect_tx(&bars[V_FULL + st], kTileBytes);
        tma_load_4d(smem + kOffV + st * kTileBytes, &mv, &bars[V_FULL + st], 0, j * kBN, h, b);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------ MMA issuer --------------------------------------------------
    // Warp-uniform control flow: all 32 lanes run the loop and the barrier waits, so the descriptors (hoisted out of the
    // loop, per-K-step constants added) stay in uniform registers; only tcgen05.mma / commit are issued by the leader lane.
    {
      constexpr uint32_t idesc_s = make_idesc(kBM, kBN, false, false);  // S = Q K^T      (both K-major)
      constexpr uint32_t idesc_o = make_idesc(kBM, kDh, false, true);   // O = P V        (V is MN-major: [keys][d])
      const uint64_t dQ = sdesc_k0(smem_u32(smem + kOffQ));
      const uint64_t dK0 = sdesc_k0(smem_u32(smem + kOffK)), dV0 = sdesc_mn0(smem_u32(smem + kOffV));
      const bool leader = lane == 0;
      const int n_tail = VBX_EXP_TAIL ? ((N - (nkv - 1) * kBN + 15) & ~15) : kBN;  // GEMM width of the last key tile
      const uint32_t idesc_s_tail = make_idesc(kBM, n_tail, false, false);
      mbar_wait(&bars[Q_FULL], 0);
      auto issue_s = [&](int j) {  // S_j = Q K_j^T
        const int st = j & 1;
        const uint64_t dK = dK0 + (uint64_t)st * (kTileBytes >> 4);
        const uint32_t idesc_sj = (VBX_EXP_TAIL && j == nkv - 1) ? idesc_s_tail : idesc_s;
        TRACE(3, j, 0);
        mbar_wait(&bars[K_FULL + st], (j >> 1) & 1);
        TRACE(3, j, 1);
        mbar_wait(&bars[S_FREE], (j & 1) ^ 1);  // the softmax warps hold S_{j-1} in registers
        TRACE(3, j, 2);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < kDh / 16; ++k) umma_bf16(tmem_base, dQ + koff_k(k), dK + koff_k(k), idesc_sj, k > 0);
          umma_commit(&bars[S_FULL]);
          umma_commit(&bars[K_EMPTY + st]);
        }
        __syncwarp();
        TRACE(3, j, 3);
      };
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        const uint64_t dV = dV0 + (uint64_t)st * (kTileBytes >> 4);
        // S_{j+1} goes to the tensor pipe BEFORE P_j V_j: it only needs S to be released (which happens just before P_j is
        // published), so the next tile's logits are ready ~300 clk after the softmax of this tile instead of ~800
        if (j + 1 < nkv) issue_s(j + 1);
        mbar_wait(&bars[V_FULL + st], (j >> 1) & 1);
        TRACE(3, j, 4);
        mbar_wait(&bars[P_FULL], j & 1);
        TRACE(3, j, 5);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < kBN / 16; ++k)  // A = P from TMEM: 8 columns (16 keys) per K-step
            if (!VBX_EXP_TAIL || j < nkv - 1 || k * 16 < n_tail)
              umma_bf16_ts(tmem_base + kColO, tmem_base + kColP + k * 8, dV + koff_mn(k), idesc_o, k > 0);
          umma_commit(&bars[O_FULL]);
          umma_commit(&bars[V_EMPTY + st]);
        }
        __syncwarp();
        TRACE(3, j, 6);
      }
    }
  } else {
    // ------------------------------------------------ softmax: two threads per query row ---------------------------
    const int half = warp >> 2;                 // which 64 key columns of S / which 32 columns of O
    const int r = (warp & 3) * 32 + lane;       // query row of the tile == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    float m = -FLT_MAX, l = 0.f;                // l: partial row sum over this thread's column half
    float alpha_prev = 0.f;                     // rescale factor of the tile whose P V is still to be accumulated
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;

    for (int j = 0; j < nkv; ++j) {
      const int k0 = j * kBN;
      const bool masked_tile = (key_mask != nullptr) || (k0 + kBN > N);  // CTA-uniform
      const float* bias = s_bias + (j & 1) * kBN + half * 64;
      if (masked_tile) {
        if (half == 0) {
          const int key = k0 + r;
          float v = 0.f;
          if (key >= N) v = -INFINITY;                                           // the key does not exist
          else if (key_mask != nullptr && !key_mask[(int64_t)b * N + key]) v = -FLT_MAX;  // masked_fill(-finfo.max)
          s_bias[(j & 1) * kBN + r] = v;
        }
        named_bar_sync(1, 256);
      }
      if (threadIdx.x == 0) TRACE(4, j, 0);
      mbar_wait(&bars[S_FULL], j & 1);
      if (threadIdx.x == 0) TRACE(4, j, 1);
      tc_fence_after();
      // pass 1: partial max over this thread's 64 columns (scaled, masked logits; log2 domain)
      float mx = -FLT_MAX;
      const int valid_cols = N - k0;  // >= 128 except in the tail tile
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (VBX_EXP_TAIL && half * 64 + c * 32 >= valid_cols) continue;  // warp-uniform: every key of this chunk is padding
        float s[32];
        tmem_ld32(t_lane + half * 64 + c * 32, s);
        if (masked_tile) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaf(s[i], scale_log2, bias[c * 32 + i]));
        } else {
          float m4[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
#if VBX_FWD_F32X2
#pragma unroll
          for (int i = 0; i < 32; i += 8) {   // sm_100 3-input max: half the instructions of the max chain
            m4[0] = fmax3(m4[0], s[i], s[i + 1]);
            m4[1] = fmax3(m4[1], s[i + 2], s[i + 3]);
            m4[2] = fmax3(m4[2], s[i + 4], s[i + 5]);
            m4[3] = fmax3(m4[3], s[i + 6], s[i + 7]);
          }
#else
#pragma unroll
          for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], s[i]);
#endif
          mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * scale_log2);  // scale > 0
        }
      }
      s_max[((j & 1) * 2 + half) * kBM + r] = mx;
      if (threadIdx.x == 0) TRACE(4, j, 2);
      named_bar_sync(2, 256);
      if (threadIdx.x == 0) TRACE(4, j, 3);
      mx = fmaxf(fmaxf(mx, s_max[((j & 1) * 2 + (half ^ 1)) * kBM + r]), m);
      const float alpha = ex2(m - mx);
      m = mx;
      // deferred accumulation of the PREVIOUS tile: O_acc = O_acc * alpha_{j-1} + P_{j-1} V_{j-1}.  By now that GEMM has long
      // retired, so this wait is off the critical path; it must precede this tile's P store (P / O_tile are single-buffered).
      if (j > 0) {
        mbar_wait(&bars[O_FULL], (j - 1) & 1);
        tc_fence_after();
        float v[32];
        tmem_ld32(t_lane + kColO + half * 32, v);
        fma_scale32(acc, alpha_prev, v);
        tc_fence_before();
      }
      alpha_prev = alpha;
      if (threadIdx.x == 0) TRACE(4, j, 6);
      // pass 2: p = 2^(t - m), partial row sum, P -> TMEM (bf16 pairs: column kColP + key/2 of this row's lane)
      float rowsum = 0.f;
      const float neg_m = -m;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (VBX_EXP_TAIL && half * 64 + c * 32 >= valid_cols) {
          // warp-uniform: all-padding chunk.  Its P columns belong to K-steps >= n_tail/16 that the P V GEMM does not issue,
          // so nothing is loaded, computed or stored; only the S hand-off still has to happen.
          if (c == 1) {
            tc_fence_before();
            mbar_arrive(&bars[S_FREE]);
          }
          continue;
        }
        float s[32];
        tmem_ld32(t_lane + half * 64 + c * 32, s);
        if (c == 1) {  // this thread's part of S is in registers: release it to the MMA warp
          tc_fence_before();
          mbar_arrive(&bars[S_FREE]);
        }
        if (masked_tile) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float bb = bias[c * 32 + i];
            float t = fmaf(s[i], scale_log2, bb);
            // S columns >= n_tail of the tail tile were not written by the narrowed GEMM: whatever bits they hold must not
            // reach the row sum (NaN + -inf = NaN), so non-existent keys are selected, not added, to -inf
            if (VBX_EXP_TAIL) t = (bb == -INFINITY) ? -INFINITY : t;
            s[i] = ex2(t + neg_m);
          }
        } else {
#if VBX_FWD_F32X2
#pragma unroll
          for (int i = 0; i < 32; i += 2) {   // every key exists: no -inf here.  Packed fp32 FMA: same rounding, half the issue slots
            const float2 t = __ffma2_rn(make_float2(s[i], s[i + 1]), make_float2(scale_log2, scale_log2), make_float2(neg_m, neg_m));
            s[i] = VBX_EX2_AT(i, t.x);
            s[i + 1] = VBX_EX2_AT(i + 1, t.y);
          }
#else
#pragma unroll
          for (int i = 0; i < 32; ++i) s[i] = VBX_EX2_AT(i, fmaf(s[i], scale_log2, neg_m));  // every key exists: no -inf here
#endif
        }
#if VBX_FWD_F32X2
        float2 r2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int i = 0; i < 32; i += 4) {     // same four partial sums (i & 3) as the scalar chain, two per packed add
          r2[0] = __fadd2_rn(r2[0], make_float2(s[i], s[i + 1]));
          r2[1] = __fadd2_rn(r2[1], make_float2(s[i + 2], s[i + 3]));
        }
        rowsum += (r2[0].x + r2[0].y) + (r2[1].x + r2[1].y);
#else
        float r4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 32; ++i) r4[i & 3] += s[i];
        rowsum += (r4[0] + r4[1]) + (r4[2] + r4[3]);
#endif
        uint32_t pk[16];
#pragma unroll
        for (int x = 0; x < 16; ++x) {
          __nv_bfloat162 t2 = f2bf(s[2 * x], s[2 * x + 1]);
          pk[x] = *reinterpret_cast<uint32_t*>(&t2);
        }
        tmem_st16(t_lane + kColP + half * 32 + c * 16, pk);
      }
      l = fmaf(l, alpha, rowsum);
      if (threadIdx.x == 0) TRACE(4, j, 4);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&bars[P_FULL]);
      if (threadIdx.x == 0) TRACE(4, j, 5);
    }
    {  // the last tile's P V
      mbar_wait(&bars[O_FULL], (nkv - 1) & 1);
      tc_fence_after();
      float v[32];
      tmem_ld32(t_lane + kColO + half * 32, v);
      fma_scale32(acc, alpha_prev, v);
      tc_fence_before();
    }
    // epilogue: total row sum (both halves), normalise, merge heads ('b h n d -> b n (h d)'), log-sum-exp for the backward
    named_bar_sync(2, 256);                       // everyone is done reading s_max of the last tiles
    s_max[half * kBM + r] = l;
    named_bar_sync(2, 256);
    l += s_max[(half ^ 1) * kBM + r];
    const int q = q0 + r;
    if (q < N) {
      const float inv_l = 1.0f / l;
      uint16_t* dst = o + (((int64_t)b * N + q) * H + h) * kDh + half * 32;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = acc[c * 8 + i] * inv_l;
        stg_16(dst + c * 8, pack8(t));
      }
      if (lse != nullptr && half == 0) lse[((int64_t)b * H + h) * N + q] = m + log2f(l);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, kTmemCols);
}

// =====================================================================================================================
// forward, second generation (default; VBX_ATTN_FWD_V1=1 selects the kernel above for A/B runs)
// =====================================================================================================================
// Same CTA shape, roles, barriers and TMEM map as the first kernel (S: [0,128)  O: [128,192)  P: [192,256), 2 CTAs / SM).  What
// changed is the per-tile dependency chain of the softmax warps, which -- not any pipe -- bounded the first kernel (two CTAs x
// ~3700 clk per tile against a 1024-clk MUFU floor; profiles/README.md):
//   * ONE pass over S: each thread pulls its 64 logits out of TMEM once, releases S to the MMA warp at once (S_{j+1} = Q K^T
//     runs under this tile's exponentials), and keeps them in registers for the max and the exp.
//   * O stays in TMEM: P_j V_j accumulates straight into the O columns (no per-tile read-modify-write of O in registers).
//   * LAZY rescale: the running max used for the exponentials is only advanced (and O, l rescaled through one TMEM
//     load / multiply / store) when the true row max has grown by more than 2^8 -- p <= 256 is harmless in bf16 / fp32 and the
//     final normalisation by l (accumulated against the same stale max) cancels it exactly.  After the first tiles of a row
//     the rescale is rare, so the P_{j-1} V_{j-1} -> O_j dependency leaves the critical path.
//   * the two threads of a row agree on the row max through a 64-thread named barrier of just their two warps.
//   * the tail key tile (N' = 8*128 + 16) runs its S GEMM at N = roundup16(valid), its P V GEMM over valid/16 K-steps, and
//     32-column chunks that are all padding are skipped outright (no TMEM traffic, no exponentials).
// EXPERIMENT: the two CTAs of an SM start together and stay in lock step (both in their exponential phase, fighting for the MUFU
// pipe, then both outside it with the pipe idle: MUFU utilisation 56 % in the round-2 trace).  With VBX_FWD_STAGGER = n > 0 the
// SECOND first-wave CTA to land on an SM (per-SM arrival counter in device memory, zeroed by the host before the launch) lets its
// softmax warps spin n clocks before their first tile; every later CTA starts when a predecessor ends and inherits the offset.
#ifndef VBX_FWD_STAGGER
#define VBX_FWD_STAGGER 0
#endif
// debug bisection of the v2 forward (bit 0: rescale O on every tile; bit 1: release S only at the end of the tile;
// bit 2: CTA-wide barrier for the max exchange; bit 3: no tail narrowing)
#ifndef VBX_V2_DBG
#define VBX_V2_DBG 0
#endif
// Which forward ships: 1 = two-pass softmax, O accumulated in registers (round 1, + tail narrowing); 2 = single pass, O resident in
// TMEM, lazy rescale; 3 = two query tiles per CTA with alternating exponential phases.  Measured on the same B200 at the bench
// geometry (B=64, H=16, N'=1040; tools/kbench.py, warm-ups excluded): v1 611.5 us, v2 631.5 us, v3 733 us -- the later
// generations remove work the trace blamed (second TMEM pass, O round trip, lock-stepped MUFU phases) without getting faster,
// see DESIGN.md section 5 for what that says about the real limiter.  All three are parity-green and bitwise repeatable.
#ifndef VBX_ATTN_FWD_DEFAULT
#define VBX_ATTN_FWD_DEFAULT 1
#endif
__device__ unsigned g_sm_slot[256];

// 64-thread named barrier of the two warps that own the same 32 query rows (ids 2..5; literal ids so that ptxas does not
// reserve all 16 hardware barriers for the CTA)
VBX_DEVINL void pair_sync(int quarter) {
#if VBX_V2_DBG & 4
  named_bar_sync(2, 256);
  return;
#endif
  switch (quarter) {
    case 0: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 5, 64;" ::: "memory"); break;
  }
}

__global__ void __launch_bounds__(fwd::kThreads, 2)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                 const __grid_constant__ CUtensorMap mv, const uint8_t* __restrict__ key_mask, float scale_log2,
                 uint16_t* __restrict__ o, float* __restrict__ lse, int N, int H) {
  using namespace fwd;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + NUM_BARS * 8);
  float* s_bias = reinterpret_cast<float*>(smem + kOffBias);  // [2][128]: 0 / -FLT_MAX / -inf per key of the tile
  float* s_max = reinterpret_cast<float*>(smem + kOffMax);    // [2][2][128]: per-tile partial row max of each column half

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kBM, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (N + kBN - 1) / kBN;
  const int n_tail = (VBX_V2_DBG & 8) ? kBN : ((N - (nkv - 1) * kBN + 15) & ~15);   // GEMM width of the last key tile (16 .. 128)

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(&bars[Q_FULL], 1);
    mbar_init(&bars[K_FULL], 1);
    mbar_init(&bars[K_FULL + 1], 1);
    mbar_init(&bars[K_EMPTY], 1);
    mbar_init(&bars[K_EMPTY + 1], 1);
    mbar_init(&bars[V_FULL], 1);
    mbar_init(&bars[V_FULL + 1], 1);
    mbar_init(&bars[V_EMPTY], 1);
    mbar_init(&bars[V_EMPTY + 1], 1);
    mbar_init(&bars[S_FULL], 1);
    mbar_init(&bars[S_FREE], 256);
    mbar_init(&bars[P_FULL], 256);
    mbar_init(&bars[O_FULL], 1);
    fence_barrier_init();
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&mq);
    tma_prefetch_desc(&mk);
    tma_prefetch_desc(&mv);
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  if (VBX_FWD_STAGGER > 0 && threadIdx.x == 32) {
    const unsigned lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    unsigned slot = 0;
    if (lin < 2u * kNumSM) {   // first wave only
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      slot = atomicAdd(&g_sm_slot[smid & 255u], 1u) & 1u;
    }
    tmem_slot[1] = slot;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------ TMA producer ------------------------------------------------
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars[Q_FULL], kTileBytes);
      tma_load_4d(smem + kOffQ, &mq, &bars[Q_FULL], 0, q0, h, b);
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        if (j + 2 < nkv) {  // warm L2 two tiles ahead
          tma_prefetch_l2_4d(&mk, 0, (j + 2) * kBN, h, b);
          tma_prefetch_l2_4d(&mv, 0, (j + 2) * kBN, h, b);
        }
        mbar_wait(&bars[K_EMPTY + st], ((j >> 1) & 1) ^ 1);  // S(j-2) has retired
        mbar_arrive_expect_tx(&bars[K_FULL + st], kTileBytes);
        tma_load_4d(smem + kOffK + st * kTileBytes, &mk, &bars[K_FULL + st], 0, j * kBN, h, b);
        mbar_wait(&bars[V_EMPTY + st], ((j >> 1) & 1) ^ 1);  // PV(j-2) has retired
        mbar_arrive_expect_tx(&bars[V_FULL + st], kTileBytes);
        tma_load_4d(smem + kOffV + st * kTileBytes, &mv, &bars[V_FULL + st], 0, j * kBN, h, b);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------ MMA issuer (v2) ---------------------------------------------
    constexpr uint32_t idesc_s = make_idesc(kBM, kBN, false, false);  // S = Q K^T      (both K-major)
    constexpr uint32_t idesc_o = make_idesc(kBM, kDh, false, true);   // O += P V       (V is MN-major: [keys][d])
    const uint64_t dQ = sdesc_k0(smem_u32(smem + kOffQ));
    const uint64_t dK0 = sdesc_k0(smem_u32(smem + kOffK)), dV0 = sdesc_mn0(smem_u32(smem + kOffV));
    const bool leader = lane == 0;
    const uint32_t idesc_s_tail = make_idesc(kBM, n_tail, false, false);
    mbar_wait(&bars[Q_FULL], 0);
    auto issue_s = [&](int j) {  // S_j = Q K_j^T
      const int st = j & 1;
      const uint64_t dK = dK0 + (uint64_t)st * (kTileBytes >> 4);
      const uint32_t idesc_sj = (j == nkv - 1) ? idesc_s_tail : idesc_s;
      TRACE(3, j, 0);
      mbar_wait(&bars[K_FULL + st], (j >> 1) & 1);
      TRACE(3, j, 1);
      mbar_wait(&bars[S_FREE], (j & 1) ^ 1);  // the softmax warps hold S_{j-1} in registers
      TRACE(3, j, 2);
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k) umma_bf16(tmem_base, dQ + koff_k(k), dK + koff_k(k), idesc_sj, k > 0);
        umma_commit(&bars[S_FULL]);
        umma_commit(&bars[K_EMPTY + st]);
      }
      __syncwarp();
      TRACE(3, j, 3);
    };
    issue_s(0);
    for (int j = 0; j < nkv; ++j) {
      const int st = j & 1;
      const uint64_t dV = dV0 + (uint64_t)st * (kTileBytes >> 4);
      if (j + 1 < nkv) issue_s(j + 1);  // only needs S released, which now happens right after the softmax warps' single load
      mbar_wait(&bars[V_FULL + st], (j >> 1) & 1);
      TRACE(3, j, 4);
      mbar_wait(&bars[P_FULL], j & 1);
      TRACE(3, j, 5);
      tc_fence_after();
      if (leader) {
        const int ksteps = (j == nkv - 1) ? n_tail / 16 : kBN / 16;
#pragma unroll
        for (int k = 0; k < kBN / 16; ++k)  // A = P from TMEM: 8 columns (16 keys) per K-step; O accumulates across tiles
          if (k < ksteps) umma_bf16_ts(tmem_base + kColO, tmem_base + kColP + k * 8, dV + koff_mn(k), idesc_o, (j > 0) || (k > 0));
        umma_commit(&bars[O_FULL]);
        umma_commit(&bars[V_EMPTY + st]);
      }
      __syncwarp();
      TRACE(3, j, 6);
    }
  } else {
    // ------------------------------------------------ softmax: two threads per query row ---------------------------
    const int half = warp >> 2;                 // which 64 key columns of S / which 32 columns of O
    const int r = (warp & 3) * 32 + lane;       // query row of the tile == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    float m_used = -FLT_MAX, l = 0.f;           // max the exponentials are taken against; partial row sum of this half
    if (VBX_FWD_STAGGER > 0 && tmem_slot[1] != 0u) {   // second CTA of this SM: start half a tile period late
      const long long t0 = clock64();
      while (clock64() - t0 < (long long)VBX_FWD_STAGGER) {
      }
    }

    for (int j = 0; j < nkv; ++j) {
      const int k0 = j * kBN;
      const bool masked_tile = (key_mask != nullptr) || (k0 + kBN > N);  // CTA-uniform
      const float* bias = s_bias + (j & 1) * kBN + half * 64;
      if (masked_tile) {
        if (half == 0) {
          const int key = k0 + r;
          float v = 0.f;
          if (key >= N) v = -INFINITY;                                           // the key does not exist
          else if (key_mask != nullptr && !key_mask[(int64_t)b * N + key]) v = -FLT_MAX;  // masked_fill(-finfo.max)
          s_bias[(j & 1) * kBN + r] = v;
        }
        named_bar_sync(1, 256);
      }
      const int valid = N - k0;                                  // >= 128 except in the tail tile
      const bool live0 = (VBX_V2_DBG & 8) || half * 64 < valid, live1 = (VBX_V2_DBG & 8) || half * 64 + 32 < valid;   // warp-uniform
      if (threadIdx.x == 0) TRACE(4, j, 0);
      mbar_wait(&bars[S_FULL], j & 1);
      if (threadIdx.x == 0) TRACE(4, j, 1);
      tc_fence_after();
      float s[64];
      {
        // one tcgen05.ld + tcgen05.wait::ld per 32 columns (the pattern of tmem_ld32): with two loads in flight behind a single
        // wait the compiler is free to shuffle the first load's destination registers before the wait, i.e. before the hardware
        // has written them -- seen on the B200 as sporadic wrong logits (non-repeatable outputs)
        if (live0) tmem_ld32(t_lane + half * 64, *reinterpret_cast<float(*)[32]>(&s[0]));
        if (live1) tmem_ld32(t_lane + half * 64 + 32, *reinterpret_cast<float(*)[32]>(&s[32]));
      }
#if !(VBX_V2_DBG & 2)
      tc_fence_before();
      mbar_arrive(&bars[S_FREE]);                                // this thread's part of S is in registers
#endif
      // local max over this thread's live columns (log2 domain)
      float mx = -FLT_MAX;
      if (masked_tile) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (!(c ? live1 : live0)) continue;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float bb = bias[c * 32 + i];
            float t = fmaf(s[c * 32 + i], scale_log2, bb);
            // columns >= n_tail of the tail tile were not written by the narrowed GEMM (stale bits, possibly NaN): select
            t = (bb == -INFINITY) ? -INFINITY : t;
            s[c * 32 + i] = t;
            mx = fmaxf(mx, t);
          }
        }
      } else {
        float m4[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
#pragma unroll
        for (int i = 0; i < 64; ++i) m4[i & 3] = fmaxf(m4[i & 3], s[i]);
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * scale_log2;  // scale > 0
      }
      s_max[((j & 1) * 2 + half) * kBM + r] = mx;
      pair_sync(warp & 3);
      mx = fmaxf(mx, s_max[((j & 1) * 2 + (half ^ 1)) * kBM + r]);
      if (threadIdx.x == 0) TRACE(4, j, 2);
      bool o_ready = false;                                      // O_FULL(j-1) already waited for in this tile
      if (j == 0) {
        m_used = mx;
      } else {
        const float m_new = fmaxf(m_used, mx);
        const bool need = (VBX_V2_DBG & 1) ? true : (m_new - m_used) > 8.0f;   // lazy: only when p could exceed 2^8
        if (__any_sync(0xffffffffu, need)) {                     // TMEM accesses are warp-wide: lanes that do not need it use 1
          const float alpha = need ? ex2(m_used - m_new) : 1.0f;
          mbar_wait(&bars[O_FULL], (j - 1) & 1);                 // P_{j-1} V_{j-1} has been accumulated
          tc_fence_after();
          o_ready = true;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {                          // 16 columns at a time: s[64] is live across this block
            float ov[16];
            tmem_ld16(t_lane + kColO + half * 32 + c * 16, ov);
            uint32_t ou[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) ou[i] = __float_as_uint(ov[i] * alpha);
            tmem_st16(t_lane + kColO + half * 32 + c * 16, ou);
          }
          l *= alpha;
          if (need) m_used = m_new;
        }
      }
      const float neg_m = -m_used;
      // exponentials, partial row sum, P -> TMEM (bf16 pairs: column kColP + key/2 of this row's lane)
      float rowsum = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (!(c ? live1 : live0)) continue;                      // all-padding chunk: its K-steps are not issued either
        float* sc = s + c * 32;
        if (masked_tile) {
#pragma unroll
          for (int i = 0; i < 32; ++i) sc[i] = ex2(sc[i] + neg_m);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) sc[i] = VBX_EX2_AT(i, fmaf(sc[i], scale_log2, neg_m));  // every key exists: no -inf here
        }
        float r4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 32; ++i) r4[i & 3] += sc[i];
        rowsum += (r4[0] + r4[1]) + (r4[2] + r4[3]);
        uint32_t pk[16];
#pragma unroll
        for (int x = 0; x < 16; ++x) {
          __nv_bfloat162 t2 = f2bf(sc[2 * x], sc[2 * x + 1]);
          pk[x] = *reinterpret_cast<uint32_t*>(&t2);
        }
        if (j > 0 && !o_ready) {                                 // P is single-buffered: P_{j-1} V_{j-1} must have read it
          mbar_wait(&bars[O_FULL], (j - 1) & 1);
          tc_fence_after();
          o_ready = true;
        }
        tmem_st16(t_lane + kColP + half * 32 + c * 16, pk);
      }
      l += rowsum;
      if (j > 0 && !o_ready) {
        // a warp whose chunks are all padding (tail tile) stored nothing, but it must still OBSERVE this phase of O_FULL: a parity
        // wait that skips a phase sees the barrier "already flipped" the next time and returns before P V has finished (the
        // epilogue then read O early: sporadic wrong rows at N' = 8*128 + 16, caught by tools/debug_attn_repeat.py)
        mbar_wait(&bars[O_FULL], (j - 1) & 1);
        tc_fence_after();
      }
      if (threadIdx.x == 0) TRACE(4, j, 4);
      tmem_st_wait();
      tc_fence_before();
#if VBX_V2_DBG & 2
      mbar_arrive(&bars[S_FREE]);
#endif
      mbar_arrive(&bars[P_FULL]);
      if (threadIdx.x == 0) TRACE(4, j, 5);
    }
    // epilogue: O (accumulated in TMEM), total row sum (both halves), normalise, merge heads, log-sum-exp for the backward
    mbar_wait(&bars[O_FULL], (nkv - 1) & 1);
    tc_fence_after();
    float acc[32];
    tmem_ld32(t_lane + kColO + half * 32, acc);
    tc_fence_before();
    s_max[((nkv & 1) * 2 + half) * kBM + r] = l;                 // the buffer the last tile did not use
    pair_sync(warp & 3);
    l += s_max[((nkv & 1) * 2 + (half ^ 1)) * kBM + r];
    const int q = q0 + r;
    if (q < N) {
      const float inv_l = 1.0f / l;
      uint16_t* dst = o + (((int64_t)b * N + q) * H + h) * kDh + half * 32;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = acc[c * 8 + i] * inv_l;
        stg_16(dst + c * 8, pack8(t));
      }
      if (lse != nullptr && half == 0) lse[((int64_t)b * H + h) * N + q] = m_used + log2f(l);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, kTmemCols);
}

// =====================================================================================================================
// forward, third generation: TWO query tiles per CTA with strictly alternating exponential phases
// =====================================================================================================================
// The v2 trace (profiles/README.md) showed the two co-resident CTAs of an SM in lock step: both in their exponential phase at
// the same time (each at half the MUFU rate: 2210 clk for 1024 clk of work), then both outside it with the MUFU pipe idle
// (~1450 clk) -- 56 % MUFU utilisation with MUFU as the binding resource (16 ex2 / clk / SM = 1024 clk per 128 x 128 tile).
// Two independent CTAs cannot be forced out of phase; two GROUPS inside one CTA can: this kernel runs one CTA per SM, 256 query
// rows (two 128-row tiles A and B of the same (batch, head)), each with its own 8 softmax warps, its own MMA-issuing warp and
// its own 256 TMEM columns (S | O | P as in v2), sharing one K/V stream (3-stage TMA ring: K and V are fetched once for both
// tiles).  A pair of named barriers passes a token between the groups: a group takes the token before its exponentials and
// hands it over right after them, so the exponential phases strictly alternate (A, B, A, B ...) and everything else a group
// does per tile -- P store, S load, row max, lazy rescale -- runs while the OTHER group owns the MUFU pipe.
namespace fwd3 {
constexpr int kStages = 3;
constexpr uint32_t kOffQ = 0, kOffK = 32768, kOffV = kOffK + kStages * 16384, kOffBar = kOffV + kStages * 16384,   // 131072
                   kOffBias = kOffBar + 256, kOffMax = kOffBias + 2 * 2 * kBN * 4, kSmemBytes = kOffMax + 2 * 2 * 2 * kBM * 4;
enum { Q_FULL = 0, K_FULL = 1, K_EMPTY = 4, V_FULL = 7, V_EMPTY = 10, S_FULL = 13, S_FREE = 15, P_FULL = 17, O_FULL = 19, NUM_BARS = 21 };
static_assert(NUM_BARS * 8 + 8 <= 256, "barrier block");
constexpr uint32_t kColO = 128, kColP = 192, kGroupCols = 256;
constexpr int kProducerWarp = 16, kMmaWarp0 = 17;          // warps 0-7: softmax of tile A, 8-15: tile B, 17 / 18: MMA issuers of A / B
constexpr int kThreads = 19 * 32;
}  // namespace fwd3

VBX_DEVINL void bar_sync_id(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
VBX_DEVINL void bar_arrive_id(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__global__ void __launch_bounds__(fwd3::kThreads, 1)
attn_fwd3_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                 const __grid_constant__ CUtensorMap mv, const uint8_t* __restrict__ key_mask, float scale_log2,
                 uint16_t* __restrict__ o, float* __restrict__ lse, int N, int H) {
  using namespace fwd3;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + NUM_BARS * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q_base = blockIdx.x * 2 * kBM;
  const bool two = q_base + kBM < N;                          // tile B holds at least one real query row
  const int nkv = (N + kBN - 1) / kBN;
  const int n_tail = (N - (nkv - 1) * kBN + 15) & ~15;        // GEMM width of the last key tile (16 .. 128)

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(&bars[Q_FULL], 1);
    for (int s_ = 0; s_ < kStages; ++s_) {
      mbar_init(&bars[K_FULL + s_], 1);
      mbar_init(&bars[K_EMPTY + s_], two ? 2 : 1);            // both groups' GEMMs read the stage
      mbar_init(&bars[V_FULL + s_], 1);
      mbar_init(&bars[V_EMPTY + s_], two ? 2 : 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&bars[S_FULL + g], 1);
      mbar_init(&bars[S_FREE + g], 256);
      mbar_init(&bars[P_FULL + g], 256);
      mbar_init(&bars[O_FULL + g], 1);
    }
    fence_barrier_init();
  }
  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&mq);
    tma_prefetch_desc(&mk);
    tma_prefetch_desc(&mv);
  }
  if (warp == kMmaWarp0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProducerWarp) {
    // ------------------------------------------------ TMA producer ------------------------------------------------
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars[Q_FULL], (two ? 2u : 1u) * kTileBytes);
      tma_load_4d(smem + kOffQ, &mq, &bars[Q_FULL], 0, q_base, h, b);
      if (two) tma_load_4d(smem + kOffQ + kTileBytes, &mq, &bars[Q_FULL], 0, q_base + kBM, h, b);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % kStages;
        const uint32_t ph = ((j / kStages) & 1) ^ 1;
        if (j + kStages < nkv) {  // warm L2 for the tile after the ring
          tma_prefetch_l2_4d(&mk, 0, (j + kStages) * kBN, h, b);
          tma_prefetch_l2_4d(&mv, 0, (j + kStages) * kBN, h, b);
        }
        mbar_wait(&bars[K_EMPTY + st], ph);
        mbar_arrive_expect_tx(&bars[K_FULL + st], kTileBytes);
        tma_load_4d(smem + kOffK + st * kTileBytes, &mk, &bars[K_FULL + st], 0, j * kBN, h, b);
        mbar_wait(&bars[V_EMPTY + st], ph);
        mbar_arrive_expect_tx(&bars[V_FULL + st], kTileBytes);
        tma_load_4d(smem + kOffV + st * kTileBytes, &mv, &bars[V_FULL + st], 0, j * kBN, h, b);
      }
    }
  } else if (warp >= kMmaWarp0) {
    // ------------------------------------------------ MMA issuer of group g ---------------------------------------
    const int g = warp - kMmaWarp0;
    if (g == 0 || two) {
      constexpr uint32_t idesc_s = make_idesc(kBM, kBN, false, false);  // S = Q K^T      (both K-major)
      constexpr uint32_t idesc_o = make_idesc(kBM, kDh, false, true);   // O += P V       (V is MN-major: [keys][d])
      const uint32_t tb = tmem_base + (uint32_t)g * kGroupCols;
      const uint64_t dQ = sdesc_k0(smem_u32(smem + kOffQ + g * kTileBytes));
      const uint64_t dK0 = sdesc_k0(smem_u32(smem + kOffK)), dV0 = sdesc_mn0(smem_u32(smem + kOffV));
      const bool leader = lane == 0;
      const uint32_t idesc_s_tail = make_idesc(kBM, n_tail, false, false);
      mbar_wait(&bars[Q_FULL], 0);
      auto issue_s = [&](int j) {  // S_j = Q K_j^T
        const int st = j % kStages;
        const uint64_t dK = dK0 + (uint64_t)st * (kTileBytes >> 4);
        const uint32_t idesc_sj = (j == nkv - 1) ? idesc_s_tail : idesc_s;
        mbar_wait(&bars[K_FULL + st], (j / kStages) & 1);
        mbar_wait(&bars[S_FREE + g], (j & 1) ^ 1);  // the softmax warps hold S_{j-1} in registers
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < kDh / 16; ++k) umma_bf16(tb, dQ + koff_k(k), dK + koff_k(k), idesc_sj, k > 0);
          umma_commit(&bars[S_FULL + g]);
          umma_commit(&bars[K_EMPTY + st]);
        }
        __syncwarp();
      };
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % kStages;
        const uint64_t dV = dV0 + (uint64_t)st * (kTileBytes >> 4);
        if (j + 1 < nkv) issue_s(j + 1);
        mbar_wait(&bars[V_FULL + st], (j / kStages) & 1);
        if (g == 0) TRACE(3, j, 4);
        mbar_wait(&bars[P_FULL + g], j & 1);
        if (g == 0) TRACE(3, j, 5);
        tc_fence_after();
        if (leader) {
          const int ksteps = (j == nkv - 1) ? n_tail / 16 : kBN / 16;
#pragma unroll
          for (int k = 0; k < kBN / 16; ++k)
            if (k < ksteps) umma_bf16_ts(tb + kColO, tb + kColP + k * 8, dV + koff_mn(k), idesc_o, (j > 0) || (k > 0));
          umma_commit(&bars[O_FULL + g]);
          umma_commit(&bars[V_EMPTY + st]);
        }
        __syncwarp();
        if (g == 0) TRACE(3, j, 6);
      }
    }
  } else {
    // ------------------------------------------------ softmax: group g = warp / 8, two threads per query row -------
    const int g = warp >> 3;
    if (g == 0 || two) {
      const int wl = warp & 7;
      const int half = wl >> 2;                   // which 64 key columns of S / which 32 columns of O
      const int quarter = wl & 3;                 // TMEM lane quarter (= warp % 4)
      const int r = quarter * 32 + lane;          // query row of the tile == TMEM lane
      const int q0 = q_base + g * kBM;
      const uint32_t t_lane = tmem_base + (uint32_t)g * kGroupCols + ((uint32_t)(quarter * 32) << 16);
      float* s_bias = reinterpret_cast<float*>(smem + kOffBias) + g * 2 * kBN;     // [2][128] per group
      float* s_max = reinterpret_cast<float*>(smem + kOffMax) + g * 2 * 2 * kBM;   // [2][2][128] per group
      const int bias_bar = 1 + g, pair_bar = 3 + g * 4 + quarter, tok_mine = 11 + g, tok_other = 12 - g;
      uint64_t* bS_FULL = &bars[S_FULL + g];
      uint64_t* bS_FREE = &bars[S_FREE + g];
      uint64_t* bP_FULL = &bars[P_FULL + g];
      uint64_t* bO_FULL = &bars[O_FULL + g];
      float m_used = -FLT_MAX, l = 0.f;
      if (two && g == 1) bar_arrive_id(11, 512);  // the token starts with group A

      for (int j = 0; j < nkv; ++j) {
        const int k0 = j * kBN;
        const bool masked_tile = (key_mask != nullptr) || (k0 + kBN > N);  // CTA-uniform
        const float* bias = s_bias + (j & 1) * kBN + half * 64;
        if (masked_tile) {
          if (half == 0) {
            const int key = k0 + r;
            float v = 0.f;
            if (key >= N) v = -INFINITY;
            else if (key_mask != nullptr && !key_mask[(int64_t)b * N + key]) v = -FLT_MAX;
            s_bias[(j & 1) * kBN + r] = v;
          }
          bar_sync_id(bias_bar, 256);
        }
        const int valid = N - k0;
        const bool live0 = half * 64 < valid, live1 = half * 64 + 32 < valid;   // warp-uniform
        if (threadIdx.x == 0) TRACE(4, j, 0);
        mbar_wait(bS_FULL, j & 1);
        if (threadIdx.x == 0) TRACE(4, j, 1);
        tc_fence_after();
        float s[64];
        if (live0) tmem_ld32(t_lane + half * 64, *reinterpret_cast<float(*)[32]>(&s[0]));
        if (live1) tmem_ld32(t_lane + half * 64 + 32, *reinterpret_cast<float(*)[32]>(&s[32]));
        tc_fence_before();
        mbar_arrive(bS_FREE);
        float mx = -FLT_MAX;
        if (masked_tile) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (!(c ? live1 : live0)) continue;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float bb = bias[c * 32 + i];
              float t = fmaf(s[c * 32 + i], scale_log2, bb);
              t = (bb == -INFINITY) ? -INFINITY : t;   // columns >= n_tail hold stale bits
              s[c * 32 + i] = t;
              mx = fmaxf(mx, t);
            }
          }
        } else {
          float m4[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
#pragma unroll
          for (int i = 0; i < 64; ++i) m4[i & 3] = fmaxf(m4[i & 3], s[i]);
          mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * scale_log2;
        }
        s_max[((j & 1) * 2 + half) * kBM + r] = mx;
        bar_sync_id(pair_bar, 64);
        mx = fmaxf(mx, s_max[((j & 1) * 2 + (half ^ 1)) * kBM + r]);
        bool o_ready = false;
        if (j == 0) {
          m_used = mx;
        } else {
          const float m_new = fmaxf(m_used, mx);
          const bool need = (m_new - m_used) > 8.0f;
          if (__any_sync(0xffffffffu, need)) {
            const float alpha = need ? ex2(m_used - m_new) : 1.0f;
            mbar_wait(bO_FULL, (j - 1) & 1);
            tc_fence_after();
            o_ready = true;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
              float ov[16];
              tmem_ld16(t_lane + kColO + half * 32 + c * 16, ov);
              uint32_t ou[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) ou[i] = __float_as_uint(ov[i] * alpha);
              tmem_st16(t_lane + kColO + half * 32 + c * 16, ou);
            }
            l *= alpha;
            if (need) m_used = m_new;
          }
        }
        const float neg_m = -m_used;
        if (threadIdx.x == 0) TRACE(4, j, 2);
        // ---- exponential phase: this group owns the MUFU pipe between token_wait and token_pass ----
        if (two) bar_sync_id(tok_mine, 512);
        if (threadIdx.x == 0) TRACE(4, j, 3);
        float rowsum = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (!(c ? live1 : live0)) continue;
          float* sc = s + c * 32;
          if (masked_tile) {
#pragma unroll
            for (int i = 0; i < 32; ++i) sc[i] = ex2(sc[i] + neg_m);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) sc[i] = VBX_EX2_AT(i, fmaf(sc[i], scale_log2, neg_m));
          }
          float r4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 32; ++i) r4[i & 3] += sc[i];
          rowsum += (r4[0] + r4[1]) + (r4[2] + r4[3]);
          uint32_t pk[16];
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            __nv_bfloat162 t2 = f2bf(sc[2 * x], sc[2 * x + 1]);
            pk[x] = *reinterpret_cast<uint32_t*>(&t2);
          }
          if (j > 0 && !o_ready) {   // P is single-buffered: P_{j-1} V_{j-1} must have read it (long done by now)
            mbar_wait(bO_FULL, (j - 1) & 1);
            tc_fence_after();
            o_ready = true;
          }
          tmem_st16(t_lane + kColP + half * 32 + c * 16, pk);
        }
        if (two) bar_arrive_id(tok_other, 512);
        if (threadIdx.x == 0) TRACE(4, j, 4);
        l += rowsum;
        if (j > 0 && !o_ready) {   // EVERY warp observes every O_FULL phase (see v2)
          mbar_wait(bO_FULL, (j - 1) & 1);
          tc_fence_after();
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bP_FULL);
        if (threadIdx.x == 0) TRACE(4, j, 5);
      }
      if (two && g == 0) bar_sync_id(11, 512);   // absorbs group B's last hand-over
      // epilogue
      mbar_wait(bO_FULL, (nkv - 1) & 1);
      tc_fence_after();
      float acc[32];
      tmem_ld32(t_lane + kColO + half * 32, acc);
      tc_fence_before();
      s_max[((nkv & 1) * 2 + half) * kBM + r] = l;
      bar_sync_id(pair_bar, 64);
      l += s_max[((nkv & 1) * 2 + (half ^ 1)) * kBM + r];
      const int q = q0 + r;
      if (q < N) {
        const float inv_l = 1.0f / l;
        uint16_t* dst = o + (((int64_t)b * N + q) * H + h) * kDh + half * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float t[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = acc[c * 8 + i] * inv_l;
          stg_16(dst + c * 8, pack8(t));
        }
        if (lse != nullptr && half == 0) lse[((int64_t)b * H + h) * N + q] = m_used + log2f(l);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == fwd3::kMmaWarp0) tmem_dealloc(tmem_base, 512);
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
constexpr int kStages = 3;  // Q/dO ring: the load of tile i+2 is issued as soon as the GEMMs of tile i-1 retire
constexpr int kDsBufs = VBX_EXP_DSBUF ? 2 : 1;                   // dS^T tiles in shared memory
constexpr uint32_t kDqStageBytes = VBX_EXP_DSBUF ? 16384 : 32768;  // dQ staging: one 32-column half, or both
constexpr uint32_t kOffK = 0, kOffV = 16384, kOffQ = 32768, kOffdO = kOffQ + kStages * 16384,
                   kOffdST = kOffdO + kStages * 16384, kOffdQ = kOffdST + kDsBufs * 32768, kOffBar = kOffdQ + kDqStageBytes,
                   kOffStat = kOffBar + 256;  // per compute warp, double buffered: [16 warps][2][32 -lse | 32 delta] f32
constexpr int kComputeWarps = 16;           // 4 threads per key row: 32 query columns of S^T / dP^T each
constexpr int kProducerWarp = 16, kMmaWarp = 17, kFlushWarp0 = 18, kMmaWarpB = 22, kMmaWarpC = 23;  // B, C: VBX_BWD_ISSUERS == 3
constexpr uint32_t kSmemBytes = kOffStat + kComputeWarps * 2 * 64 * 4;  // ~216 KB: one CTA per SM (TMEM: all 512 columns)
static_assert(kSmemBytes <= 232448, "shared memory budget");
enum { KV_FULL = 0, QD_FULL = 1, QD_EMPTY = 4, ST_FULL = 7, ST_FREE = 8, DS_FULL = 9, DQ_FULL = 10, DQ_FREE = 11,
       PT_FREE = 12,   // (VBX_EXP_DSBUF) dV_i has retired: P^T may be overwritten, and so may the dS^T buffer of tile i-1
       ALL_DONE = 13,  // (VBX_EXP_DSBUF) every dV / dK GEMM of the CTA has retired
       DSB_FREE = 14,  // (3 issuers) [2]: dK_i AND dQ_i have retired: dS^T buffer i%2 may be overwritten (one barrier per buffer,
                       // so a waiter is never two phases behind)
       NUM_BARS = 16 };
static_assert(NUM_BARS * 8 + 4 <= 256, "barrier block");
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColST = 0, kColDPT = 128, kColDV = 256, kColDK = 320, kColDQ = 384, kColPT = 448;  // P^T: bf16 pairs
constexpr int kThreads = VBX_BWD_ISSUERS == 3 ? 768 : 704;  // warps 0-15 compute, 16 TMA producer, 17 MMA issuer (A), 18-21 dQ flush,
                                                            // 22-23 MMA issuers B and C
}  // namespace bwd

// One CTA per (key tile, head, batch); loops over the query tiles.  Everything is computed TRANSPOSED (keys on the
// TMEM lanes) so that P^T and dS^T come out K-major for the dV / dK GEMMs, which accumulate in TMEM across the loop:
//   S^T = K Q^T, dP^T = V dO^T  ->  P^T = 2^(c S^T - lse), dS^T = P^T (dP^T - delta)
//   dV += P^T dO,  dK += dS^T Q,  dQ_i = dS K  (fp32, scaled by `scale`, staged in shared memory and added into dq by
//   ONE TMA reduce-add per 128x32 block -- per-lane red.global atomics cost ~8000 clk per tile, 6x the five GEMMs)
// 22 warps: 0-15 compute (key row 32*(w%4)+lane, query-column quarter w/4: FOUR threads per row -> 4 warps per SM
// sub-partition hide the ex2 / TMEM / shared-memory latencies of the exp phase), 16 TMA producer, 17 MMA issuer, 18-21 dQ
// flush (TMEM -> shared -> TMA reduce-add, off the compute warps' critical path).  Software pipeline: the S^T/dP^T GEMMs of tile i+1 are issued BEFORE the
// dV/dK/dQ GEMMs of tile i, and the compute warps keep P^T/dS^T of tile i+1 in registers until those GEMMs have finished
// reading the shared-memory operand tiles -- so tensor pipe and exp/FMA pipes overlap instead of alternating.
__global__ void __launch_bounds__(bwd::kThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                const __grid_constant__ CUtensorMap mv, const __grid_constant__ CUtensorMap mdo,
                const __grid_constant__ CUtensorMap mdq, const uint8_t* __restrict__ key_mask, float scale, float scale_log2,
                const float* __restrict__ lse, const float* __restrict__ delta, uint16_t* __restrict__ dk,
                uint16_t* __restrict__ dv, int64_t dv_bs, int64_t dv_ns, int N, int H) {
  using namespace bwd;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + NUM_BARS * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * kBN, h = blockIdx.y, b = blockIdx.z;
  const int nq = (N + kBM - 1) / kBM;
  const int64_t bh = (int64_t)b * H + h;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(&bars[KV_FULL], 1);
    for (int s_ = 0; s_ < kStages; ++s_) {
      mbar_init(&bars[QD_FULL + s_], 1);
      mbar_init(&bars[QD_EMPTY + s_], VBX_BWD_ISSUERS == 3 ? 2 : 1);   // 3 issuers: warp A's and warp B's GEMMs both read the stage
    }
    mbar_init(&bars[ST_FULL], 1);
    mbar_init(&bars[ST_FREE], kComputeWarps * 32);
    mbar_init(&bars[DS_FULL], kComputeWarps * 32);
    mbar_init(&bars[DQ_FULL], 1);
    mbar_init(&bars[DQ_FREE], 128);
    if (VBX_EXP_DSBUF) {
      mbar_init(&bars[PT_FREE], 1);
      mbar_init(&bars[ALL_DONE], 1);
    }
    mbar_init(&bars[DSB_FREE], 2);
    mbar_init(&bars[DSB_FREE + 1], 2);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProducerWarp) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars[KV_FULL], 2 * kTileBytes);
      tma_load_4d(smem + kOffK, &mk, &bars[KV_FULL], 0, k0, h, b);
      tma_load_4d(smem + kOffV, &mv, &bars[KV_FULL], 0, k0, h, b);
      for (int i = 0; i < nq; ++i) {
        const int st = i % kStages;
        if (i + kStages < nq) {  // warm L2 for the tile after the ring
          tma_prefetch_l2_4d(&mq, 0, (i + kStages) * kBM, h, b);
          tma_prefetch_l2_4d(&mdo, 0, (i + kStages) * kBM, h, b);
        }
        mbar_wait(&bars[QD_EMPTY + st], ((i / kStages) & 1) ^ 1);
        TRACE(2, i, 0);
        mbar_arrive_expect_tx(&bars[QD_FULL + st], 2 * kTileBytes);
        tma_load_4d(smem + kOffQ + st * kTileBytes, &mq, &bars[QD_FULL + st], 0, i * kBM, h, b);
        tma_load_4d(smem + kOffdO + st * kTileBytes, &mdo, &bars[QD_FULL + st], 0, i * kBM, h, b);
      }
    }
#if VBX_BWD_ISSUERS == 1
  } else if (warp == kMmaWarp) {
    // warp-uniform control flow, hoisted descriptors (see the forward kernel): only the leader lane issues MMAs / commits
    {
      constexpr uint32_t idesc_kk = make_idesc(128, 128, false, false);
      constexpr uint32_t idesc_kmn = make_idesc(128, kDh, false, true);
      constexpr uint32_t idesc_mnmn = make_idesc(128, kDh, true, true);
      const bool leader = lane == 0;
      const uint64_t dKk = sdesc_k0(smem_u32(smem + kOffK)), dKmn = sdesc_mn0(smem_u32(smem + kOffK));
      const uint64_t dVk = sdesc_k0(smem_u32(smem + kOffV));
      const uint64_t dQk0 = sdesc_k0(smem_u32(smem + kOffQ)), dQmn0 = sdesc_mn0(smem_u32(smem + kOffQ));
      const uint64_t dOk0 = sdesc_k0(smem_u32(smem + kOffdO)), dOmn0 = sdesc_mn0(smem_u32(smem + kOffdO));
      const uint64_t dSk = sdesc_k0(smem_u32(smem + kOffdST)), dSmn = sdesc_mn0(smem_u32(smem + kOffdST));
      const int n_qt = VBX_EXP_TAIL ? ((N - (nq - 1) * kBM + 15) & ~15) : kBM;  // GEMM width of the last query tile
      const uint32_t idesc_kk_tail = make_idesc(128, n_qt, false, false);
      auto issue_s = [&](int i) {  // S^T = K Q_i^T, dP^T = V dO_i^T
        const uint64_t so = (uint64_t)(i % kStages) * (kTileBytes >> 4);
        const uint32_t idesc_i = (VBX_EXP_TAIL && i == nq - 1) ? idesc_kk_tail : idesc_kk;
        TRACE(0, i, 0);
        mbar_wait(&bars[QD_FULL + i % kStages], (i / kStages) & 1);
        TRACE(0, i, 1);
        mbar_wait(&bars[ST_FREE], (i & 1) ^ 1);  // the compute warps hold tile i-1's S^T/dP^T in registers
        TRACE(0, i, 2);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < kDh / 16; ++k)
            umma_bf16(tmem_base + kColST, dKk + koff_k(k), dQk0 + so + koff_k(k), idesc_i, k > 0);
#pragma unroll
          for (int k = 0; k < kDh / 16; ++k)
            umma_bf16(tmem_base + kColDPT, dVk + koff_k(k), dOk0 + so + koff_k(k), idesc_i, k > 0);
          umma_commit(&bars[ST_FULL]);
        }
        __syncwarp();
        TRACE(0, i, 3);
      };
      mbar_wait(&bars[KV_FULL], 0);
      issue_s(0);
      for (int i = 0; i < nq; ++i) {
        const int st = i % kStages;
        const uint64_t so = (uint64_t)st * (kTileBytes >> 4);
        const uint64_t dso = (uint64_t)(i & (kDsBufs - 1)) * (32768 >> 4);  // which dS^T buffer (VBX_EXP_DSBUF)
        if (i + 1 < nq) issue_s(i + 1);  // runs on the tensor pipe while the compute warps finish tile i
        TRACE(0, i, 4);
        mbar_wait(&bars[DS_FULL], i & 1);
        TRACE(0, i, 5);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < kBM / 16; ++k)  // dV += P^T dO   (TS mode: P^T read from TMEM, only dO from shared memory)
            if (!VBX_EXP_TAIL || i < nq - 1 || k * 16 < n_qt)
              umma_bf16_ts(tmem_base + kColDV, tmem_base + kColPT + k * 8, dOmn0 + so + koff_mn(k), idesc_kmn, (i > 0) || (k > 0));
          if (VBX_EXP_DSBUF) umma_commit(&bars[PT_FREE]);  // dV_i (and everything before it) retired -> P^T is free
#pragma unroll
          for (int k = 0; k < kBM / 16; ++k)  // dK += dS^T Q
            if (!VBX_EXP_TAIL || i < nq - 1 || k * 16 < n_qt)
              umma_bf16(tmem_base + kColDK, dSk + dso + koff_k(k), dQmn0 + so + koff_mn(k), idesc_kmn, (i > 0) || (k > 0));
        }
        __syncwarp();
        mbar_wait(&bars[DQ_FREE], (i & 1) ^ 1);  // the flush warps have read dQ_{i-1} out of TMEM
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < kBN / 16; ++k)  // dQ_i = dS K    (A = dS^T read MN-major)
            umma_bf16(tmem_base + kColDQ, dSmn + dso + koff_mn(k), dKmn + koff_mn(k), idesc_mnmn, k > 0);
          umma_commit(&bars[DQ_FULL]);
          umma_commit(&bars[QD_EMPTY + st]);
          if (VBX_EXP_DSBUF && i == nq - 1) umma_commit(&bars[ALL_DONE]);
        }
        __syncwarp();
        TRACE(0, i, 6);
      }
    }
#else
  } else if (warp == kMmaWarp || warp == kMmaWarpB || warp == kMmaWarpC) {
    // ---- three issuing warps (warp-uniform control flow; the leader lane issues).  tcgen05.mma's of different threads are not
    // ordered against each other, so EVERY dependency between the three streams goes through an mbarrier, and every barrier
    // that guards a resource read by two streams counts both commits (QD_EMPTY: A and B; DSB_FREE: B and C). ----
    constexpr uint32_t idesc_kk = make_idesc(128, 128, false, false);
    constexpr uint32_t idesc_kmn = make_idesc(128, kDh, false, true);
    constexpr uint32_t idesc_mnmn = make_idesc(128, kDh, true, true);
    const bool leader = lane == 0;
    const int n_qt = VBX_EXP_TAIL ? ((N - (nq - 1) * kBM + 15) & ~15) : kBM;  // GEMM width of the last query tile
    if (warp == kMmaWarp) {
      // A: S^T = K Q_i^T, dP^T = V dO_i^T for every tile, as early as the compute warps have pulled tile i-1 out of TMEM
      const uint64_t dKk = sdesc_k0(smem_u32(smem + kOffK)), dVk = sdesc_k0(smem_u32(smem + kOffV));
      const uint64_t dQk0 = sdesc_k0(smem_u32(smem + kOffQ)), dOk0 = sdesc_k0(smem_u32(smem + kOffdO));
      const uint32_t idesc_kk_tail = make_idesc(128, n_qt, false, false);
      mbar_wait(&bars[KV_FULL], 0);
      for (int i = 0; i < nq; ++i) {
        const int st = i % kStages;
        const uint64_t so = (uint64_t)st * (kTileBytes >> 4);
        const uint32_t idesc_i = (VBX_EXP_TAIL && i == nq - 1) ? idesc_kk_tail : idesc_kk;
        TRACE(0, i, 0);
        mbar_wait(&bars[QD_FULL + st], (i / kStages) & 1);
        TRACE(0, i, 1);
        mbar_wait(&bars[ST_FREE], (i & 1) ^ 1);
        TRACE(0, i, 2);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < kDh / 16; ++k)
            umma_bf16(tmem_base + kColST, dKk + koff_k(k), dQk0 + so + koff_k(k), idesc_i, k > 0);
#pragma unroll
          for (int k = 0; k < kDh / 16; ++k)
            umma_bf16(tmem_base + kColDPT, dVk + koff_k(k), dOk0 + so + koff_k(k), idesc_i, k > 0);
          umma_commit(&bars[ST_FULL]);
          umma_commit(&bars[QD_EMPTY + st]);      // this warp's share: its GEMMs have read Q_i / dO_i
        }
        __syncwarp();
        TRACE(0, i, 3);
      }
    } else if (warp == kMmaWarpB) {
      // B: dV += P^T dO (TS mode), dK += dS^T Q
      const uint64_t dQmn0 = sdesc_mn0(smem_u32(smem + kOffQ)), dOmn0 = sdesc_mn0(smem_u32(smem + kOffdO));
      const uint64_t dSk = sdesc_k0(smem_u32(smem + kOffdST));
      for (int i = 0; i < nq; ++i) {
        const int st = i % kStages;
        const uint64_t so = (uint64_t)st * (kTileBytes >> 4);
        const uint64_t dso = (uint64_t)(i & 1) * (32768 >> 4);
        mbar_wait(&bars[QD_FULL + st], (i / kStages) & 1);   // (long complete: makes the TMA writes visible to this thread)
        TRACE(5, i, 0);
        mbar_wait(&bars[DS_FULL], i & 1);
        TRACE(5, i, 1);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < kBM / 16; ++k)
            if (!VBX_EXP_TAIL || i < nq - 1 || k * 16 < n_qt)
              umma_bf16_ts(tmem_base + kColDV, tmem_base + kColPT + k * 8, dOmn0 + so + koff_mn(k), idesc_kmn, (i > 0) || (k > 0));
          umma_commit(&bars[PT_FREE]);            // dV_i retired -> P^T (TMEM) may be overwritten
#pragma unroll
          for (int k = 0; k < kBM / 16; ++k)
            if (!VBX_EXP_TAIL || i < nq - 1 || k * 16 < n_qt)
              umma_bf16(tmem_base + kColDK, dSk + dso + koff_k(k), dQmn0 + so + koff_mn(k), idesc_kmn, (i > 0) || (k > 0));
          umma_commit(&bars[QD_EMPTY + st]);      // this warp's share
          umma_commit(&bars[DSB_FREE + (i & 1)]);
          if (i == nq - 1) umma_commit(&bars[ALL_DONE]);
        }
        __syncwarp();
        TRACE(5, i, 2);
      }
    } else {
      // C: dQ_i = dS K (A = dS^T read MN-major)
      const uint64_t dKmn = sdesc_mn0(smem_u32(smem + kOffK)), dSmn = sdesc_mn0(smem_u32(smem + kOffdST));
      mbar_wait(&bars[KV_FULL], 0);
      for (int i = 0; i < nq; ++i) {
        const uint64_t dso = (uint64_t)(i & 1) * (32768 >> 4);
        TRACE(6, i, 0);
        mbar_wait(&bars[DS_FULL], i & 1);
        TRACE(6, i, 1);
        mbar_wait(&bars[DQ_FREE], (i & 1) ^ 1);   // the flush warps have read dQ_{i-1} out of TMEM
        TRACE(6, i, 2);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < kBN / 16; ++k)
            umma_bf16(tmem_base + kColDQ, dSmn + dso + koff_mn(k), dKmn + koff_mn(k), idesc_mnmn, k > 0);
          umma_commit(&bars[DQ_FULL]);
          umma_commit(&bars[DSB_FREE + (i & 1)]);
        }
        __syncwarp();
        TRACE(6, i, 3);
      }
    }
#endif
  } else if (warp >= kFlushWarp0 && warp < kFlushWarp0 + 4) {
    // ------------------------------------------------ dQ flush warps -----------------------------------------------
    // dQ_i: TMEM -> registers -> shared -> TMA reduce-add into dq.  Warp w owns query rows [32(w%4), +32): a contiguous,
    // 1024-byte aligned 4 KB slice of each [128 rows][32 fp32] SWIZZLE_128B staging block, reduced with its own 32x32 TMA
    // ops (bulk groups are per thread: lane 0 waits for its previous group before the slice is overwritten).
    const int r = (warp & 3) * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
#if VBX_EXP_DSBUF
    // one 16 KB staging block: the two 32-column halves of dQ_i go through this warp's 4 KB slice one after the other
    for (int i = 0; i < nq; ++i) {
      mbar_wait(&bars[DQ_FULL], i & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (lane == 0) tma_wait_group_read0();  // the previous reduce-add has read this slice
        __syncwarp();
        float v[32];
        tmem_ld32(t_lane + kColDQ + c * 32, v);
        if (c == 1) {
          tc_fence_before();
          mbar_arrive(&bars[DQ_FREE]);
        }
#pragma unroll
        for (int qd = 0; qd < 8; ++qd) {
          float4 o4 = make_float4(v[qd * 4] * scale, v[qd * 4 + 1] * scale, v[qd * 4 + 2] * scale, v[qd * 4 + 3] * scale);
          *reinterpret_cast<float4*>(smem + kOffdQ + r * 128 + ((qd ^ (r & 7)) << 4)) = o4;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_4d(&mdq, smem + kOffdQ + (warp & 3) * 4096, c * 32, i * kBM + (warp & 3) * 32, h, b);
          tma_commit_group();
        }
      }
    }
#else
    for (int i = 0; i < nq; ++i) {
      mbar_wait(&bars[DQ_FULL], i & 1);
      tc_fence_after();
      if (lane == 0) tma_wait_group_read0();
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float v[32];
        tmem_ld32(t_lane + kColDQ + c * 32, v);
        if (c == 1) {  // dQ_i is in registers: the MMA warp may start dQ_{i+1}
          tc_fence_before();
          mbar_arrive(&bars[DQ_FREE]);
        }
#pragma unroll
        for (int qd = 0; qd < 8; ++qd) {
          float4 o4 = make_float4(v[qd * 4] * scale, v[qd * 4 + 1] * scale, v[qd * 4 + 2] * scale, v[qd * 4 + 3] * scale);
          *reinterpret_cast<float4*>(smem + kOffdQ + c * kTileBytes + r * 128 + ((qd ^ (r & 7)) << 4)) = o4;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {  // rows beyond N are clipped by the tensor map
        tma_reduce_add_4d(&mdq, smem + kOffdQ + (warp & 3) * 4096, 0, i * kBM + (warp & 3) * 32, h, b);
        tma_reduce_add_4d(&mdq, smem + kOffdQ + kTileBytes + (warp & 3) * 4096, 32, i * kBM + (warp & 3) * 32, h, b);
        tma_commit_group();
      }
    }
#endif
    if (lane == 0) tma_wait_group0();
  } else {
    // ------------------------------------------------ compute warps ------------------------------------------------
    const int qr = warp >> 2;                // which 32 query columns of S^T / dP^T, which 16 columns of dV / dK
    const int r = (warp & 3) * 32 + lane;    // TMEM lane: key row
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int key = k0 + r;
    const bool dead_row = (key >= N) || (key_mask != nullptr && !key_mask[(int64_t)b * N + key]);  // P^T row is zero
    const bool warp_dead = VBX_EXP_TAIL && (k0 + (warp & 3) * 32) >= N;  // all 32 key rows of this warp are padding
    // -lse / delta of this warp's 32 query columns live in a warp-private, double-buffered shared slice: lane l loads
    // column l one tile ahead (register prefetch) and the warp only needs __syncwarp -- no block-wide barrier.
    float* stat = reinterpret_cast<float*>(smem + kOffStat) + warp * 2 * 64;
    auto load_stats = [&](int i, float& nl, float& dl) {
      const int q = i * kBM + qr * 32 + lane;
      nl = -INFINITY;  // -lse = -inf -> p = 0 for padded queries
      dl = 0.f;
      if (i < nq && q < N) nl = -lse[bh * N + q], dl = delta[bh * N + q];
    };
    float nl, dl;
    load_stats(0, nl, dl);

    for (int i = 0; i < nq; ++i) {
      float* lrow_w = stat + (i & 1) * 64;
      lrow_w[lane] = nl;
      lrow_w[32 + lane] = dl;
      load_stats(i + 1, nl, dl);  // next tile: latency hides behind this tile's math
      __syncwarp();
      if (threadIdx.x == 0) TRACE(1, i, 1);
      mbar_wait(&bars[ST_FULL], i & 1);
      if (threadIdx.x == 0) TRACE(1, i, 2);
      tc_fence_after();
      uint32_t pk[16], dsk[16];  // P^T / dS^T of this thread's 32 columns, packed bf16x2, held until smem / TMEM are free
      const int valid_q = N - i * kBM;  // >= 128 except in the tail query tile
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (VBX_EXP_TAIL && (qr * 32 + c * 16 >= valid_q || warp_dead)) {
          // warp-uniform: every query of this 16-column chunk (columns the narrowed GEMMs did not even write), or every key
          // row of this warp, is padding: exact zeros without touching TMEM or the exp unit
          if (c == 1) {
            tc_fence_before();
            mbar_arrive(&bars[ST_FREE]);
          }
#pragma unroll
          for (int x = 0; x < 8; ++x) pk[c * 8 + x] = dsk[c * 8 + x] = 0u;
          continue;
        }
        float s[16], dp[16];
        tmem_ld16(t_lane + kColST + qr * 32 + c * 16, s);
        tmem_ld16(t_lane + kColDPT + qr * 32 + c * 16, dp);
        if (c == 1) {
          tc_fence_before();
          mbar_arrive(&bars[ST_FREE]);
        }
        const float* lrow = lrow_w + c * 16;       // -lse
        const float* drow = lrow_w + 32 + c * 16;  // delta
#if VBX_FWD_F32X2
        // packed fp32 (FFMA2 / FADD2 / FMUL2): same roundings as the scalar lines, three issue slots per pair fewer.  The 32 -lse /
        // delta values of this warp's query columns are the same for every lane: 16-byte broadcast loads (one shared-memory
        // wavefront per four columns; the 8-byte version cost 512 wavefronts per tile of a kernel bound by shared-memory bandwidth)
#pragma unroll
        for (int x = 0; x < 16; x += 4) {
          const float4 lr = *reinterpret_cast<const float4*>(lrow + x), dr = *reinterpret_cast<const float4*>(drow + x);
          const float2 a01 = __ffma2_rn(make_float2(s[x], s[x + 1]), make_float2(scale_log2, scale_log2), make_float2(lr.x, lr.y));
          const float2 a23 = __ffma2_rn(make_float2(s[x + 2], s[x + 3]), make_float2(scale_log2, scale_log2), make_float2(lr.z, lr.w));
          // padded queries (-lse = -inf) may come out as 2^-125 instead of 0 from the polynomial: they only ever multiply
          // zero-filled dO / Q rows and clipped dQ rows
          const float p0 = VBX_EX2_AT(x, a01.x), p1 = VBX_EX2_AT(x + 1, a01.y), p2 = VBX_EX2_AT(x + 2, a23.x), p3 = VBX_EX2_AT(x + 3, a23.y);
          const float2 d01 = __fmul2_rn(make_float2(p0, p1), __fadd2_rn(make_float2(dp[x], dp[x + 1]), make_float2(-dr.x, -dr.y)));
          const float2 d23 = __fmul2_rn(make_float2(p2, p3), __fadd2_rn(make_float2(dp[x + 2], dp[x + 3]), make_float2(-dr.z, -dr.w)));
          __nv_bfloat162 pa = f2bf(p0, p1), pb = f2bf(p2, p3), da = f2bf(d01.x, d01.y), db = f2bf(d23.x, d23.y);
          pk[c * 8 + (x >> 1)] = *reinterpret_cast<uint32_t*>(&pa);
          pk[c * 8 + (x >> 1) + 1] = *reinterpret_cast<uint32_t*>(&pb);
          dsk[c * 8 + (x >> 1)] = *reinterpret_cast<uint32_t*>(&da);
          dsk[c * 8 + (x >> 1) + 1] = *reinterpret_cast<uint32_t*>(&db);
        }
#else
#pragma unroll
        for (int x = 0; x < 16; x += 2) {
          // padded queries (-lse = -inf) may come out as 2^-125 instead of 0 from the polynomial: they only ever multiply
          // zero-filled dO / Q rows and clipped dQ rows
          const float p0 = VBX_EX2_AT(x, fmaf(s[x], scale_log2, lrow[x]));
          const float p1 = VBX_EX2_AT(x + 1, fmaf(s[x + 1], scale_log2, lrow[x + 1]));
          const float d0 = p0 * (dp[x] - drow[x]);
          const float d1 = p1 * (dp[x + 1] - drow[x + 1]);
          __nv_bfloat162 pp = f2bf(p0, p1), dd = f2bf(d0, d1);
          pk[c * 8 + (x >> 1)] = *reinterpret_cast<uint32_t*>(&pp);
          dsk[c * 8 + (x >> 1)] = *reinterpret_cast<uint32_t*>(&dd);
        }
#endif
      }
      if (dead_row) {  // rare (tail tile / masked keys): the whole row of P^T and dS^T is zero
#pragma unroll
        for (int x = 0; x < 16; ++x) pk[x] = dsk[x] = 0u;
      }
      if (threadIdx.x == 0) TRACE(1, i, 3);
      if (i > 0) {
        // shipped: ALL of tile i-1's GEMMs have retired: P^T (TMEM) and the single dS^T tile (shared) may be overwritten.
        // VBX_EXP_DSBUF: dV_{i-1} has retired (so has every GEMM of tile i-2): P^T and dS^T buffer i%2 may be overwritten.
        mbar_wait(&bars[VBX_EXP_DSBUF ? PT_FREE : DQ_FULL], (i - 1) & 1);
#if VBX_BWD_ISSUERS == 3
        // separate issuing warps: dV_{i-1} retiring no longer implies that dK_{i-2} / dQ_{i-2} (which read dS^T buffer i%2) have
        if (i >= 2) mbar_wait(&bars[DSB_FREE + (i & 1)], ((i - 2) >> 1) & 1);
#endif
        tc_fence_after();
      }
      if (threadIdx.x == 0) TRACE(1, i, 4);
      // P^T -> TMEM (A operand of the TS-mode dV GEMM: never touches shared memory); dS^T -> shared (K-major for dK, and
      // read MN-major for dQ).  This thread's 32 query columns = 4 16-byte chunks of 64-query sub-tile qr/2.
      tmem_st16(t_lane + kColPT + qr * 16, pk);
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        const int cc = (qr & 1) * 4 + qd;
        const uint32_t off = (i & (kDsBufs - 1)) * 32768 + (qr >> 1) * kSubTileBytes + r * 128 + ((cc ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(smem + kOffdST + off) = make_uint4(dsk[qd * 4], dsk[qd * 4 + 1], dsk[qd * 4 + 2], dsk[qd * 4 + 3]);
      }
      tmem_st_wait();
      tc_fence_before();
      fence_proxy_async();
      if (threadIdx.x == 0) TRACE(1, i, 5);
      mbar_arrive(&bars[DS_FULL]);
      if (threadIdx.x == 0) TRACE(1, i, 6);
    }
    if (VBX_EXP_DSBUF) mbar_wait(&bars[ALL_DONE], 0);  // these warps did not follow DQ_FULL phase by phase
    else mbar_wait(&bars[DQ_FULL], (nq - 1) & 1);
    tc_fence_after();
    // all GEMMs have retired: write this thread's 16 columns of dV and dK for its key row.  The TMEM loads are
    // .sync.aligned (whole warp, converged); only the global stores are predicated on the key being real.
    {
      const bool live = key < N;
      uint16_t* dvp = dv + (int64_t)b * dv_bs + (int64_t)(live ? key : 0) * dv_ns + h * kDh + qr * 16;
      uint16_t* dkp = dk + qk_vec_off(b, h, live ? key : 0, H, N) + qr * 16;
      float v[16];
      tmem_ld16(t_lane + kColDV + qr * 16, v);
      if (live) {
        stg_16(dvp, pack8(&v[0]));
        stg_16(dvp + 8, pack8(&v[8]));
      }
      tmem_ld16(t_lane + kColDK + qr * 16, v);
#pragma unroll
      for (int x = 0; x < 16; ++x) v[x] *= scale;
      if (live) {
        stg_16(dkp, pack8(&v[0]));
        stg_16(dkp + 8, pack8(&v[8]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == bwd::kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

// =====================================================================================================================
// tcgen05 / TMA self-test: C[128x128] = A B^T with K = 128, every operand-major combination, manual or TMA staging
// =====================================================================================================================
template <bool kTma>
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mb,
                     const uint16_t* __restrict__ A, const uint16_t* __restrict__ Bm, float* __restrict__ C, int a_mn, int b_mn,
                     int a_tmem) {
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
    tmem_alloc(tmem_slot, 256);
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
  if (a_tmem) {
    // TS mode: row m of A -> TMEM lane m, columns [128, 192): column c holds elements (2c, 2c+1) of the row
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16) + 128;
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(A + threadIdx.x * 128);
    for (int c = 0; c < 4; ++c) {
      uint32_t v[16];
      for (int x = 0; x < 16; ++x) v[x] = arow[c * 16 + x];
      tmem_st16(t_row + c * 16, v);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc(128, 128, a_mn != 0, b_mn != 0);
    const uint32_t aA = smem_u32(sA), aB = smem_u32(sB);
    for (int k = 0; k < 8; ++k) {
      if (a_tmem)
        umma_bf16_ts(tmem_base, tmem_base + 128 + k * 8, b_mn ? desc_mnmajor(aB, k) : desc_kmajor(aB, k), idesc, k > 0);
      else
        umma_bf16(tmem_base, a_mn ? desc_mnmajor(aA, k) : desc_kmajor(aA, k),
                  b_mn ? desc_mnmajor(aB, k) : desc_kmajor(aB, k), idesc, k > 0);
    }
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
  if (warp == 0) tmem_dealloc(tmem_base, 256);
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

// esize = 2 (bf16, box 64 wide) or 4 (fp32, box 32 wide): either way one box row is 128 bytes = one swizzle row
static int make_tmap(CUtensorMap* out, const void* base, int64_t inner, int64_t N, int64_t H, int64_t B, int64_t n_stride,
                     int64_t h_stride, int64_t b_stride, int box_rows, int esize = 2) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return VBX_E_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (n_stride * esize) % 16 || (h_stride * esize) % 16 || (b_stride * esize) % 16)
    return VBX_E_ALIGN;
  cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)N, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)n_stride * esize, (cuuint64_t)h_stride * esize, (cuuint64_t)b_stride * esize};
  cuuint32_t box[4] = {(cuuint32_t)(128 / esize), (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult rc = fn(out, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                   const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? VBX_OK : VBX_E_DRIVER;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, int64_t inner, int64_t rows, int64_t row_pitch, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return VBX_E_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_pitch * 2) % 16 || inner <= 0 || rows <= 0 || box_rows <= 0 || box_rows > 256)
    return VBX_E_ALIGN;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_pitch * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
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
// (n, h, b) element strides of q^ / k^ / dq^ in the tensor maps (common.cuh: VBX_QK_TOKEN_MAJOR)
#if VBX_QK_TOKEN_MAJOR
#define VBX_QK_STRIDES(N, H) (H) * kDh, kDh, (N) * (H) * kDh
#else
#define VBX_QK_STRIDES(N, H) kDh, (N) * kDh, (H) * (N) * kDh
#endif

extern "C" int vbx_attn_fwd(const uint16_t* q, const uint16_t* k, const uint16_t* v, int64_t v_bs, int64_t v_ns,
                            const uint8_t* key_mask, float scale, uint16_t* o, float* lse, int64_t B, int64_t H, int64_t N,
                            void* stream) {
  VBX_REQUIRE(q && k && v && o, VBX_E_NULL);
  VBX_REQUIRE(B > 0 && H > 0 && N > 0 && B < 65536 && H < 65536 && N < (1 << 24), VBX_E_SHAPE);
  VBX_REQUIRE(VBX_ALIGNED16(o), VBX_E_ALIGN);
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = make_tmap_bf16_4d(&mq, q, N, H, B, VBX_QK_STRIDES(N, H), kBM)) != VBX_OK) return rc;
  if ((rc = make_tmap_bf16_4d(&mk, k, N, H, B, VBX_QK_STRIDES(N, H), kBN)) != VBX_OK) return rc;
  if ((rc = make_tmap_bf16_4d(&mv, v, N, H, B, v_ns, kDh, v_bs, kBN)) != VBX_OK) return rc;
  // per call, not once per process: the attribute is per device, and a host process may drive several GPUs
  cudaError_t ce = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd::kSmemBytes);
  if (ce != cudaSuccess) return (int)ce;
  dim3 grid((unsigned)((N + kBM - 1) / kBM), (unsigned)H, (unsigned)B);
  static const int fwd_ver = [] {   // VBX_ATTN_FWD = 1 | 2 | 3 selects the forward kernel generation (A/B runs); default below
    const char* e = getenv("VBX_ATTN_FWD");
    if (getenv("VBX_ATTN_FWD_V1") != nullptr && getenv("VBX_ATTN_FWD_V1")[0] == '1') return 1;
    return (e != nullptr && e[0] >= '1' && e[0] <= '3') ? e[0] - '0' : VBX_ATTN_FWD_DEFAULT;
  }();
  if (fwd_ver == 1) {
    attn_fwd_kernel<<<grid, fwd::kThreads, fwd::kSmemBytes, (cudaStream_t)stream>>>(mq, mk, mv, key_mask, scale * kLog2e, o, lse, (int)N,
                                                                         (int)H);
    return VBX_LAUNCH_RC();
  }
  if (fwd_ver == 3) {
    ce = cudaFuncSetAttribute(attn_fwd3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd3::kSmemBytes);
    if (ce != cudaSuccess) return (int)ce;
    dim3 grid3((unsigned)((N + 2 * kBM - 1) / (2 * kBM)), (unsigned)H, (unsigned)B);
    attn_fwd3_kernel<<<grid3, fwd3::kThreads, fwd3::kSmemBytes, (cudaStream_t)stream>>>(mq, mk, mv, key_mask, scale * kLog2e, o, lse,
                                                                                 (int)N, (int)H);
    return VBX_LAUNCH_RC();
  }
  ce = cudaFuncSetAttribute(attn_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd::kSmemBytes);
  if (ce != cudaSuccess) return (int)ce;
  if (VBX_FWD_STAGGER > 0) {   // per-SM arrival counters of the stagger experiment (library-owned static device memory)
    void* slots = nullptr;
    if ((ce = cudaGetSymbolAddress(&slots, g_sm_slot)) != cudaSuccess) return (int)ce;
    if ((ce = cudaMemsetAsync(slots, 0, sizeof(unsigned) * 256, (cudaStream_t)stream)) != cudaSuccess) return (int)ce;
  }
  attn_fwd2_kernel<<<grid, fwd::kThreads, fwd::kSmemBytes, (cudaStream_t)stream>>>(mq, mk, mv, key_mask, scale * kLog2e, o, lse, (int)N,
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
  CUtensorMap mq, mk, mv, mdo, mdq;
  int rc;
  if ((rc = make_tmap_bf16_4d(&mq, q, N, H, B, VBX_QK_STRIDES(N, H), kBM)) != VBX_OK) return rc;
  if ((rc = make_tmap_bf16_4d(&mk, k, N, H, B, VBX_QK_STRIDES(N, H), kBN)) != VBX_OK) return rc;
  if ((rc = make_tmap_bf16_4d(&mv, v, N, H, B, v_ns, kDh, v_bs, kBN)) != VBX_OK) return rc;
  if ((rc = make_tmap_bf16_4d(&mdo, dout, N, H, B, H * kDh, kDh, N * H * kDh, kBM)) != VBX_OK) return rc;
  if ((rc = make_tmap(&mdq, dq, kDh, N, H, B, VBX_QK_STRIDES(N, H), 32, 4)) != VBX_OK) return rc;  // 32x32 fp32 boxes
  cudaError_t ce = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd::kSmemBytes);
  if (ce != cudaSuccess) return (int)ce;  // per call: the attribute is per device
  cudaStream_t s = (cudaStream_t)stream;
  attn_delta_kernel<<<grid_for(B * N * H, 32, 8), 256, 0, s>>>(o, dout, delta, B, N, (int)H);
  dim3 grid((unsigned)((N + kBN - 1) / kBN), (unsigned)H, (unsigned)B);
  attn_bwd_kernel<<<grid, bwd::kThreads, bwd::kSmemBytes, s>>>(mq, mk, mv, mdo, mdq, key_mask, scale, scale * kLog2e, lse, delta, dk, dv,
                                                     dv_bs, dv_ns, (int)N, (int)H);
  return VBX_LAUNCH_RC();
}

#ifdef VBX_TRACE
// ---- MMA throughput microbenchmark (trace build only): cycles for `iters` x 8 back-to-back MMAs of one configuration ----
namespace vbx {
template <int NACC, bool ATMEM>
__global__ void __launch_bounds__(128, 1) umma_bench_kernel(long long* out, int n, int a_mn, int b_mn, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * kSubTileBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 4 * (int)kSubTileBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tmem_slot;
  if (warp == 1) {  // warp-uniform issue loop, leader lane issues
    const uint32_t idesc = make_idesc(128, n, a_mn != 0, b_mn != 0);
    const uint64_t a0 = a_mn ? sdesc_mn0(smem_u32(smem)) : sdesc_k0(smem_u32(smem));
    const uint64_t b0 = b_mn ? sdesc_mn0(smem_u32(smem + 2 * kSubTileBytes)) : sdesc_k0(smem_u32(smem + 2 * kSubTileBytes));
    const uint64_t as = a_mn ? koff_mn(1) : 2, bs = b_mn ? koff_mn(1) : 2;
    const bool leader = (threadIdx.x & 31) == 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (leader) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // accumulators rotate over NACC independent TMEM tiles (compile-time addresses)
          const uint32_t d = tb + (uint32_t)(k % NACC) * 64;
          if (ATMEM) umma_bf16_ts(d, tb + 256 + (k & 3) * 8, b0 + (k & 3) * bs, idesc, 1);
          else umma_bf16(d, a0 + (k & 3) * as, b0 + (k & 3) * bs, idesc, 1);
        }
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (leader) umma_commit(&bars[0]);
    mbar_wait(&bars[0], 0);
    const long long t2 = clock64();
    if (leader) {
      out[0] = t1 - t0;  // issue time
      out[1] = t2 - t0;  // until all retired
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}
}  // namespace vbx
extern "C" int vbx_debug_umma_bench(long long* out, int n, int a_mn, int b_mn, int a_tmem, int iters, int ctas, int nacc) {
  const int smem = 4 * ptx::kSubTileBytes + 64;
#define VBX_BENCH(NA, AT)                                                                                         \
  {                                                                                                               \
    cudaFuncSetAttribute(vbx::umma_bench_kernel<NA, AT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);      \
    vbx::umma_bench_kernel<NA, AT><<<ctas, 128, smem>>>(out, n, a_mn, b_mn, iters);                               \
  }
  if (a_tmem) {
    if (nacc == 1) VBX_BENCH(1, true) else if (nacc == 2) VBX_BENCH(2, true) else VBX_BENCH(4, true)
  } else {
    if (nacc == 1) VBX_BENCH(1, false) else if (nacc == 2) VBX_BENCH(2, false) else VBX_BENCH(4, false)
  }
#undef VBX_BENCH
  return (int)cudaGetLastError();
}
// ---- same measurement with SEVERAL issuing warps (each on its own accumulators): is the ~55 clk floor of N <= 64 MMAs a
// property of the tensor pipe, or of one thread's issue path (ptxas wraps every tcgen05.mma in an ELECT / BRA.U.ANY loop)? ----
namespace vbx {
template <bool ATMEM>
__global__ void __launch_bounds__(160, 1) umma_bench_mw_kernel(long long* out, int n, int a_mn, int b_mn, int iters, int nissuers) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * kSubTileBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 4 * (int)kSubTileBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tmem_slot;
  if (warp >= 1 && warp <= nissuers) {
    const int w = warp - 1;
    const uint32_t idesc = make_idesc(128, n, a_mn != 0, b_mn != 0);
    const uint64_t a0 = a_mn ? sdesc_mn0(smem_u32(smem)) : sdesc_k0(smem_u32(smem));
    const uint64_t b0 = b_mn ? sdesc_mn0(smem_u32(smem + 2 * kSubTileBytes)) : sdesc_k0(smem_u32(smem + 2 * kSubTileBytes));
    const uint64_t as = a_mn ? koff_mn(1) : 2, bs = b_mn ? koff_mn(1) : 2;
    const bool leader = (threadIdx.x & 31) == 0;
    const uint32_t acc = tb + (uint32_t)w * 128;      // this warp's accumulator (n <= 128 columns); A-in-TMEM lives at col 448+
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (leader) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (ATMEM) umma_bf16_ts(acc, tb + 448 + (k & 3) * 8, b0 + (k & 3) * bs, idesc, 1);
          else umma_bf16(acc, a0 + (k & 3) * as, b0 + (k & 3) * bs, idesc, 1);
        }
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (leader) umma_commit(&bars[w]);
    mbar_wait(&bars[w], 0);
    const long long t2 = clock64();
    if (leader) {
      out[2 * w] = t1 - t0;
      out[2 * w + 1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}
}  // namespace vbx
extern "C" int vbx_debug_umma_bench_mw(long long* out, int n, int a_mn, int b_mn, int a_tmem, int iters, int ctas, int nissuers) {
  const int smem = 4 * ptx::kSubTileBytes + 128;
  if (nissuers < 1 || nissuers > 3 || n > 128) return VBX_E_SHAPE;   // 3 x 128 accumulator columns + 64 for the TMEM A operand
  if (a_tmem) {
    cudaFuncSetAttribute(vbx::umma_bench_mw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    vbx::umma_bench_mw_kernel<true><<<ctas, 160, smem>>>(out, n, a_mn, b_mn, iters, nissuers);
  } else {
    cudaFuncSetAttribute(vbx::umma_bench_mw_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    vbx::umma_bench_mw_kernel<false><<<ctas, 160, smem>>>(out, n, a_mn, b_mn, iters, nissuers);
  }
  return (int)cudaGetLastError();
}
extern "C" int vbx_debug_set_trace(void* dev_ptr) {
  return (int)cudaMemcpyToSymbol(vbx::g_trace, &dev_ptr, sizeof(void*));
}
#endif

extern "C" int vbx_umma_selftest(const uint16_t* a, const uint16_t* b, float* c, int variant, void* stream) {
  VBX_REQUIRE(a && b && c, VBX_E_NULL);
  VBX_REQUIRE(variant >= 0 && variant < 16, VBX_E_SHAPE);
  const int b_mn = variant & 1, a_mn = (variant >> 1) & 1, tma = (variant >> 2) & 1, a_tmem = (variant >> 3) & 1;
  VBX_REQUIRE(!(a_tmem && a_mn), VBX_E_UNSUPPORTED);  // a TMEM A operand cannot be transposed
  CUtensorMap ma, mb;
  int rc;
  if ((rc = make_tmap(&ma, a, 128, 128, 1, 1, 128, 128 * 128, 128 * 128, 128)) != VBX_OK) return rc;
  if ((rc = make_tmap(&mb, b, 128, 128, 1, 1, 128, 128 * 128, 128 * 128, 128)) != VBX_OK) return rc;
  const int smem = 4 * ptx::kSubTileBytes + 64;
  cudaStream_t s = (cudaStream_t)stream;
  if (tma) {
    cudaFuncSetAttribute(umma_selftest_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    umma_selftest_kernel<true><<<1, 128, smem, s>>>(ma, mb, a, b, c, a_mn, b_mn, a_tmem);
  } else {
    cudaFuncSetAttribute(umma_selftest_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    umma_selftest_kernel<false><<<1, 128, smem, s>>>(ma, mb, a, b, c, a_mn, b_mn, a_tmem);
  }
  return VBX_LAUNCH_RC();
}
