// bf16 GEMM on tcgen05 tensor cores with fused epilogues (sm_100a):  C[M, N] = A[M, K] W[N, K]^T (+ bias), both operands
// K-major (activations [tokens, features] and nn.Linear weights [out, in] as they lie in memory).
//
//   mode PLAIN      C = A W^T + b                       -> bf16 [M, N]                          (vp.py:320, 333, 348, 1078, 1092)
//   mode GEGLU      h = A W1^T + b1 ; g = gelu_erf(h[:, Fp:]) * h[:, :Fp]   -> h bf16 [M, 2Fp] (saved for the backward), g bf16 [M, Fp]
//   mode GEGLU_NOH  same, h never written (inference: no backward)                               (vp.py:337-346)
// The GEGLU modes replace the Linear(D, 2F) GEMM *and* the GEGLU pass over its output: the (value | gate) pair of every output
// column lands in the same accumulator tile (columns [0,128) = value rows j0.. of W1, columns [128,256) = gate rows Fp + j0..),
// so bias, rounding to bf16 (the reference's autocast output dtype), exact-erf GELU and the product happen on the accumulator
// before anything is written: the 2Fp-wide intermediate is written once (training) or not at all (inference) and never re-read.
//
// Persistent, warp-specialised CTA (one per SM, 192 threads), tile 128 x 256 x 64, 4-stage TMA ring (48 KB per stage):
//   warp 0     TMA producer: A box {64, 128} + two W boxes {64, 128} per stage, SWIZZLE_128B (= the canonical K-major UMMA layout)
//   warp 1     TMEM allocator (all 512 columns: two 256-column fp32 accumulators) + tcgen05.mma issuer (M = 128, N = 256, K = 16)
//   warps 2-5  epilogue: tcgen05.ld the accumulator (one row per thread), bias / GELU in registers, bf16 tiles staged through a
//              private 4 KB shared-memory slice per warp in the SWIZZLE_128B pattern and written with TMA stores (clipped at the
//              tensor edges by the tensor map), overlapped with the next tile's main loop through the second accumulator.
#include <cstdlib>

#include "umma.cuh"

namespace vbx {
using namespace ptx;

namespace gemm {
constexpr int kBM = 128, kBN = 256, kBK = 64, kStages = 4;
constexpr uint32_t kABytes = kBM * kBK * 2, kBBytes = kBN * kBK * 2, kStageBytes = kABytes + kBBytes;   // 16 + 32 KB
constexpr uint32_t kOffOut = kStages * kStageBytes;                 // 2 x [128 rows][64 bf16] output staging blocks
constexpr uint32_t kOutBlockBytes = kBM * 64 * 2;
constexpr uint32_t kOffBar = kOffOut + 2 * kOutBlockBytes;
enum { FULL = 0, EMPTY = kStages, TFULL = 2 * kStages, TEMPTY = 2 * kStages + 2, NUM_BARS = 2 * kStages + 4 };
constexpr uint32_t kSmemBytes = kOffBar + NUM_BARS * 8 + 16;
static_assert(kSmemBytes <= 232448, "shared memory budget (227 KB)");
constexpr int kThreads = 192;
enum Mode { PLAIN = 0, GEGLU = 1, GEGLU_NOH = 2 };
}  // namespace gemm

VBX_DEVINL void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
VBX_DEVINL void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
VBX_DEVINL void tma_wait_group_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
// single-slice variant of store_block (below): the slice's previous store must have finished reading before it is rewritten
VBX_DEVINL void store_block_single(uint8_t* slice, const uint32_t (&pk)[32], const CUtensorMap* map, int col, int row, int lane) {
  if (lane == 0) tma_wait_group_read0();
  __syncwarp();
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<uint4*>(slice + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(map, slice, col, row);
    tma_commit_group();
  }
}
// one box from global memory into the SAME shared-memory offset of every CTA of the cluster named in cta_mask; each destination
// CTA's mbarrier (same offset) receives the complete_tx
VBX_DEVINL void tma_load_2d_multicast(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
// arrive (when all prior tcgen05 ops of this thread have completed) on the mbarrier at this offset in every CTA of cta_mask
VBX_DEVINL void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
VBX_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
VBX_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// 32 packed bf16 pairs (64 columns of this thread's row) -> this warp's [32 rows][128 B] SWIZZLE_128B slice -> one TMA store.
// `slice` alternates between the two staging blocks: before it is overwritten, the store issued two blocks ago (the last one
// that read it) must have finished reading -- at most ONE younger bulk group may still be pending.
VBX_DEVINL void store_block(uint8_t* slice, const uint32_t (&pk)[32], const CUtensorMap* map, int col, int row, int lane) {
  if (lane == 0) tma_wait_group_read1();
  __syncwarp();
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<uint4*>(slice + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(map, slice, col, row);
    tma_commit_group();
  }
}

VBX_DEVINL float bf16_to_float(uint16_t v) { return __uint_as_float((uint32_t)v << 16); }
VBX_DEVINL uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = f2bf(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// bias[col .. col+32) (bf16) -> fp32; all threads of the CTA read the same addresses (L1 broadcast).  Columns at or beyond
// n_cols (clipped by the TMA store anyway) read the last valid entry: nothing is fetched out of bounds for any N % 8 == 0.
VBX_DEVINL void load_bias32(const uint16_t* bias, int col, int n_cols, float (&b)[32]) {
  if (col + 32 > n_cols) {
#pragma unroll
    for (int i = 0; i < 32; ++i) b[i] = bf16_to_float(bias[min(col + i, n_cols - 1)]);
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float t[8];
    unpack8(*reinterpret_cast<const uint4*>(bias + col + 8 * i), t);
#pragma unroll
    for (int x = 0; x < 8; ++x) b[8 * i + x] = t[x];
  }
}

// mA: A [M, K]; mW: W [N_total, K] (box 128 rows); output maps (box {64, 32}): PLAIN: mO0 = C.  GEGLU: mO0 = value half of h,
// mO1 = gate half of h, mO2 = g.  n_tiles = column tiles (PLAIN: ceil(N / 256); GEGLU: ceil(Fp / 128)).  gate_row0: row of W
// where the second 128-row box of a tile starts, relative to the first: PLAIN 128, GEGLU Fp.
// CLUSTER = 2: two CTAs of a cluster work on vertically adjacent tiles (same W rows).  Each loads ITS activation tile and only
// HALF of the W tile, multicast into both CTAs' shared memory: W traffic from L2 per tile halves (48 -> 32 KB per k-block).
// A stage may then only be refilled once BOTH CTAs' MMAs have read it: EMPTY counts two commits, each multicast to both CTAs.
template <int MODE, int CLUSTER>
__global__ void __launch_bounds__(gemm::kThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mW, const __grid_constant__ CUtensorMap mO0,
                 const __grid_constant__ CUtensorMap mO1, const __grid_constant__ CUtensorMap mO2, const uint16_t* __restrict__ bias,
                 int M, int n_cols, int K, int m_tiles, int n_tiles, int second_box_row) {
  using namespace gemm;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + NUM_BARS * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (K + kBK - 1) / kBK;
  constexpr int kColsPerTile = (MODE == PLAIN) ? 256 : 128;   // output columns a tile advances by
  // work items: CLUSTER == 1: tile t -> (m_blk, n_blk) = (t / n_tiles, t % n_tiles), CTA b takes t = b, b + grid, ...
  //             CLUSTER == 2: pair p -> m_blk = 2 (p / n_tiles) + rank, n_blk = p % n_tiles, cluster c takes p = c, c + grid/2, ...
  const int rank = CLUSTER == 2 ? (int)cluster_ctarank() : 0;
  const int nwork = CLUSTER == 2 ? ((m_tiles + 1) / 2) * n_tiles : m_tiles * n_tiles;
  const int w0 = CLUSTER == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int wstep = CLUSTER == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
#define VBX_TILE_COORDS(wi)                                                                   \
  const int m0 = (CLUSTER == 2 ? 2 * ((wi) / n_tiles) + rank : (wi) / n_tiles) * kBM,        \
            j0 = ((wi) % n_tiles) * kColsPerTile

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars[FULL + s], 1);
      mbar_init(&bars[EMPTY + s], CLUSTER);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars[TFULL + a], 1);
      mbar_init(&bars[TEMPTY + a], 128);
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mA);
    tma_prefetch_desc(&mW);
    tma_prefetch_desc(&mO0);
    if (MODE == GEGLU) tma_prefetch_desc(&mO1);
    if (MODE != PLAIN) tma_prefetch_desc(&mO2);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (CLUSTER == 2) cluster_sync_all();   // the peer's barriers are initialised before anything is multicast to / arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer ------------------------------------------------
    if (lane == 0) {
      int it = 0;
      for (int wi = w0; wi < nwork; wi += wstep) {
        VBX_TILE_COORDS(wi);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(&bars[EMPTY + s], ((it / kStages) & 1) ^ 1);
          uint8_t* st = smem + s * kStageBytes;
          mbar_arrive_expect_tx(&bars[FULL + s], kStageBytes);
          tma_load_2d(st, &mA, &bars[FULL + s], kb * kBK, m0);
          if (CLUSTER == 2) {   // this CTA fetches rows [128 rank, +128) of the 256-row W tile for BOTH CTAs
            tma_load_2d_multicast(st + kABytes + rank * (kBBytes / 2), &mW, &bars[FULL + s], kb * kBK,
                                  (rank ? second_box_row : 0) + j0, (uint16_t)3);
          } else {
            tma_load_2d(st + kABytes, &mW, &bars[FULL + s], kb * kBK, j0);
            tma_load_2d(st + kABytes + kBBytes / 2, &mW, &bars[FULL + s], kb * kBK, second_box_row + j0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer --------------------------------------------------
    constexpr uint32_t idesc = make_idesc(kBM, kBN, false, false);
    const uint64_t dA0 = sdesc_k0(smem_u32(smem)), dB0 = sdesc_k0(smem_u32(smem + kABytes));
    const bool leader = lane == 0;
    int it = 0, tcount = 0;
    for (int wi = w0; wi < nwork; wi += wstep, ++tcount) {
      const int acc = tcount & 1;
      mbar_wait(&bars[TEMPTY + acc], ((tcount >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * kBN;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % kStages;
        mbar_wait(&bars[FULL + s], (it / kStages) & 1);
        tc_fence_after();
        if (leader) {
          const uint64_t so = (uint64_t)s * (kStageBytes >> 4);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) umma_bf16(d_tmem, dA0 + so + koff_k(k), dB0 + so + koff_k(k), idesc, (kb | k) != 0);
          if (CLUSTER == 2) umma_commit_multicast(&bars[EMPTY + s], (uint16_t)3);
          else umma_commit(&bars[EMPTY + s]);
          if (kb == nkb - 1) umma_commit(&bars[TFULL + acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------ epilogue ----------------------------------------------------
    const int q = warp & 3;                                   // TMEM lane quarter this warp may access
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* slice0 = smem + kOffOut + q * 4096;
    uint8_t* slice1 = slice0 + kOutBlockBytes;
    int tcount = 0, blk = 0;
    for (int wi = w0; wi < nwork; wi += wstep, ++tcount) {
      const int acc = tcount & 1;
      VBX_TILE_COORDS(wi);
      const int row = m0 + q * 32;
      mbar_wait(&bars[TFULL + acc], (tcount >> 1) & 1);
      tc_fence_after();
      const uint32_t t_acc = t_lane + (uint32_t)acc * kBN;
      if (MODE == PLAIN) {
#pragma unroll 1
        for (int jj = 0; jj < 4; ++jj) {                       // 64 output columns per block
          uint32_t pk[32];
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            float v[32], b[32];
            tmem_ld32(t_acc + jj * 64 + hf * 32, v);
            if (jj == 3 && hf == 1) {                          // last read of this accumulator: hand it back to the MMA warp
              tc_fence_before();
              mbar_arrive(&bars[TEMPTY + acc]);
            }
            if (bias != nullptr) {
              load_bias32(bias, j0 + jj * 64 + hf * 32, n_cols, b);
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] += b[i];
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[hf * 16 + i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
          }
          store_block((blk & 1) ? slice1 : slice0, pk, &mO0, j0 + jj * 64, row, lane);
          ++blk;
        }
      } else {
#pragma unroll 1
        for (int jj = 0; jj < 2; ++jj) {                       // value columns [64jj, +64) with gate columns [128 + 64jj, +64)
          uint32_t hv[32], hg[32], gg[32];
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            float v[32], g[32], b[32];
            tmem_ld32(t_acc + jj * 64 + hf * 32, v);
            tmem_ld32(t_acc + 128 + jj * 64 + hf * 32, g);
            if (jj == 1 && hf == 1) {
              tc_fence_before();
              mbar_arrive(&bars[TEMPTY + acc]);
            }
            const int col = j0 + jj * 64 + hf * 32;                       // n_cols = Fp here (a multiple of 32)
            load_bias32(bias, col, n_cols, b);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += b[i];
            load_bias32(bias + n_cols, col, n_cols, b);
#pragma unroll
            for (int i = 0; i < 32; ++i) g[i] += b[i];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              // h is a bf16 tensor in the reference (autocast output of the Linear): GEGLU sees the ROUNDED value and gate
              const __nv_bfloat162 v2 = f2bf(v[2 * i], v[2 * i + 1]), g2 = f2bf(g[2 * i], g[2 * i + 1]);
              const float2 vf = bf2f(v2), gf = bf2f(g2);
              hv[hf * 16 + i] = *reinterpret_cast<const uint32_t*>(&v2);
              hg[hf * 16 + i] = *reinterpret_cast<const uint32_t*>(&g2);
              gg[hf * 16 + i] = pack_bf16x2(gelu_f(gf.x) * vf.x, gelu_f(gf.y) * vf.y);
            }
          }
          const int col = j0 + jj * 64;
          if (MODE == GEGLU) {
            store_block((blk & 1) ? slice1 : slice0, hv, &mO0, col, row, lane);
            ++blk;
            store_block((blk & 1) ? slice1 : slice0, hg, &mO1, col, row, lane);
            ++blk;
          }
          store_block((blk & 1) ? slice1 : slice0, gg, &mO2, col, row, lane);
          ++blk;
        }
      }
    }
    if (lane == 0) tma_wait_group0();                          // shared memory must outlive the last TMA store's read
  }
  tc_fence_before();
  __syncthreads();
  if (CLUSTER == 2) cluster_sync_all();   // the peer may still multicast into this CTA's shared memory / arrive on its barriers
  if (warp == 1) tmem_dealloc(tmem_base, 512);
#undef VBX_TILE_COORDS
}

// =====================================================================================================================
// FF2 data-gradient GEMM with the GEGLU backward as its epilogue (vp.py:337-348, backward):
//   dg = dy W2            [M, Fp]   (A = dy [M, D] K-major;  B = W2^T operand [Fp, D] K-major, zero rows beyond F)
//   dh[:, :Fp]  = dg * gelu_erf(gate)              dh[:, Fp:] = dg * value * gelu_erf'(gate)          (h = [value | gate], bf16)
//   db1[c]     += sum over rows of dh[:, c]        (fp32; the bias gradient of the first Linear)
// Replaces the library dgrad GEMM *and* the stand-alone geglu_bwd pass (1.83 GB of HBM traffic per layer at cfg3: dg is never
// written, h is read once, dh is written once).  Same main loop as gemm_bf16_kernel, 3 stages, EIGHT epilogue warps (the GEGLU
// backward costs ~2.5x the forward's math per element): warps 2-5 take accumulator columns [0,128), warps 6-9 columns [128,256);
// both sets cover the four TMEM lane quarters.  Per 64-column chunk a warp
//   1. TMA-loads its [32 rows x 64] value and gate blocks of h into its two 4 KB slices (issued before it waits for the
//      accumulator, so the first chunk's latency hides behind the main loop).  Measured (tools/trip9.sh): per-thread row loads --
//      32 different 128-byte lines per warp instruction -- made the first version 1,544 us, 2.4x SLOWER than the pair it
//      replaces; with the loads stubbed out the same kernel ran in 354 us;
//   2. computes dh for its row, overwrites the slices in place and TMA-stores them;
//   3. reads the column sums back from the staged bf16 block (32 conflict-free LDS per lane; a register butterfly over 128 live
//      floats spilled) and adds them into a shared-memory vector -- the value half lives in cluster rank 0's shared memory, the
//      gate half in rank 1's (red.shared::cluster), flushed to global once per CTA at the end.  Global atomics per warp per tile
//      (12 M per launch on 5,504 addresses) cost 240 us.  Without a cluster (single tile row) the sums go to global directly.
// =====================================================================================================================
// VBX_BWD_ABL: development-only ablation mask of the backward epilogue (1: no bias-gradient sums, 4: no GELU math, 8: no dh
// stores) used by tools/gemm_bench.py to attribute its time; the product builds with 0
#ifndef VBX_BWD_ABL
#define VBX_BWD_ABL 0
#endif
#ifndef VBX_BWD_F32X2
#define VBX_BWD_F32X2 1
#endif
namespace gemmb {
constexpr int kBM = 128, kBN = 256, kBK = 64, kStages = 3;
constexpr uint32_t kABytes = kBM * kBK * 2, kBBytes = kBN * kBK * 2, kStageBytes = kABytes + kBBytes;
constexpr uint32_t kOffOut = kStages * kStageBytes;                 // 8 warps x 2 slices x 4 KB: [32 rows][64 bf16] value / gate blocks
constexpr uint32_t kOffBar = kOffOut + 8 * 8192;
enum { FULL = 0, EMPTY = kStages, TFULL = 2 * kStages, TEMPTY = 2 * kStages + 2, HFULL = 2 * kStages + 4, NUM_BARS = 2 * kStages + 12 };
constexpr uint32_t kOffDb = kOffBar + 256;                          // f32 [Fp]: column sums of dh (this rank's half of the bias gradient)
constexpr uint32_t kMaxSmem = 232448;
constexpr int kThreads = 320;
}  // namespace gemmb

VBX_DEVINL uint32_t mapa_shared(uint32_t addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
  return r;
}
VBX_DEVINL void red_add_f32_cluster(uint32_t cluster_addr, float v) {
  asm volatile("red.relaxed.cluster.shared::cluster.add.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}

template <int CLUSTER>
__global__ void __launch_bounds__(gemmb::kThreads, 1)
gemm_geglu_bwd_kernel(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mW, const __grid_constant__ CUtensorMap mO0,
                      const __grid_constant__ CUtensorMap mO1, const __grid_constant__ CUtensorMap mH, float* __restrict__ db, int M, int Fp,
                      int K, int m_tiles, int n_tiles) {
  using namespace gemmb;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + NUM_BARS * 8);
  float* s_db = reinterpret_cast<float*>(smem + kOffDb);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (K + kBK - 1) / kBK;
  const int rank = CLUSTER == 2 ? (int)cluster_ctarank() : 0;
  const int nwork = CLUSTER == 2 ? ((m_tiles + 1) / 2) * n_tiles : m_tiles * n_tiles;
  const int w0 = CLUSTER == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int wstep = CLUSTER == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
#define VBX_TILE_COORDS(wi)                                                                   \
  const int m0 = (CLUSTER == 2 ? 2 * ((wi) / n_tiles) + rank : (wi) / n_tiles) * kBM,        \
            j0 = ((wi) % n_tiles) * kBN

  if (CLUSTER == 2)
    for (int i = threadIdx.x; i < Fp; i += blockDim.x) s_db[i] = 0.f;
  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars[FULL + s], 1);
      mbar_init(&bars[EMPTY + s], CLUSTER);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars[TFULL + a], 1);
      mbar_init(&bars[TEMPTY + a], 256);
    }
    for (int w = 0; w < 8; ++w) mbar_init(&bars[HFULL + w], 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mA);
    tma_prefetch_desc(&mW);
    tma_prefetch_desc(&mO0);
    tma_prefetch_desc(&mO1);
    tma_prefetch_desc(&mH);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (CLUSTER == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int wi = w0; wi < nwork; wi += wstep) {
        VBX_TILE_COORDS(wi);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(&bars[EMPTY + s], ((it / kStages) & 1) ^ 1);
          uint8_t* st = smem + s * kStageBytes;
          mbar_arrive_expect_tx(&bars[FULL + s], kStageBytes);
          tma_load_2d(st, &mA, &bars[FULL + s], kb * kBK, m0);
          if (CLUSTER == 2) {
            tma_load_2d_multicast(st + kABytes + rank * (kBBytes / 2), &mW, &bars[FULL + s], kb * kBK, rank * 128 + j0, (uint16_t)3);
          } else {
            tma_load_2d(st + kABytes, &mW, &bars[FULL + s], kb * kBK, j0);
            tma_load_2d(st + kABytes + kBBytes / 2, &mW, &bars[FULL + s], kb * kBK, 128 + j0);
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(kBM, kBN, false, false);
    const uint64_t dA0 = sdesc_k0(smem_u32(smem)), dB0 = sdesc_k0(smem_u32(smem + kABytes));
    const bool leader = lane == 0;
    int it = 0, tcount = 0;
    for (int wi = w0; wi < nwork; wi += wstep, ++tcount) {
      const int acc = tcount & 1;
      mbar_wait(&bars[TEMPTY + acc], ((tcount >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * kBN;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % kStages;
        mbar_wait(&bars[FULL + s], (it / kStages) & 1);
        tc_fence_after();
        if (leader) {
          const uint64_t so = (uint64_t)s * (kStageBytes >> 4);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) umma_bf16(d_tmem, dA0 + so + koff_k(k), dB0 + so + koff_k(k), idesc, (kb | k) != 0);
          if (CLUSTER == 2) umma_commit_multicast(&bars[EMPTY + s], (uint16_t)3);
          else umma_commit(&bars[EMPTY + s]);
          if (kb == nkb - 1) umma_commit(&bars[TFULL + acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------ epilogue: 8 warps ---------------------------------------------
    const int q = warp & 3;                                   // TMEM lane quarter
    const int grp = (warp - 2) >> 2;                          // 0: accumulator columns [0,128), 1: [128,256)
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* slice_v = smem + kOffOut + (warp - 2) * 8192;    // value block in, d(value) block out
    uint8_t* slice_g = slice_v + 4096;                        // gate block in, d(gate) block out
    uint64_t* hbar = &bars[HFULL + (warp - 2)];
    // where this warp's column sums go: shared-memory vectors of the cluster pair (value half: rank 0, gate half: rank 1)
    const uint32_t sdb_v = CLUSTER == 2 ? mapa_shared(smem_u32(s_db), 0) : 0u;
    const uint32_t sdb_g = CLUSTER == 2 ? mapa_shared(smem_u32(s_db), 1) : 0u;
    int tcount = 0, hcount = 0;
    for (int wi = w0; wi < nwork; wi += wstep, ++tcount) {
      const int acc = tcount & 1;
      VBX_TILE_COORDS(wi);
      const int row0 = m0 + q * 32;                           // rows >= M: TMA zero-fills h and dy, so their dh rows are zeros
      const uint32_t t_acc = t_lane + (uint32_t)acc * kBN + grp * 128;
#pragma unroll 1
      for (int jj = 0; jj < 2; ++jj) {
        const int col = j0 + grp * 128 + jj * 64;             // first dg / value column of this 64-column chunk
        const bool col_ok = col < Fp;                         // Fp % 64 == 0: a chunk is entirely inside or entirely outside
        if (col_ok) {
          if (lane == 0) {
            tma_wait_group_read0();                           // the previous chunk's stores have finished reading the slices
            mbar_arrive_expect_tx(hbar, 8192);
            tma_load_2d(slice_v, &mH, hbar, col, row0);
            tma_load_2d(slice_g, &mH, hbar, Fp + col, row0);
          }
        }
        if (jj == 0) {
          mbar_wait(&bars[TFULL + acc], (tcount >> 1) & 1);
          tc_fence_after();
        }
        if (col_ok) {
          mbar_wait(hbar, hcount & 1);
          ++hcount;
          uint32_t pv[32], pg[32];
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            float d[32];
            tmem_ld32(t_acc + jj * 64 + hf * 32, d);
            if (jj == 1 && hf == 1) {                          // last read of this accumulator half by this thread
              tc_fence_before();
              mbar_arrive(&bars[TEMPTY + acc]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c = hf * 4 + i;                        // 16-byte chunk of this lane's 128-byte row, at position c ^ (lane & 7)
              const uint4 hv = *reinterpret_cast<const uint4*>(slice_v + lane * 128 + ((c ^ (lane & 7)) << 4));
              const uint4 hg = *reinterpret_cast<const uint4*>(slice_g + lane * 128 + ((c ^ (lane & 7)) << 4));
              float v8[8], g8[8], dv8[8], dg8[8];
              unpack8(hv, v8);
              unpack8(hg, g8);
#if VBX_BWD_F32X2
              // two elements per instruction on the packed-fp32 pipe (FFMA2 / FMUL2): the epilogue is instruction-issue bound
              // (~26 scalar instructions per element x 32 K elements per tile against a 13.7 k-clock main loop); same
              // formula and operation order as normal_cdf() in common.cuh, so the results are bit-identical
#pragma unroll
              for (int x = 0; x < 8; x += 2) {
                const float2 g2 = make_float2(g8[x], g8[x + 1]), v2 = make_float2(v8[x], v8[x + 1]);
                // dg is a bf16 tensor in the reference's autocast backward: round it before it is used
                const float2 dd2 = make_float2(__bfloat162float(__float2bfloat16_rn(d[i * 8 + x])),
                                               __bfloat162float(__float2bfloat16_rn(d[i * 8 + x + 1])));
                const float2 z2 = __fmul2_rn(make_float2(fabsf(g2.x), fabsf(g2.y)), make_float2(0.70710678118654752f, 0.70710678118654752f));
                const float2 den = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), z2, make_float2(1.f, 1.f));
                const float2 t2 = make_float2(rcp_approx(den.x), rcp_approx(den.y));
                const float2 a2 = __fmul2_rn(__fmul2_rn(make_float2(-1.4426950408889634f, -1.4426950408889634f), z2), z2);
                const float2 e2 = make_float2(ex2_approx(a2.x), ex2_approx(a2.y));
                float2 p2 = __ffma2_rn(make_float2(1.061405429f, 1.061405429f), t2, make_float2(-1.453152027f, -1.453152027f));
                p2 = __ffma2_rn(p2, t2, make_float2(1.421413741f, 1.421413741f));
                p2 = __ffma2_rn(p2, t2, make_float2(-0.284496736f, -0.284496736f));
                p2 = __ffma2_rn(p2, t2, make_float2(0.254829592f, 0.254829592f));
                const float2 he = __fmul2_rn(__fmul2_rn(__fmul2_rn(make_float2(0.5f, 0.5f), p2), t2), e2);     // 0.5 erfc(z)
                const float2 cdf = make_float2(g2.x >= 0.f ? 1.0f - he.x : he.x, g2.y >= 0.f ? 1.0f - he.y : he.y);
                const float2 dv2 = __fmul2_rn(dd2, __fmul2_rn(g2, cdf));
                const float2 in2 = __ffma2_rn(__fmul2_rn(g2, make_float2(0.3989422804014327f, 0.3989422804014327f)), e2, cdf);
                const float2 dg2 = __fmul2_rn(__fmul2_rn(dd2, v2), in2);
                dv8[x] = dv2.x;
                dv8[x + 1] = dv2.y;
                dg8[x] = dg2.x;
                dg8[x + 1] = dg2.y;
              }
#else
#pragma unroll
              for (int x = 0; x < 8; ++x) {
                // dg is a bf16 tensor in the reference's autocast backward: round it before it is used
                const float dd = __bfloat162float(__float2bfloat16_rn(d[i * 8 + x]));
#if VBX_BWD_ABL & 4
                dv8[x] = dd * g8[x];
                dg8[x] = dd * v8[x];
#else
                float e;
                const float cdf = normal_cdf(g8[x], e);
                dv8[x] = dd * (g8[x] * cdf);
                dg8[x] = dd * v8[x] * fmaf(g8[x] * 0.3989422804014327f, e, cdf);
#endif
              }
#endif
#pragma unroll
              for (int x = 0; x < 4; ++x) {
                pv[c * 4 + x] = pack_bf16x2(dv8[2 * x], dv8[2 * x + 1]);
                pg[c * 4 + x] = pack_bf16x2(dg8[2 * x], dg8[2 * x + 1]);
              }
            }
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint8_t* slice = half ? slice_g : slice_v;
            // each lane overwrites the row it alone has read
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint32_t* pk = half ? pg : pv;
              *reinterpret_cast<uint4*>(slice + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            }
            fence_proxy_async();
            __syncwarp();
#if !(VBX_BWD_ABL & 8)
            if (lane == 0) {
              tma_store_2d(half ? &mO1 : &mO0, slice, col, row0);
              tma_commit_group();
            }
#endif
#if !(VBX_BWD_ABL & 1)
            // column sums over this warp's 32 rows: lane l owns columns 2l, 2l+1; row r's 16-byte chunk (l / 4) sits at chunk
            // position (l / 4) ^ (r & 7) -> the 32 lanes hit 32 distinct banks
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              const uint32_t w = *reinterpret_cast<const uint32_t*>(slice + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4);
              s0 += __uint_as_float(w << 16);
              s1 += __uint_as_float(w & 0xffff0000u);
            }
            if (CLUSTER == 2) {
              const uint32_t dst = (half ? sdb_g : sdb_v) + (uint32_t)(col + 2 * lane) * 4u;
              red_add_f32_cluster(dst, s0);
              red_add_f32_cluster(dst + 4u, s1);
            } else {
              atomicAdd(db + half * Fp + col + 2 * lane, s0);
              atomicAdd(db + half * Fp + col + 2 * lane + 1, s1);
            }
#endif
          }
          __syncwarp();                                        // every lane has read its sums before lane 0 reloads the slices
        } else if (jj == 1) {                                   // (warp-uniform) nothing to do, but the accumulator must be released
          tc_fence_before();
          mbar_arrive(&bars[TEMPTY + acc]);
        }
      }
    }
    if (lane == 0) tma_wait_group0();
  }
  tc_fence_before();
  __syncthreads();
  if (CLUSTER == 2) {
    cluster_sync_all();                                        // every red of both CTAs has landed
    for (int i = threadIdx.x; i < Fp; i += blockDim.x) atomicAdd(db + rank * Fp + i, s_db[i]);
  }
  if (warp == 1) tmem_dealloc(tmem_base, 512);
#undef VBX_TILE_COORDS
}

// 2-D bf16 tensor map: `inner` contiguous elements per row, `rows` rows `row_pitch` elements apart; box {64, box_rows}, SWIZZLE_128B
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, int64_t inner, int64_t rows, int64_t row_pitch, int box_rows);

}  // namespace vbx

using namespace vbx;

static int launch_gemm(int mode, const uint16_t* a, const uint16_t* w, const uint16_t* bias, uint16_t* o0, int64_t o0_pitch, uint16_t* o1,
                       uint16_t* o2, int64_t M, int64_t n_cols, int64_t K, int64_t w_rows, void* stream) {
  using namespace gemm;
  CUtensorMap mA, mW, mO0, mO1, mO2;
  int rc;
  if ((rc = make_tmap_bf16_2d(&mA, a, K, M, K, kBM)) != VBX_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&mW, w, K, w_rows, K, 128)) != VBX_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&mO0, o0, n_cols, M, o0_pitch, 32)) != VBX_OK) return rc;
  mO1 = mO0;
  mO2 = mO0;
  if (mode == GEGLU && (rc = make_tmap_bf16_2d(&mO1, o1, n_cols, M, o0_pitch, 32)) != VBX_OK) return rc;
  if (mode != PLAIN && (rc = make_tmap_bf16_2d(&mO2, o2, n_cols, M, n_cols, 32)) != VBX_OK) return rc;
  const int m_tiles = (int)((M + kBM - 1) / kBM);
  const int n_tiles = (int)(mode == PLAIN ? (n_cols + 255) / 256 : (n_cols + 127) / 128);
  // VBX_GEMM_CLUSTER=1 turns the 2-CTA W multicast off (A/B runs); a single-tile-row problem has nothing to share
  static const bool no_cluster = getenv("VBX_GEMM_CLUSTER") != nullptr && getenv("VBX_GEMM_CLUSTER")[0] == '1';
  const bool cluster2 = !no_cluster && m_tiles >= 2;
  int grid;
  if (cluster2) {
    const int pairs = ((m_tiles + 1) / 2) * n_tiles;
    grid = 2 * (pairs < kNumSM / 2 ? pairs : kNumSM / 2);
  } else {
    grid = m_tiles * n_tiles < kNumSM ? m_tiles * n_tiles : kNumSM;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster2 ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t ce;
  const int iM = (int)M, iN = (int)n_cols, iK = (int)K;
#define VBX_GEMM_LAUNCH(MODE_, SECOND)                                                                                              \
  {                                                                                                                                  \
    const int second = (int)(SECOND);                                                                                                \
    auto kern = cluster2 ? gemm_bf16_kernel<MODE_, 2> : gemm_bf16_kernel<MODE_, 1>;                                                  \
    ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);                                   \
    if (ce != cudaSuccess) return (int)ce;                                                                                           \
    ce = cudaLaunchKernelEx(&cfg, kern, mA, mW, mO0, mO1, mO2, bias, iM, iN, iK, m_tiles, n_tiles, second);                          \
    if (ce != cudaSuccess) return (int)ce;                                                                                           \
  }
  if (mode == PLAIN) {
    VBX_GEMM_LAUNCH(PLAIN, 128)
  } else if (mode == GEGLU) {
    VBX_GEMM_LAUNCH(GEGLU, n_cols)
  } else {
    VBX_GEMM_LAUNCH(GEGLU_NOH, n_cols)
  }
#undef VBX_GEMM_LAUNCH
  return VBX_LAUNCH_RC();
}

extern "C" int vbx_gemm_bf16(const uint16_t* a, const uint16_t* w, const uint16_t* bias, uint16_t* c, int64_t M, int64_t N, int64_t K,
                             void* stream) {
  VBX_REQUIRE(a && w && c, VBX_E_NULL);
  VBX_REQUIRE(M > 0 && N >= 32 && K > 0 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31) && N % 8 == 0 && K % 8 == 0, VBX_E_SHAPE);
  VBX_REQUIRE(VBX_ALIGNED16(a) && VBX_ALIGNED16(w) && VBX_ALIGNED16(c) && (!bias || VBX_ALIGNED16(bias)), VBX_E_ALIGN);
  return launch_gemm(gemm::PLAIN, a, w, bias, c, N, nullptr, nullptr, M, N, K, N, stream);
}

extern "C" int vbx_ff1_geglu(const uint16_t* x, const uint16_t* w1, const uint16_t* b1, uint16_t* h, uint16_t* g, int64_t M, int64_t Fp,
                             int64_t K, void* stream) {
  VBX_REQUIRE(x && w1 && b1 && g, VBX_E_NULL);
  VBX_REQUIRE(M > 0 && Fp >= 32 && K > 0 && M < (1ll << 31) && Fp < (1ll << 30) && K < (1ll << 31) && Fp % 32 == 0 && K % 8 == 0,
              VBX_E_SHAPE);
  VBX_REQUIRE(VBX_ALIGNED16(x) && VBX_ALIGNED16(w1) && VBX_ALIGNED16(b1) && VBX_ALIGNED16(g) && (!h || VBX_ALIGNED16(h)), VBX_E_ALIGN);
  if (h != nullptr) return launch_gemm(gemm::GEGLU, x, w1, b1, h, 2 * Fp, h + Fp, g, M, Fp, K, 2 * Fp, stream);
  return launch_gemm(gemm::GEGLU_NOH, x, w1, b1, g, Fp, nullptr, g, M, Fp, K, 2 * Fp, stream);
}

extern "C" int vbx_ff2_dgrad_geglu_bwd(const uint16_t* dy, const uint16_t* w2t, const uint16_t* h, uint16_t* dh, float* db1, int64_t M,
                                       int64_t Fp, int64_t K, void* stream) {
  using namespace gemmb;
  VBX_REQUIRE(dy && w2t && h && dh && db1, VBX_E_NULL);
  VBX_REQUIRE(M > 0 && Fp >= 64 && K > 0 && M < (1ll << 31) && Fp < (1ll << 30) && K < (1ll << 31) && Fp % 64 == 0 && K % 8 == 0,
              VBX_E_SHAPE);
  VBX_REQUIRE(VBX_ALIGNED16(dy) && VBX_ALIGNED16(w2t) && VBX_ALIGNED16(h) && VBX_ALIGNED16(dh), VBX_E_ALIGN);
  CUtensorMap mA, mW, mO0, mO1, mH;
  int rc;
  if ((rc = make_tmap_bf16_2d(&mA, dy, K, M, K, kBM)) != VBX_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&mW, w2t, K, Fp, K, 128)) != VBX_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&mH, h, 2 * Fp, M, 2 * Fp, 32)) != VBX_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&mO0, dh, Fp, M, 2 * Fp, 32)) != VBX_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&mO1, dh + Fp, Fp, M, 2 * Fp, 32)) != VBX_OK) return rc;
  const int m_tiles = (int)((M + kBM - 1) / kBM), n_tiles = (int)((Fp + kBN - 1) / kBN);
  static const bool no_cluster = getenv("VBX_GEMM_CLUSTER") != nullptr && getenv("VBX_GEMM_CLUSTER")[0] == '1';
  const bool cluster2 = !no_cluster && m_tiles >= 2;
  int grid;
  if (cluster2) {
    const int pairs = ((m_tiles + 1) / 2) * n_tiles;
    grid = 2 * (pairs < kNumSM / 2 ? pairs : kNumSM / 2);
  } else {
    grid = m_tiles * n_tiles < kNumSM ? m_tiles * n_tiles : kNumSM;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  const uint32_t smem_bytes = kOffDb + (uint32_t)(cluster2 ? Fp * 4 : 0);
  if (smem_bytes > kMaxSmem) return VBX_E_UNSUPPORTED;   // the per-CTA half of the bias-gradient vector must fit (Fp <= 4800)
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster2 ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  auto kern = cluster2 ? gemm_geglu_bwd_kernel<2> : gemm_geglu_bwd_kernel<1>;
  cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (ce != cudaSuccess) return (int)ce;
  const int iM = (int)M, iF = (int)Fp, iK = (int)K;
  ce = cudaLaunchKernelEx(&cfg, kern, mA, mW, mO0, mO1, mH, db1, iM, iF, iK, m_tiles, n_tiles);
  if (ce != cudaSuccess) return (int)ce;
  return VBX_LAUNCH_RC();
}
