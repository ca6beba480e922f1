// sm_100a primitives used by the attention kernels: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 MMA / TMEM, and the
// shared-memory / instruction descriptors for bf16 operands in the 128-byte-swizzle canonical layouts.
//
// Shared-memory tile convention used everywhere in this library ("SW128 tile"): R rows x 64 bf16 (128 bytes per row),
// row r at byte r*128, and the 16-byte chunk c of row r stored at chunk position c ^ (r & 7).  The tile base is
// 1024-byte aligned.  This is exactly what a TMA load with CU_TENSOR_MAP_SWIZZLE_128B and a {64, R} box writes, and it
// is simultaneously
//   * the canonical K-major  SW128 operand layout (rows = M or N index, the 64 columns = K)      [desc_kmajor]
//   * the canonical MN-major SW128 operand layout (rows = K index,      the 64 columns = M or N) [desc_mnmajor]
// so one tile image can feed tcgen05.mma in either role by flipping the major bit of the instruction descriptor.
// Wider operands (128 columns) are two such tiles `kSubTileBytes` apart.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace vbx {
namespace ptx {

VBX_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------------------------
VBX_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
VBX_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
VBX_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
VBX_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
VBX_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
VBX_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// generic-proxy writes to smem -> visible to the async proxy (UMMA / TMA reads)
VBX_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
VBX_DEVINL void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---- TMA --------------------------------------------------------------------------------------------------------
VBX_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
VBX_DEVINL void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// warm L2 for a box that a later tma_load_4d will fetch (takes DRAM latency off the load's critical path)
VBX_DEVINL void tma_prefetch_l2_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// smem -> global element-wise ADD through the TMA unit (reduction performed at L2): replaces per-lane red.global atomics
VBX_DEVINL void tma_reduce_add_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
VBX_DEVINL void tma_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
VBX_DEVINL void tma_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
VBX_DEVINL void tma_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- TMEM -------------------------------------------------------------------------------------------------------
VBX_DEVINL void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
VBX_DEVINL void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
VBX_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
VBX_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
VBX_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
VBX_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (taddr.lane + i), columns [col, col+32)
VBX_DEVINL void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 16 consecutive 32-bit columns, registers -> TMEM (thread i writes lane taddr.lane + i).  Used to place a bf16
// A operand (two K-consecutive elements per 32-bit column, row = lane) for TS-mode MMAs.  Caller waits with tmem_st_wait().
VBX_DEVINL void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
          taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
VBX_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// issue-only forms (caller batches several loads / stores behind ONE tcgen05.wait): v / r must be indexed statically
VBX_DEVINL void tmem_ld32_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns, registers -> TMEM; caller waits with tmem_st_wait()
VBX_DEVINL void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}

// 16-column variant of tmem_ld32
VBX_DEVINL void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- tcgen05.mma ------------------------------------------------------------------------------------------------
constexpr uint32_t kSubTileBytes = 128 * 128;  // one [128 rows][64 bf16] SW128 tile

// instruction descriptor, kind::f16, bf16 x bf16 -> f32  (bit layout: cute/arch/mma_sm100_desc.hpp InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                          // c_format  = F32
         | (1u << 7)                        // a_format  = BF16
         | (1u << 10)                       // b_format  = BF16
         | ((a_mn_major ? 1u : 0u) << 15)   // a_major
         | ((b_mn_major ? 1u : 0u) << 16)   // b_major
         | ((uint32_t)(N >> 3) << 17)       // n_dim
         | ((uint32_t)(M >> 4) << 24);      // m_dim
}

// shared-memory matrix descriptor (SmemDescriptor): SWIZZLE_128B, version 1 (Blackwell)
VBX_DEVINL uint64_t make_sdesc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version
  d |= (uint64_t)2 << 61;  // layout type SWIZZLE_128B
  return d;
}
// operand whose 64-wide contiguous dimension is K: K-step k (16 elements) lives in sub-tile k/4 at byte (k%4)*32.
// 8-row core-matrix groups are 1024 B apart (SBO); LBO is unused for swizzled K-major layouts.
VBX_DEVINL uint64_t desc_kmajor(uint32_t tile_addr, int kstep) {
  return make_sdesc(tile_addr + (uint32_t)(kstep >> 2) * kSubTileBytes + (uint32_t)(kstep & 3) * 32, 16, 1024);
}
// operand whose 64-wide contiguous dimension is M (or N): rows are K; K-step k starts at row 16k (2048 B);
// 8-row groups 1024 B apart (SBO); the second 64-wide M/N half is the next sub-tile (LBO).
VBX_DEVINL uint64_t desc_mnmajor(uint32_t tile_addr, int kstep) {
  return make_sdesc(tile_addr + (uint32_t)kstep * 2048, kSubTileBytes, 1024);
}

// Hoisted forms: build the descriptor of K-step 0 once per operand tile, then ADD a compile-time constant per K-step (the
// start-address field is the low 14 bits, in 16-byte units; shared memory is < 256 KB so the add never carries out of it).
// Rebuilding the full descriptor per MMA costs a ~90-clk dependent scalar chain in the issuing thread -- more than the MMA.
VBX_DEVINL uint64_t sdesc_k0(uint32_t tile_addr) { return make_sdesc(tile_addr, 16, 1024); }
VBX_DEVINL uint64_t sdesc_mn0(uint32_t tile_addr) { return make_sdesc(tile_addr, kSubTileBytes, 1024); }
__host__ __device__ constexpr uint64_t koff_k(int k) { return (uint64_t)(((k >> 2) * kSubTileBytes + (k & 3) * 32) >> 4); }
__host__ __device__ constexpr uint64_t koff_mn(int k) { return (uint64_t)((k * 2048) >> 4); }

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
VBX_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// TS mode: A[128 x 16] bf16 read from TMEM (lane = row, 8 columns of 2 elements), B from shared memory.  Halves the
// shared-memory operand traffic of the MMA (A cannot be transposed in this mode: it is always K-major).
VBX_DEVINL void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05 ops of this thread have completed
VBX_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// byte offset of element (row, col) inside a [rows][128 cols] bf16 operand stored as two SW128 sub-tiles
VBX_DEVINL uint32_t sw128_offset(int row, int col) {
  const int sub = col >> 6, cc = (col & 63) >> 3;
  return (uint32_t)sub * kSubTileBytes + (uint32_t)row * 128 + (uint32_t)((cc ^ (row & 7)) << 4) + (uint32_t)(col & 7) * 2;
}

}  // namespace ptx

// ---- host: TMA tensor maps ----------------------------------------------------------------------------------------
// 4-D bf16 tensor (d = 64 contiguous, n, h, b) with arbitrary element strides; box = {64, box_rows, 1, 1}, SWIZZLE_128B.
int make_tmap_bf16_4d(CUtensorMap* out, const void* base, int64_t N, int64_t H, int64_t B, int64_t n_stride, int64_t h_stride,
                      int64_t b_stride, int box_rows);

}  // namespace vbx
