// Operand packing: every fp32 master weight / bias of the trunk -> its bf16 tensor-core operand copy, ALL in one launch.
//
// The reference reaches the same state through autocast: each `F.linear` under `torch.autocast(bfloat16)` casts its fp32
// weight and bias to bf16 at use (trainer.py:267), i.e. ~14 cast kernels per layer per forward (plus zero-padding copies for
// the GEGLU operands, whose inner width int(8D/3) is not 16-byte aligned).  Here a table of segments
//   { src f32 [rows x cols], row pitch src_pitch }  ->  { dst bf16, row pitch dst_pitch }
// is walked by one grid: one warp per row, 16-byte loads where the row allows it.  Destination padding is never written
// (the caller allocates the operand buffers zero-filled once).
#include "common.cuh"

namespace vbx {

struct PackSeg {          // mirrors the int64 [n_seg, 6] table built by pack.py
  const float* src;
  uint16_t* dst;
  int64_t rows, cols, src_pitch, dst_pitch;
};

// row_start: exclusive prefix sum of rows over the segments (int64 [n_seg + 1])
__global__ void __launch_bounds__(256) pack_bf16_kernel(const PackSeg* __restrict__ segs, const int64_t* __restrict__ row_start,
                                                         int n_seg) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t total = row_start[n_seg];
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < total; row += warps) {
    int lo = 0, hi = n_seg;                       // largest s with row_start[s] <= row (warp-uniform search)
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (row_start[mid] <= row) lo = mid; else hi = mid;
    }
    const PackSeg sg = segs[lo];
    const int64_t r = row - row_start[lo];
    const float* s = sg.src + r * sg.src_pitch;
    uint16_t* d = sg.dst + r * sg.dst_pitch;
    const int64_t n = sg.cols;
    if (((reinterpret_cast<uintptr_t>(s) & 15) == 0) && ((reinterpret_cast<uintptr_t>(d) & 7) == 0)) {
      const int64_t n4 = n >> 2;
#pragma unroll 4
      for (int64_t i = lane; i < n4; i += 32) {
        const uint4 u = ldg_nc_16(s + 4 * i);
        __nv_bfloat162 a = f2bf(__uint_as_float(u.x), __uint_as_float(u.y)), b = f2bf(__uint_as_float(u.z), __uint_as_float(u.w));
        *reinterpret_cast<uint2*>(d + 4 * i) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
      }
      for (int64_t i = (n4 << 2) + lane; i < n; i += 32) d[i] = __bfloat16_as_ushort(__float2bfloat16_rn(s[i]));
    } else if (((reinterpret_cast<uintptr_t>(s) & 7) == 0) && ((reinterpret_cast<uintptr_t>(d) & 3) == 0)) {
      const int64_t n2 = n >> 1;                  // rows of an odd-pitch matrix (e.g. [D, 2730] fp32): 8-byte aligned only
      for (int64_t i = lane; i < n2; i += 32) {
        const float2 u = *reinterpret_cast<const float2*>(s + 2 * i);
        __nv_bfloat162 a = f2bf(u.x, u.y);
        *reinterpret_cast<uint32_t*>(d + 2 * i) = *reinterpret_cast<uint32_t*>(&a);
      }
      if ((n & 1) && lane == 0) d[n - 1] = __bfloat16_as_ushort(__float2bfloat16_rn(s[n - 1]));
    } else {
      for (int64_t i = lane; i < n; i += 32) d[i] = __bfloat16_as_ushort(__float2bfloat16_rn(s[i]));
    }
  }
}

// dst (f32, row pitch dst_pitch) += src (bf16, row pitch src_pitch) over a [rows x cols] block: the accumulation of a bf16 weight
// gradient (fast bf16-output library GEMM) into the fp32 master gradient in ONE pass of 10 B / element -- what autograd does in
// two (cast to fp32: 6 B, then add into .grad: 12 B).
__global__ void __launch_bounds__(256) accum_bf16_2d_kernel(float* __restrict__ dst, int64_t dst_pitch, const uint16_t* __restrict__ src,
                                                             int64_t src_pitch, int64_t rows, int64_t cols) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    float* d = dst + r * dst_pitch;
    const uint16_t* s = src + r * src_pitch;
    if (((reinterpret_cast<uintptr_t>(d) & 15) == 0) && ((reinterpret_cast<uintptr_t>(s) & 7) == 0)) {
      const int64_t n4 = cols >> 2;
#pragma unroll 4
      for (int64_t i = lane; i < n4; i += 32) {
        const uint2 u = *reinterpret_cast<const uint2*>(s + 4 * i);
        const float2 a = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&u.x)), b = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
        const uint4 g = ldg_16(d + 4 * i);
        stg_16(d + 4 * i, make_uint4(__float_as_uint(__uint_as_float(g.x) + a.x), __float_as_uint(__uint_as_float(g.y) + a.y),
                                     __float_as_uint(__uint_as_float(g.z) + b.x), __float_as_uint(__uint_as_float(g.w) + b.y)));
      }
      for (int64_t i = (n4 << 2) + lane; i < cols; i += 32) d[i] += __uint_as_float((uint32_t)s[i] << 16);
    } else {
      for (int64_t i = lane; i < cols; i += 32) d[i] += __uint_as_float((uint32_t)s[i] << 16);
    }
  }
}

// table form: segment i adds a bf16 [rows x cols] block into an f32 block (same 6 x int64 layout as PackSeg, with src = bf16 and
// dst = f32): all 96 gamma/beta weight gradients of a backward in ONE launch
struct AccumSeg {
  const uint16_t* src;
  float* dst;
  int64_t rows, cols, src_pitch, dst_pitch;
};
__global__ void __launch_bounds__(256) accum_bf16_table_kernel(const AccumSeg* __restrict__ segs, const int64_t* __restrict__ row_start,
                                                                int n_seg) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t total = row_start[n_seg];
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < total; row += warps) {
    int lo = 0, hi = n_seg;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (row_start[mid] <= row) lo = mid; else hi = mid;
    }
    const AccumSeg sg = segs[lo];
    const int64_t r = row - row_start[lo];
    float* d = sg.dst + r * sg.dst_pitch;
    const uint16_t* s = sg.src + r * sg.src_pitch;
    const int64_t cols = sg.cols;
    if (((reinterpret_cast<uintptr_t>(d) & 15) == 0) && ((reinterpret_cast<uintptr_t>(s) & 7) == 0)) {
      const int64_t n4 = cols >> 2;
#pragma unroll 4
      for (int64_t i = lane; i < n4; i += 32) {
        const uint2 u = *reinterpret_cast<const uint2*>(s + 4 * i);
        const float2 a = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&u.x)), b = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
        const uint4 g = ldg_16(d + 4 * i);
        stg_16(d + 4 * i, make_uint4(__float_as_uint(__uint_as_float(g.x) + a.x), __float_as_uint(__uint_as_float(g.y) + a.y),
                                     __float_as_uint(__uint_as_float(g.z) + b.x), __float_as_uint(__uint_as_float(g.w) + b.y)));
      }
      for (int64_t i = (n4 << 2) + lane; i < cols; i += 32) d[i] += __uint_as_float((uint32_t)s[i] << 16);
    } else {
      for (int64_t i = lane; i < cols; i += 32) d[i] += __uint_as_float((uint32_t)s[i] << 16);
    }
  }
}

}  // namespace vbx

using namespace vbx;

extern "C" int vbx_accum_bf16_table(const void* segs, const int64_t* row_start, int64_t n_seg, int64_t total_rows, void* stream) {
  VBX_REQUIRE(segs && row_start, VBX_E_NULL);
  VBX_REQUIRE(n_seg > 0 && n_seg < (1 << 30) && total_rows > 0, VBX_E_SHAPE);
  static_assert(sizeof(AccumSeg) == 48, "table layout: 6 x int64 per segment");
  accum_bf16_table_kernel<<<grid_for(total_rows, 8, 8), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const AccumSeg*>(segs), row_start,
                                                                                       (int)n_seg);
  return VBX_LAUNCH_RC();
}

extern "C" int vbx_accum_bf16_2d(float* dst, int64_t dst_pitch, const uint16_t* src, int64_t src_pitch, int64_t rows, int64_t cols,
                                 void* stream) {
  VBX_REQUIRE(dst && src, VBX_E_NULL);
  VBX_REQUIRE(rows > 0 && cols > 0 && dst_pitch >= cols && src_pitch >= cols, VBX_E_SHAPE);
  accum_bf16_2d_kernel<<<grid_for(rows, 8, 8), 256, 0, (cudaStream_t)stream>>>(dst, dst_pitch, src, src_pitch, rows, cols);
  return VBX_LAUNCH_RC();
}

extern "C" int vbx_pack_bf16(const void* segs, const int64_t* row_start, int64_t n_seg, int64_t total_rows, void* stream) {
  VBX_REQUIRE(segs && row_start, VBX_E_NULL);
  VBX_REQUIRE(n_seg > 0 && n_seg < (1 << 30) && total_rows > 0, VBX_E_SHAPE);
  VBX_REQUIRE((reinterpret_cast<uintptr_t>(segs) & 7) == 0, VBX_E_ALIGN);
  static_assert(sizeof(PackSeg) == 48, "table layout: 6 x int64 per segment");
  pack_bf16_kernel<<<grid_for(total_rows, 8, 8), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const PackSeg*>(segs), row_start,
                                                                                (int)n_seg);
  return VBX_LAUNCH_RC();
}
