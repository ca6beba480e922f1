// Depthwise conv positional embedding (ConvPositionEmbed, vp.py:203-233) as a channel-last shared-memory stencil,
// fused with: key mask before and after (vp.py:222-224, 230-231), exact GELU, the caller's residual add
// (vp.py:826, 1080), the bf16 -> fp32 promotion of the residual stream, and the register-token pack (vp.py:422-425).
//
// Tile: 64 channels x 128 tokens per CTA (4 warps).  The (128 + 30) x 64 bf16 input tile is staged once in shared
// memory with 16-byte coalesced loads; each thread then owns 2 channels x 32 consecutive tokens and slides a register
// window over its column (one 4-byte LDS per 2x31 FMAs), so the stencil is FMA-bound inside the SM and reads every
// input row once per CTA.  The reference moves the tensor 8 times (2 transposes, conv, gelu, 2 masks, add, pack).
#include "common.cuh"

namespace vbx {

constexpr int kCT = 64;    // channels per CTA
constexpr int kNT = 128;   // tokens per CTA
constexpr int kKMax = 31;  // taps (smaller odd kernels are centred and zero padded)
constexpr int kHalo = kKMax - 1;
constexpr int kRowsIn = kNT + kHalo;  // 158
constexpr int kGroup = 8;             // outputs per register block

// stage rows [n0-15, n0-15+158) x 64 channels of `src` (bf16 [B,N,C]) into smem (zero outside [0,N)).
VBX_DEVINL void stage_tile(const uint16_t* __restrict__ src, uint32_t* tile /*[158][32] bf16x2*/, int64_t b, int64_t n0, int c0,
                           int64_t N, int C) {
  for (int i = threadIdx.x; i < kRowsIn * 8; i += blockDim.x) {
    const int r = i >> 3, q = i & 7;
    const int64_t n = n0 - kHalo / 2 + r;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (n >= 0 && n < N) v = ldg_nc_16(src + (b * N + n) * C + c0 + q * 8);
    *reinterpret_cast<uint4*>(tile + r * 32 + q * 4) = v;
  }
}

__global__ void __launch_bounds__(128) convpos_fwd_kernel(const uint16_t* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, const uint8_t* __restrict__ mask,
                                                           const float* __restrict__ reg, float* __restrict__ y,
                                                           uint16_t* __restrict__ pre, int64_t N, int C, int K, int64_t R) {
  __shared__ __align__(16) uint32_t tile[kRowsIn * 32];
  __shared__ float mrow[kRowsIn];
  const int c0 = blockIdx.x * kCT;
  const int64_t n0 = (int64_t)blockIdx.y * kNT, b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  stage_tile(x, tile, b, n0, c0, N, C);
  for (int r = threadIdx.x; r < kRowsIn; r += blockDim.x) {
    const int64_t n = n0 - kHalo / 2 + r;
    mrow[r] = (n >= 0 && n < N && (mask == nullptr || mask[b * N + n])) ? 1.f : 0.f;
  }
  // taps of this thread's two channels, centred in a 31-tap window
  float w0[kKMax], w1[kKMax];
  const int pad = (kKMax - K) / 2;
#pragma unroll
  for (int j = 0; j < kKMax; ++j) {
    const int jj = j - pad;
    const bool in = jj >= 0 && jj < K;
    w0[j] = in ? w[(c0 + 2 * lane) * K + jj] : 0.f;
    w1[j] = in ? w[(c0 + 2 * lane + 1) * K + jj] : 0.f;
  }
  const float b0 = bias[c0 + 2 * lane], b1 = bias[c0 + 2 * lane + 1];
  if (blockIdx.y == 0 && reg != nullptr) {  // register tokens on the left (vp.py:425), fp32
    for (int i = threadIdx.x; i < (int)R * (kCT / 4); i += blockDim.x) {
      const int r = i / (kCT / 4), q = i % (kCT / 4);
      *reinterpret_cast<float4*>(y + (b * (R + N) + r) * C + c0 + q * 4) = *reinterpret_cast<const float4*>(reg + r * C + c0 + q * 4);
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int g = 0; g < 32 / kGroup; ++g) {
    const int base = warp * 32 + g * kGroup;  // first output row of this block (tile-relative); input row = base + i
    float a0[kGroup], a1[kGroup];
#pragma unroll
    for (int o = 0; o < kGroup; ++o) a0[o] = b0, a1[o] = b1;
    float2 centre[kGroup];
#pragma unroll
    for (int i = 0; i < kGroup + kHalo; ++i) {
      const float2 raw = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&tile[(base + i) * 32 + lane]));
      const float m = mrow[base + i];
      const float v0 = raw.x * m, v1 = raw.y * m;
      if (i >= kHalo / 2 && i < kHalo / 2 + kGroup) centre[i - kHalo / 2] = raw;  // un-masked x for the residual
#pragma unroll
      for (int o = 0; o < kGroup; ++o) {
        const int j = i - o;
        if (j >= 0 && j < kKMax) {
          a0[o] = fmaf(w0[j], v0, a0[o]);
          a1[o] = fmaf(w1[j], v1, a1[o]);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < kGroup; ++o) {
      const int64_t n = n0 + base + o;
      if (n < N) {
        const float m = mrow[base + o + kHalo / 2];
        if (pre != nullptr) *reinterpret_cast<__nv_bfloat162*>(pre + (b * N + n) * C + c0 + 2 * lane) = f2bf(a0[o], a1[o]);
        float2 out;
        out.x = fmaf(gelu_f(a0[o]), m, centre[o].x);
        out.y = fmaf(gelu_f(a1[o]), m, centre[o].y);
        *reinterpret_cast<float2*>(y + (b * (R + N) + R + n) * C + c0 + 2 * lane) = out;
      }
    }
  }
}

// backward: g = dy * gelu'(pre) * m ;  dx[n] = dy[n] + m[n] * sum_j w[j] g[n + 15 - j] ;
//           dw[j] += sum_n g[n] * (x m)[n + j - 15] ;  dbias += sum_n g[n] ;  dreg[r] += sum_b dy[b, r]
// grid.z strides the batch so each CTA keeps its dw partial sums in registers over several samples.
__global__ void __launch_bounds__(128) convpos_bwd_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ pre,
                                                           const float* __restrict__ w, const uint8_t* __restrict__ mask,
                                                           const float* __restrict__ dy, uint16_t* __restrict__ dx,
                                                           float* __restrict__ dw, float* __restrict__ dbias,
                                                           float* __restrict__ dreg, int64_t B, int64_t N, int C, int K,
                                                           int64_t R) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint32_t* xt = reinterpret_cast<uint32_t*>(smem_raw);                      // x tile, bf16x2   [158][32]
  float2* gt = reinterpret_cast<float2*>(smem_raw + kRowsIn * 32 * 4);        // g tile, fp32 x2  [158][32]
  float* mrow = reinterpret_cast<float*>(smem_raw + kRowsIn * 32 * 12);       // [158]
  float* red = reinterpret_cast<float*>(gt);                                  // reused after the loop: [4][32][64]
  const int c0 = blockIdx.x * kCT;
  const int64_t n0 = (int64_t)blockIdx.y * kNT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pad = (kKMax - K) / 2;
  float w0[kKMax], w1[kKMax], dw0[kKMax], dw1[kKMax];
#pragma unroll
  for (int j = 0; j < kKMax; ++j) {
    const int jj = j - pad;
    const bool in = jj >= 0 && jj < K;
    w0[j] = in ? w[(c0 + 2 * lane) * K + jj] : 0.f;
    w1[j] = in ? w[(c0 + 2 * lane + 1) * K + jj] : 0.f;
    dw0[j] = dw1[j] = 0.f;
  }
  float db0 = 0.f, db1 = 0.f;
  for (int64_t b = blockIdx.z; b < B; b += gridDim.z) {
    __syncthreads();
    stage_tile(x, xt, b, n0, c0, N, C);
    for (int r = threadIdx.x; r < kRowsIn; r += blockDim.x) {
      const int64_t n = n0 - kHalo / 2 + r;
      mrow[r] = (n >= 0 && n < N && (mask == nullptr || mask[b * N + n])) ? 1.f : 0.f;
    }
    __syncthreads();
    // g tile (with halo): one (row, channel pair) per thread-iteration, coalesced along channels
    for (int i = threadIdx.x; i < kRowsIn * 32; i += blockDim.x) {
      const int r = i >> 5, cp = i & 31;
      const int64_t n = n0 - kHalo / 2 + r;
      float2 g = make_float2(0.f, 0.f);
      if (n >= 0 && n < N && mrow[r] != 0.f) {
        const float2 d = *reinterpret_cast<const float2*>(dy + (b * (R + N) + R + n) * C + c0 + 2 * cp);
        const float2 u = bf2f(*reinterpret_cast<const __nv_bfloat162*>(pre + (b * N + n) * C + c0 + 2 * cp));
        g.x = d.x * gelu_grad_f(u.x);
        g.y = d.y * gelu_grad_f(u.y);
      }
      gt[i] = g;
    }
    if (blockIdx.y == 0 && dreg != nullptr) {
      for (int i = threadIdx.x; i < (int)R * kCT; i += blockDim.x) {
        const int r = i / kCT, c = i % kCT;
        atomicAdd(dreg + r * C + c0 + c, dy[(b * (R + N) + r) * C + c0 + c]);
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int gi = 0; gi < 32 / kGroup; ++gi) {
      const int base = warp * 32 + gi * kGroup;
      float a0[kGroup], a1[kGroup];
      float2 gc[kGroup];
#pragma unroll
      for (int o = 0; o < kGroup; ++o) {
        a0[o] = a1[o] = 0.f;
        gc[o] = gt[(base + o + kHalo / 2) * 32 + lane];
        db0 += gc[o].x;
        db1 += gc[o].y;
      }
#pragma unroll
      for (int i = 0; i < kGroup + kHalo; ++i) {
        const float2 gv = gt[(base + i) * 32 + lane];
        const float2 raw = bf2f(*reinterpret_cast<const __nv_bfloat162*>(&xt[(base + i) * 32 + lane]));
        const float m = mrow[base + i];
        const float x0 = raw.x * m, x1 = raw.y * m;
#pragma unroll
        for (int o = 0; o < kGroup; ++o) {
          const int j = i - o;  // input row (base+i) = output row (base+o) + j - 15
          if (j >= 0 && j < kKMax) {
            // dx[base+o] += w[30-j] * g[(base+o) + 15 - (30-j)] = w[30-j] * g[tile row base+o+j] = g[base+i]
            a0[o] = fmaf(w0[kKMax - 1 - j], gv.x, a0[o]);
            a1[o] = fmaf(w1[kKMax - 1 - j], gv.y, a1[o]);
            // dw[j] += g[out row base+o] * (x m)[in row base+i]
            dw0[j] = fmaf(gc[o].x, x0, dw0[j]);
            dw1[j] = fmaf(gc[o].y, x1, dw1[j]);
          }
        }
      }
#pragma unroll
      for (int o = 0; o < kGroup; ++o) {
        const int64_t n = n0 + base + o;
        if (n < N) {
          const float m = mrow[base + o + kHalo / 2];
          const float2 d = *reinterpret_cast<const float2*>(dy + (b * (R + N) + R + n) * C + c0 + 2 * lane);
          *reinterpret_cast<__nv_bfloat162*>(dx + (b * N + n) * C + c0 + 2 * lane) = f2bf(fmaf(a0[o], m, d.x), fmaf(a1[o], m, d.y));
        }
      }
    }
  }
  // reduce the 4 warps (same channels) through smem (the g tile is dead now), then one atomic per (channel, tap) per CTA
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kKMax; ++j) {
    red[(warp * 32 + j) * kCT + 2 * lane] = dw0[j];
    red[(warp * 32 + j) * kCT + 2 * lane + 1] = dw1[j];
  }
  red[(warp * 32 + kKMax) * kCT + 2 * lane] = db0;
  red[(warp * 32 + kKMax) * kCT + 2 * lane + 1] = db1;
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * kCT; i += blockDim.x) {
    const int j = i / kCT, c = i % kCT;
    const float sum = red[(0 * 32 + j) * kCT + c] + red[(1 * 32 + j) * kCT + c] + red[(2 * 32 + j) * kCT + c] +
                      red[(3 * 32 + j) * kCT + c];
    if (j < kKMax) {
      const int jj = j - pad;
      if (jj >= 0 && jj < K) atomicAdd(dw + (c0 + c) * K + jj, sum);
    } else {
      atomicAdd(dbias + c0 + c, sum);
    }
  }
}

}  // namespace vbx

using namespace vbx;

extern "C" int vbx_convpos_fwd(const uint16_t* x, const float* w, const float* bias, const uint8_t* mask, const float* reg,
                               float* y, uint16_t* pre, int64_t B, int64_t N, int64_t C, int64_t K, int64_t R, void* stream) {
  VBX_REQUIRE(x && w && bias && y && (R == 0 || reg), VBX_E_NULL);
  VBX_REQUIRE(B > 0 && B < 65536 && N > 0 && C > 0 && C % kCT == 0 && K > 0 && (K & 1) && R >= 0, VBX_E_SHAPE);
  VBX_REQUIRE(K <= kKMax, VBX_E_UNSUPPORTED);
  VBX_REQUIRE(VBX_ALIGNED16(x) && VBX_ALIGNED16(y) && (!reg || VBX_ALIGNED16(reg)) && (!pre || VBX_ALIGNED16(pre)), VBX_E_ALIGN);
  dim3 grid((unsigned)(C / kCT), (unsigned)((N + kNT - 1) / kNT), (unsigned)B);
  convpos_fwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, w, bias, mask, reg, y, pre, N, (int)C, (int)K, R);
  return VBX_LAUNCH_RC();
}

extern "C" int vbx_convpos_bwd(const uint16_t* x, const uint16_t* pre, const float* w, const uint8_t* mask, const float* dy,
                               uint16_t* dx, float* dw, float* dbias, float* dreg, int64_t B, int64_t N, int64_t C, int64_t K,
                               int64_t R, void* stream) {
  VBX_REQUIRE(x && pre && w && dy && dx && dw && dbias && (R == 0 || dreg), VBX_E_NULL);
  VBX_REQUIRE(B > 0 && N > 0 && C > 0 && C % kCT == 0 && K > 0 && (K & 1) && R >= 0, VBX_E_SHAPE);
  VBX_REQUIRE(K <= kKMax, VBX_E_UNSUPPORTED);
  VBX_REQUIRE(VBX_ALIGNED16(x) && VBX_ALIGNED16(pre) && VBX_ALIGNED16(dy) && VBX_ALIGNED16(dx), VBX_E_ALIGN);
  dim3 grid((unsigned)(C / kCT), (unsigned)((N + kNT - 1) / kNT), (unsigned)(B < 8 ? B : 8));
  const int smem = kRowsIn * 32 * 12 + kRowsIn * 4;
  cudaFuncSetAttribute(convpos_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  convpos_bwd_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(x, pre, w, mask, dy, dx, dw, dbias, dreg, B, N, (int)C, (int)K, R);
  return VBX_LAUNCH_RC();
}
