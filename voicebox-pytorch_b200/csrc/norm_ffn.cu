// HBM-bound fused passes of the Voicebox trunk: residual-add + (adaptive) RMSNorm, GEGLU.
// One read and one write of every activation per pass, 16-byte vector accesses, fp32 math.
#include "common.cuh"

namespace vbx {

// Element ownership of the row kernels: within chunk c (256 elements) lane l owns the two 4-element groups starting at
// c*256 + l*4 and c*256 + 128 + l*4.  Every warp-level access is then ONE fully coalesced segment: 512 B for fp32 (16 B per
// lane), 256 B for bf16 (8 B per lane).  The first version gave each lane 8 CONTIGUOUS elements: its two 16-byte fp32
// accesses each touched half of 32 different 32-byte sectors, doubling the L2 <-> SM sector traffic of x / dx (0.71-0.75 of
// the HBM roofline in round 1 although DRAM bytes matched the algorithmic count).
VBX_DEVINL int grp(int c, int hf, int lane) { return c * 256 + hf * 128 + lane * 4; }
// ------------------------------------------------------------------------------------------------------------------
// residual add + (adaptive) RMSNorm, forward.   One warp per token row; element ownership: grp() above.
// ------------------------------------------------------------------------------------------------------------------
#ifndef VBX_ADARMS_FWD_ROWS
#define VBX_ADARMS_FWD_ROWS 1
#endif
constexpr int kAdarmsFwdRows = VBX_ADARMS_FWD_ROWS;   // rows per warp (A/B knob: 1 = one row per warp, the round-1 shape)
template <int C>
__global__ void __launch_bounds__(256, (C <= 4) ? 4 : 2) adarms_fwd_kernel(const float* __restrict__ x_in, int64_t xbs, int64_t row0,
                                                          const uint16_t* __restrict__ branch,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          int per_batch, float* x_out, uint16_t* __restrict__ h,
                                                          float* __restrict__ rstd, int64_t B, int64_t rows, int D) {
  const int lane = threadIdx.x & 31;
  const float sqrt_d = sqrtf((float)D);
  // one block per 8 * kAdarmsFwdRows consecutive rows, a warp takes every 8th of them; the hardware block scheduler balances the
  // tail (a grid-stride loop over all rows costs up to 1/7 in imbalance).  With more than one row per warp, lane 0 announces the
  // warp's NEXT row to L2 before it starts on the current one.
  const int nwarp = blockDim.x >> 5;
#pragma unroll 1
  for (int rr = 0; rr < kAdarmsFwdRows; ++rr) {
  const int64_t row = ((int64_t)blockIdx.x * kAdarmsFwdRows + rr) * nwarp + (threadIdx.x >> 5);
  if (row < B * rows) {
    const int64_t b = row / rows, r = row - b * rows;
    const float* xi = x_in + b * xbs + (row0 + r) * D;
    float v[C][8];
    float ss = 0.f;
    const uint16_t* bri = branch != nullptr ? branch + b * xbs + (row0 + r) * D : nullptr;
    if (kAdarmsFwdRows > 1 && rr + 1 < kAdarmsFwdRows) {
      const int64_t nxt = row + nwarp;
      if (lane == 0 && nxt < B * rows && (D & 7) == 0) {
        const int64_t b2 = nxt / rows, r2 = nxt - b2 * rows;
        prefetch_l2_bulk(x_in + b2 * xbs + (row0 + r2) * D, (uint32_t)D * 4u);
        if (branch != nullptr) prefetch_l2_bulk(branch + b2 * xbs + (row0 + r2) * D, (uint32_t)D * 2u);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int e = grp(c, hf, lane);
        if (e < D) {
          ld4f(xi + e, &v[c][hf * 4], true);
          if (bri != nullptr) {
            float br[4];
            ld4h(bri + e, br);
#pragma unroll
            for (int i = 0; i < 4; ++i) v[c][hf * 4 + i] += br[i];
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) ss = fmaf(v[c][hf * 4 + i], v[c][hf * 4 + i], ss);
        }
      }
    }
    if (x_out != nullptr) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int e = grp(c, hf, lane);
          if (e < D) st4f(x_out + row * D + e, &v[c][hf * 4]);
        }
      }
    }
    ss = warp_sum(ss);
    const float rinv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps (vp.py:247)
    const float scale = sqrt_d * rinv;                    // ... then * sqrt(dim)
    if (rstd != nullptr && lane == 0) rstd[row] = rinv;
    const float* g = gamma + (per_batch ? b * D : 0);
    const float* bt = beta ? beta + (per_batch ? b * D : 0) : nullptr;
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int e = grp(c, hf, lane);
        if (e < D) {
          float o[4];
          const float4 g0 = *reinterpret_cast<const float4*>(g + e);
          const float gg[4] = {g0.x, g0.y, g0.z, g0.w};
          if (bt) {
            const float4 b0 = *reinterpret_cast<const float4*>(bt + e);
            const float bb[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = fmaf(v[c][hf * 4 + i] * scale, gg[i], bb[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = v[c][hf * 4 + i] * scale * gg[i];
          }
          st4h(h + row * D + e, o);
        }
      }
    }
  }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward.  CTA = (batch b, chunk of RC rows); warps stride the chunk; dgamma/dbeta reduced warp -> smem -> global.
// (lane l owns 8 CONTIGUOUS elements (c*32+l)*8.. here: fp32 rows go through the 256-bit ld8f / st8f, one request per sector;
//  the grp() mapping of the forward cost 22 % in this kernel -- twice the load / store instructions for the bf16 streams)
//   g = dh*gamma*sqrt(D);  xh = x/||x||;  dx = (g - xh*(xh.g))/||x|| + dx_res;  dgamma += dh*xh*sqrt(D);  dbeta += dh
// ------------------------------------------------------------------------------------------------------------------
// rows per CTA are chosen on the host so that the CTA count is just under a whole number of waves (2 CTAs per SM)

// dgamma/dbeta partial sums live in per-warp PRIVATE shared-memory slices (no atomics, no barriers in the row loop),
// laid out [warp][2][c][half][lane] float4 so every access is conflict-free; this keeps the kernel near 100
// registers (2 CTAs/SM) instead of 254 with register accumulators.
template <int C>
__global__ void __launch_bounds__(256, 2) adarms_bwd_kernel(const float* __restrict__ x, int64_t xbs, int64_t row0,
                                                             const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                             int per_batch, const uint16_t* __restrict__ dh,
                                                             const float* __restrict__ dx_res, float* __restrict__ dx,
                                                             uint16_t* __restrict__ dbranch, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, int64_t rows, int D, int rows_per_cta) {
  extern __shared__ float4 red4[];  // [8 warps][2][C*2*32] float4
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kSlice = C * 2 * 32;  // float4 per (warp, which)
  float4* my_g = red4 + (warp * 2 + 0) * kSlice;
  float4* my_b = red4 + (warp * 2 + 1) * kSlice;
  const int64_t b = blockIdx.y;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r_end = min(rows, r_begin + rows_per_cta);
  const float sqrt_d = sqrtf((float)D);
#pragma unroll
  for (int j = 0; j < C * 2; ++j) my_g[j * 32 + lane] = my_b[j * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* g = gamma + (per_batch ? b * D : 0);

  for (int64_t r = r_begin + warp; r < r_end; r += (blockDim.x >> 5)) {
    const int64_t row = b * rows + r;
    const float* xi = x + b * xbs + (row0 + r) * D;
    if (lane == 0 && (D & 7) == 0) {   // this warp's next row; and the residual-stream gradient this row needs after the reduction
      if (dx_res != nullptr) prefetch_l2_bulk(dx_res + row * D, (uint32_t)D * 4u);
      const int64_t rn = r + (blockDim.x >> 5);
      if (rn < r_end) {
        prefetch_l2_bulk(xi + (int64_t)(blockDim.x >> 5) * D, (uint32_t)D * 4u);
        prefetch_l2_bulk(dh + (row + (blockDim.x >> 5)) * D, (uint32_t)D * 2u);
      }
    }
    const float rinv = rstd[row];
    const float s1 = sqrt_d * rinv;  // d h / d x leading factor
    float xv[C][8], gv[C][8];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int e = (c * 32 + lane) * 8;
      if (e < D) {
        ld8f(xi + e, xv[c]);
        unpack8(ldg_nc_16(dh + row * D + e), gv[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int e = (c * 32 + lane) * 8;
      if (e < D) {
        const float4 g0 = *reinterpret_cast<const float4*>(g + e), g1 = *reinterpret_cast<const float4*>(g + e + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float4 ag = my_g[(c * 2 + hf) * 32 + lane], ab = my_b[(c * 2 + hf) * 32 + lane];
          float* pg = reinterpret_cast<float*>(&ag);
          float* pb = reinterpret_cast<float*>(&ab);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = hf * 4 + k;
            pb[k] += gv[c][i];                               // dbeta  += dh
            pg[k] = fmaf(gv[c][i] * xv[c][i], s1, pg[k]);    // dgamma += dh * xhat * sqrt(D)
            gv[c][i] *= gg[i];                               // dh * gamma
            dot = fmaf(gv[c][i], xv[c][i], dot);
          }
          my_g[(c * 2 + hf) * 32 + lane] = ag;
          my_b[(c * 2 + hf) * 32 + lane] = ab;
        }
      }
    }
    dot = warp_sum(dot);
    const float s2 = s1 * rinv * rinv * dot;  // projection on x
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int e = (c * 32 + lane) * 8;
      if (e < D) {
        float o[8];
        if (dx_res != nullptr) ld8f(dx_res + row * D + e, o);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = fmaf(gv[c][i], s1, -xv[c][i] * s2);
          o[i] = (dx_res != nullptr) ? o[i] + d : d;
        }
        st8f(dx + row * D + e, o);
        if (dbranch != nullptr) stg_16(dbranch + row * D + e, pack8(o));
      }
    }
  }
  __syncthreads();
  // cross-warp reduce + one global atomic per column per CTA
  const float* redf = reinterpret_cast<const float*>(red4);
  for (int col = threadIdx.x; col < D; col += blockDim.x) {
    // element col = (c*32+l)*8 + hf*4 + k  ->  private slot ((c*2+hf)*32 + l)*4 + k
    const int c = col >> 8, l = (col >> 3) & 31, hf = (col >> 2) & 1, k = col & 3;
    const int slot = ((c * 2 + hf) * 32 + l) * 4 + k;
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      sg += redf[(w * 2 + 0) * kSlice * 4 + slot];
      sb += redf[(w * 2 + 1) * kSlice * 4 + slot];
    }
    atomicAdd(dgamma + (per_batch ? b * D : 0) + col, sg);
    if (dbeta != nullptr) atomicAdd(dbeta + (per_batch ? b * D : 0) + col, sb);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// GEGLU: out = gelu_erf(gate) * value, value = h[:, :Fp], gate = h[:, Fp:]
// ------------------------------------------------------------------------------------------------------------------
// Thread layout for both GEGLU kernels: threadIdx/blockIdx.x pick a fixed 8-element column vector, blockIdx.y strides the
// rows -- no integer division anywhere (a 64-bit div per grid-stride iteration made the first version issue-bound), and
// four rows are in flight per thread.
constexpr int kGegluRowsUnroll = 4;

__global__ void __launch_bounds__(128) geglu_fwd_kernel(const uint16_t* __restrict__ h, uint16_t* __restrict__ out, int64_t T,
                                                         int Fp) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) << 3;
  if (c >= Fp) return;
  const int64_t row_stride = (int64_t)gridDim.y * kGegluRowsUnroll;
  for (int64_t t0 = (int64_t)blockIdx.y * kGegluRowsUnroll; t0 < T; t0 += row_stride) {
    uint4 uv[kGegluRowsUnroll], ug[kGegluRowsUnroll];
#pragma unroll
    for (int u = 0; u < kGegluRowsUnroll; ++u) {
      if (t0 + u < T) {
        uv[u] = ldg_nc_16(h + (t0 + u) * 2 * Fp + c);
        ug[u] = ldg_nc_16(h + (t0 + u) * 2 * Fp + Fp + c);
      }
    }
#pragma unroll
    for (int u = 0; u < kGegluRowsUnroll; ++u) {
      if (t0 + u < T) {
        float val[8], gate[8], o[8];
        unpack8(uv[u], val);
        unpack8(ug[u], gate);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = gelu_f(gate[k]) * val[k];
        stg_16(out + (t0 + u) * Fp + c, pack8(o));
      }
    }
  }
}

// dbias (f32 [2*Fp], may be NULL): column sums of dh = the bias gradient of the Linear that produced h (vp.py:345).  The
// thread <-> column mapping is fixed, so the sums live in 16 registers and cost one atomic per column per thread at the end
// -- instead of a separate reduction pass re-reading all of dh (733 MB per layer at cfg3).
__global__ void __launch_bounds__(128) geglu_bwd_kernel(const uint16_t* __restrict__ h, const uint16_t* __restrict__ dout,
                                                         uint16_t* __restrict__ dh, float* __restrict__ dbias, int64_t T, int Fp) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) << 3;
  if (c >= Fp) return;
  float sv[8], sg[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) sv[k] = sg[k] = 0.f;
  const int64_t row_stride = (int64_t)gridDim.y * kGegluRowsUnroll;
  for (int64_t t0 = (int64_t)blockIdx.y * kGegluRowsUnroll; t0 < T; t0 += row_stride) {
    uint4 uv[kGegluRowsUnroll], ug[kGegluRowsUnroll], ud[kGegluRowsUnroll];
#pragma unroll
    for (int u = 0; u < kGegluRowsUnroll; ++u) {
      if (t0 + u < T) {
        uv[u] = ldg_nc_16(h + (t0 + u) * 2 * Fp + c);
        ug[u] = ldg_nc_16(h + (t0 + u) * 2 * Fp + Fp + c);
        ud[u] = ldg_nc_16(dout + (t0 + u) * Fp + c);
      }
    }
#pragma unroll
    for (int u = 0; u < kGegluRowsUnroll; ++u) {
      if (t0 + u < T) {
        float val[8], gate[8], d[8], dv[8], dg[8];
        unpack8(uv[u], val);
        unpack8(ug[u], gate);
        unpack8(ud[u], d);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float e;
          const float cdf = normal_cdf(gate[k], e);
          dv[k] = d[k] * gate[k] * cdf;                                         // d value
          dg[k] = d[k] * val[k] * fmaf(gate[k] * 0.3989422804014327f, e, cdf);  // d gate
        }
        const uint4 pv = pack8(dv), pg = pack8(dg);
        stg_16(dh + (t0 + u) * 2 * Fp + c, pv);
        stg_16(dh + (t0 + u) * 2 * Fp + Fp + c, pg);
        if (dbias != nullptr) {  // sum what was actually stored (bf16-rounded), as a reduction over dh would
          float rv[8], rg[8];
          unpack8(pv, rv);
          unpack8(pg, rg);
#pragma unroll
          for (int k = 0; k < 8; ++k) sv[k] += rv[k], sg[k] += rg[k];
        }
      }
    }
  }
  if (dbias != nullptr) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      atomicAdd(dbias + c + k, sv[k]);
      atomicAdd(dbias + Fp + c + k, sg[k]);
    }
  }
}

}  // namespace vbx

using namespace vbx;

extern "C" int vbx_adarms_fwd(const float* x_in, int64_t x_batch_stride, int64_t row0, const uint16_t* branch,
                              const float* gamma, const float* beta, int per_batch, float* x_out, uint16_t* h,
                              float* rstd, int64_t B, int64_t rows, int64_t D, void* stream) {
  VBX_REQUIRE(x_in && gamma && h, VBX_E_NULL);
  VBX_REQUIRE(B > 0 && rows > 0 && D > 0 && D % 8 == 0 && x_batch_stride % 4 == 0, VBX_E_SHAPE);
  VBX_REQUIRE(D <= 2048, VBX_E_UNSUPPORTED);
  VBX_REQUIRE(VBX_ALIGNED16(x_in) && VBX_ALIGNED16(gamma) && VBX_ALIGNED16(h) && (!branch || VBX_ALIGNED16(branch)) &&
                  (!beta || VBX_ALIGNED16(beta)) && (!x_out || VBX_ALIGNED16(x_out)),
              VBX_E_ALIGN);
  if (x_out != nullptr && x_out == x_in) VBX_REQUIRE(x_batch_stride == rows * D && row0 == 0, VBX_E_SHAPE);
  const int64_t nblk = (B * rows + 8 * kAdarmsFwdRows - 1) / (8 * kAdarmsFwdRows);
  VBX_REQUIRE(nblk < (1ll << 31), VBX_E_SHAPE);
  const int grid = (int)nblk;
  cudaStream_t s = (cudaStream_t)stream;
  const int C = (int)((D + 255) / 256);
#define LAUNCH(CC)                                                                                                          \
  adarms_fwd_kernel<CC><<<grid, 256, 0, s>>>(x_in, x_batch_stride, row0, branch, gamma, beta, per_batch, x_out, h, rstd, B, \
                                             rows, (int)D)
  if (C <= 1) LAUNCH(1);
  else if (C <= 2) LAUNCH(2);
  else if (C <= 4) LAUNCH(4);
  else LAUNCH(8);
#undef LAUNCH
  return VBX_LAUNCH_RC();
}

extern "C" int vbx_adarms_bwd(const float* x, int64_t x_batch_stride, int64_t row0, const float* rstd, const float* gamma,
                              int per_batch,
                              const uint16_t* dh, const float* dx_res, float* dx, uint16_t* dbranch, float* dgamma,
                              float* dbeta, int64_t B, int64_t rows, int64_t D, void* stream) {
  VBX_REQUIRE(x && rstd && gamma && dh && dx && dgamma, VBX_E_NULL);
  VBX_REQUIRE(B > 0 && rows > 0 && D > 0 && D % 8 == 0 && x_batch_stride % 4 == 0 && B < 65536, VBX_E_SHAPE);
  VBX_REQUIRE(D <= 2048, VBX_E_UNSUPPORTED);
  VBX_REQUIRE(VBX_ALIGNED16(x) && VBX_ALIGNED16(dh) && VBX_ALIGNED16(dx) && (!dx_res || VBX_ALIGNED16(dx_res)) &&
                  (!dbranch || VBX_ALIGNED16(dbranch)),
              VBX_E_ALIGN);
  // pick rows/CTA in [32,128] minimising ceil(waves) * rows  (time ~ number of waves x rows per CTA)
  int best_rc = 64;
  double best_cost = 1e30;
  for (int rc = 32; rc <= 128; ++rc) {
    const int64_t ctas = B * ((rows + rc - 1) / rc);
    const int64_t waves = (ctas + 2 * kNumSM - 1) / (2 * kNumSM);
    const double cost = (double)waves * rc;
    if (cost < best_cost - 1e-9) best_cost = cost, best_rc = rc;
  }
  dim3 grid((unsigned)((rows + best_rc - 1) / best_rc), (unsigned)B);
  cudaStream_t s = (cudaStream_t)stream;
  const int C = (int)((D + 255) / 256);
#define LAUNCH(CC)                                                                                                     \
  {                                                                                                                    \
    const size_t smem = 8 * 2 * (CC * 2 * 32) * sizeof(float4);                                                        \
    cudaFuncSetAttribute(adarms_bwd_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);               \
    adarms_bwd_kernel<CC><<<grid, 256, smem, s>>>(x, x_batch_stride, row0, rstd, gamma, per_batch, dh, dx_res, dx,     \
                                                  dbranch, dgamma, dbeta, rows, (int)D, best_rc);                      \
  }
  if (C <= 1) LAUNCH(1)
  else if (C <= 2) LAUNCH(2)
  else if (C <= 4) LAUNCH(4)
  else LAUNCH(8)
#undef LAUNCH
  return VBX_LAUNCH_RC();
}

extern "C" int vbx_geglu_fwd(const uint16_t* h, uint16_t* out, int64_t T, int64_t Fp, void* stream) {
  VBX_REQUIRE(h && out, VBX_E_NULL);
  VBX_REQUIRE(T > 0 && Fp > 0 && Fp % 8 == 0, VBX_E_SHAPE);
  VBX_REQUIRE(VBX_ALIGNED16(h) && VBX_ALIGNED16(out), VBX_E_ALIGN);
  const unsigned gx = (unsigned)((Fp / 8 + 127) / 128);
  int64_t gy = (T + kGegluRowsUnroll - 1) / kGegluRowsUnroll;
  const int64_t cap = (int64_t)kNumSM * 16 / gx + 1;  // ~16 CTAs of 128 threads per SM in flight, then stride
  if (gy > cap) gy = cap;
  geglu_fwd_kernel<<<dim3(gx, (unsigned)gy), 128, 0, (cudaStream_t)stream>>>(h, out, T, (int)Fp);
  return VBX_LAUNCH_RC();
}

extern "C" int vbx_geglu_bwd(const uint16_t* h, const uint16_t* dout, uint16_t* dh, float* dbias, int64_t T, int64_t Fp,
                             void* stream) {
  VBX_REQUIRE(h && dout && dh, VBX_E_NULL);
  VBX_REQUIRE(T > 0 && Fp > 0 && Fp % 8 == 0, VBX_E_SHAPE);
  VBX_REQUIRE(VBX_ALIGNED16(h) && VBX_ALIGNED16(dout) && VBX_ALIGNED16(dh), VBX_E_ALIGN);
  const unsigned gx = (unsigned)((Fp / 8 + 127) / 128);
  int64_t gy = (T + kGegluRowsUnroll - 1) / kGegluRowsUnroll;
  const int64_t cap = (int64_t)kNumSM * (dbias ? 6 : 16) / gx + 1;  // fewer, longer threads when they end in atomics
  if (gy > cap) gy = cap;
  geglu_bwd_kernel<<<dim3(gx, (unsigned)gy), 128, 0, (cudaStream_t)stream>>>(h, dout, dh, dbias, T, (int)Fp);
  return VBX_LAUNCH_RC();
}
