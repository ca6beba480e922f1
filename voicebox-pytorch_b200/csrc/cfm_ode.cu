// Conditional-flow-matching elementwise passes: noise interpolation + conditioning mask + concat (to_embed input),
// masked-mean MSE (fwd, bwd), and the fixed-grid ODE stage combine.  All HBM-bound, 8 elements (16 B of bf16) per thread.
#include "common.cuh"

namespace vbx {

// emb[b,n,0:D] = bf16(w), emb[b,n,D:2D] = bf16(flow * !mask)       (vp.py:1408-1410, 1003, 1035, 1075-1076)
__global__ void __launch_bounds__(256) cfm_embed_kernel(const float* __restrict__ x0, const float* __restrict__ x1,
                                                         const float* __restrict__ times, const uint8_t* __restrict__ cmask,
                                                         float sigma, uint16_t* __restrict__ emb, int64_t N, int D,
                                                         int64_t total_vec) {
  const int vpr = D >> 3;
  const float oms = 1.0f - sigma;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tok = i / vpr;
    const int c = (int)(i - tok * vpr) << 3;
    const int64_t b = tok / N;
    const float t = times[b];
    const float keep = (cmask != nullptr && cmask[tok]) ? 0.f : 1.f;
    float a[8], z[8], w[8], fl[8];
    ld8f(x0 + tok * D + c, a);
    ld8f(x1 + tok * D + c, z);
    const float ca = 1.0f - oms * t;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      w[k] = ca * a[k] + t * z[k];           // same op order as the reference: (1-(1-s)t)*x0 + t*x1
      fl[k] = (z[k] - oms * a[k]) * keep;    // flow, zeroed where the conditioning is masked
    }
    stg_16(emb + tok * 2 * D + c, pack8(w));
    stg_16(emb + tok * 2 * D + D + c, pack8(fl));
  }
}

__global__ void __launch_bounds__(256) embed_concat_kernel(const float* __restrict__ x, const float* __restrict__ cond,
                                                            const uint8_t* __restrict__ cmask, uint16_t* __restrict__ emb,
                                                            int D, int64_t total_vec) {
  const int vpr = D >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tok = i / vpr;
    const int c = (int)(i - tok * vpr) << 3;
    float v[8];
    if (x != nullptr) {
      ld8f(x + tok * D + c, v);
      stg_16(emb + tok * 2 * D + c, pack8(v));
    }
    if (cond != nullptr) {
      const float keep = (cmask != nullptr && cmask[tok]) ? 0.f : 1.f;
      ld8f(cond + tok * D + c, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] *= keep;
      stg_16(emb + tok * 2 * D + D + c, pack8(v));
    }
  }
}

// one warp per token: mean_d (pred - target)^2, masked, accumulated per sample     (vp.py:1104-1112)
template <bool kBackward>
__global__ void __launch_bounds__(256) masked_mse_kernel(const uint16_t* __restrict__ pred, const float* __restrict__ tgt,
                                                          const float* __restrict__ x0, const float* __restrict__ x1,
                                                          float sigma, const uint8_t* __restrict__ lmask,
                                                          const float* __restrict__ coef, float* __restrict__ num,
                                                          uint16_t* __restrict__ dpred, int64_t B, int64_t N, int D) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float oms = 1.0f - sigma, inv_d = 1.0f / (float)D;
  for (int64_t tok = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tok < B * N; tok += nwarps) {
    const int64_t b = tok / N;
    const bool on = (lmask == nullptr) || lmask[tok];
    if (!on) {
      if (kBackward)
        for (int c = lane * 8; c < D; c += 256) stg_16(dpred + tok * D + c, make_uint4(0, 0, 0, 0));
      continue;  // warp-uniform
    }
    const float cf = kBackward ? coef[b] : 0.f;
    float acc = 0.f;
    for (int c = lane * 8; c < D; c += 256) {
      float p[8], t[8];
      unpack8(ldg_nc_16(pred + tok * D + c), p);
      if (tgt != nullptr) {
        ld8f(tgt + tok * D + c, t);
      } else {
        float a[8];
        ld8f(x0 + tok * D + c, a);
        ld8f(x1 + tok * D + c, t);
#pragma unroll
        for (int k = 0; k < 8; ++k) t[k] = t[k] - oms * a[k];
      }
      if (kBackward) {
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = cf * (p[k] - t[k]);
        stg_16(dpred + tok * D + c, pack8(o));
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float d = p[k] - t[k];
          acc = fmaf(d, d, acc);
        }
      }
    }
    if (!kBackward) {
      acc = warp_sum(acc);
      if (lane == 0) atomicAdd(num + b, acc * inv_d);
    }
  }
}

// y_out = y + a*f ; a = (t[i1]-t[i0]) or half of it, read from DEVICE memory      (torchdiffeq fixed-grid stage combine)
__global__ void __launch_bounds__(256) ode_axpy_kernel(const float* y, const uint16_t* __restrict__ f,
                                                        const float* __restrict__ t, int64_t i0, int64_t i1, int half,
                                                        float* y_out, uint16_t* __restrict__ emb, float* __restrict__ t_out,
                                                        int D, int64_t total_vec) {
  const float t0 = t[i0];
  const float dt = t[i1] - t0;
  const float a = half ? 0.5f * dt : dt;
  if (t_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) t_out[0] = t0 + 0.5f * dt;
  const int vpr = D >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tok = i / vpr;
    const int c = (int)(i - tok * vpr) << 3;
    float yv[8], fv[8];
    ld8f_rw(y + tok * D + c, yv);
    unpack8(ldg_nc_16(f + tok * D + c), fv);
#pragma unroll
    for (int k = 0; k < 8; ++k) yv[k] = yv[k] + fv[k] * a;  // y0 + f0*half_dt  /  y0 + dt*f1
    st8f(y_out + tok * D + c, yv);
    if (emb != nullptr) stg_16(emb + tok * 2 * D + c, pack8(yv));
  }
}

}  // namespace vbx

using namespace vbx;

extern "C" int vbx_cfm_embed(const float* x0, const float* x1, const float* times, const uint8_t* cond_mask, float sigma,
                             uint16_t* emb, int64_t B, int64_t N, int64_t D, void* stream) {
  VBX_REQUIRE(x0 && x1 && times && emb, VBX_E_NULL);
  VBX_REQUIRE(B > 0 && N > 0 && D > 0 && D % 8 == 0, VBX_E_SHAPE);
  VBX_REQUIRE(VBX_ALIGNED16(x0) && VBX_ALIGNED16(x1) && VBX_ALIGNED16(emb), VBX_E_ALIGN);
  const int64_t total = B * N * (D / 8);
  cfm_embed_kernel<<<grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(x0, x1, times, cond_mask, sigma, emb, N, (int)D,
                                                                             total);
  return VBX_LAUNCH_RC();
}

extern "C" int vbx_embed_concat(const float* x, const float* cond, const uint8_t* cond_mask, uint16_t* emb, int64_t B,
                                int64_t N, int64_t D, void* stream) {
  VBX_REQUIRE(emb && (x || cond), VBX_E_NULL);
  VBX_REQUIRE(B > 0 && N > 0 && D > 0 && D % 8 == 0, VBX_E_SHAPE);
  VBX_REQUIRE((!x || VBX_ALIGNED16(x)) && (!cond || VBX_ALIGNED16(cond)) && VBX_ALIGNED16(emb), VBX_E_ALIGN);
  const int64_t total = B * N * (D / 8);
  embed_concat_kernel<<<grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, cond, cond_mask, emb, (int)D, total);
  return VBX_LAUNCH_RC();
}

extern "C" int vbx_masked_mse_fwd(const uint16_t* pred, const float* tgt, const float* x0, const float* x1, float sigma,
                                  const uint8_t* loss_mask, float* num, int64_t B, int64_t N, int64_t D, void* stream) {
  VBX_REQUIRE(pred && num && (tgt || (x0 && x1)), VBX_E_NULL);
  VBX_REQUIRE(B > 0 && N > 0 && D > 0 && D % 8 == 0, VBX_E_SHAPE);
  VBX_REQUIRE(VBX_ALIGNED16(pred) && (!tgt || VBX_ALIGNED16(tgt)) && (!x0 || VBX_ALIGNED16(x0)) && (!x1 || VBX_ALIGNED16(x1)),
              VBX_E_ALIGN);
  masked_mse_kernel<false><<<grid_for(B * N, 8, 8), 256, 0, (cudaStream_t)stream>>>(pred, tgt, x0, x1, sigma, loss_mask, nullptr,
                                                                                   num, nullptr, B, N, (int)D);
  return VBX_LAUNCH_RC();
}

extern "C" int vbx_masked_mse_bwd(const uint16_t* pred, const float* tgt, const float* x0, const float* x1, float sigma,
                                  const uint8_t* loss_mask, const float* coef, uint16_t* dpred, int64_t B, int64_t N,
                                  int64_t D, void* stream) {
  VBX_REQUIRE(pred && coef && dpred && (tgt || (x0 && x1)), VBX_E_NULL);
  VBX_REQUIRE(B > 0 && N > 0 && D > 0 && D % 8 == 0, VBX_E_SHAPE);
  VBX_REQUIRE(VBX_ALIGNED16(pred) && VBX_ALIGNED16(dpred) && (!tgt || VBX_ALIGNED16(tgt)) && (!x0 || VBX_ALIGNED16(x0)) &&
                  (!x1 || VBX_ALIGNED16(x1)),
              VBX_E_ALIGN);
  masked_mse_kernel<true><<<grid_for(B * N, 8, 8), 256, 0, (cudaStream_t)stream>>>(pred, tgt, x0, x1, sigma, loss_mask, coef,
                                                                                  nullptr, dpred, B, N, (int)D);
  return VBX_LAUNCH_RC();
}

extern "C" int vbx_ode_axpy(const float* y, const uint16_t* f, const float* t, int64_t i0, int64_t i1, int half, float* y_out,
                            uint16_t* emb, float* t_out, int64_t B, int64_t N, int64_t D, void* stream) {
  VBX_REQUIRE(y && f && t && y_out, VBX_E_NULL);
  VBX_REQUIRE(B > 0 && N > 0 && D > 0 && D % 8 == 0 && i0 >= 0 && i1 >= 0, VBX_E_SHAPE);
  VBX_REQUIRE(VBX_ALIGNED16(y) && VBX_ALIGNED16(f) && VBX_ALIGNED16(y_out) && (!emb || VBX_ALIGNED16(emb)), VBX_E_ALIGN);
  const int64_t total = B * N * (D / 8);
  ode_axpy_kernel<<<grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(y, f, t, i0, i1, half, y_out, emb, t_out, (int)D,
                                                                            total);
  return VBX_LAUNCH_RC();
}
