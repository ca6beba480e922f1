// Fused gradient-clip + Adam update over the FLAT parameter / gradient / moment buffers (trainer.py:274-278: clip_grad_norm_,
// optim.step): one HBM-bound pass, 16-byte accesses, 28 B/element (+2 B when the bf16 shadow copy is emitted).
#include "common.cuh"

namespace vbx {

struct AdamConsts {
  float beta1, beta2, one_minus_beta1, one_minus_beta2, step_size, inv_sqrt_bc2, eps, l2, decay;
};

VBX_DEVINL void adam_one(float& p, float g, float& m, float& v, const AdamConsts& c, float inv_scale) {
  g *= inv_scale;
  g = fmaf(c.l2, p, g);                       // L2 weight decay (torch.optim.Adam); l2 = 0 otherwise
  p *= c.decay;                               // decoupled weight decay (AdamW): 1 - lr*wd; 1 otherwise
  m = fmaf(c.one_minus_beta1, g - m, m);      // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(c.one_minus_beta2 * g, g, c.beta2 * v);
  const float denom = fmaf(sqrtf(v), c.inv_sqrt_bc2, c.eps);
  p = fmaf(-c.step_size, m / denom, p);
}

__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, uint16_t* __restrict__ p_bf16, int64_t n,
                                                         AdamConsts c, const float* __restrict__ grad_scale,
                                                         const float* __restrict__ found_inf) {
  if (found_inf != nullptr && found_inf[0] != 0.f) return;  // skipped step (the fused torch optimizers' found_inf contract)
  const float inv_scale = grad_scale != nullptr ? 1.0f / grad_scale[0] : 1.0f;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 pu = ldg_16(p + 4 * i), gu = ldg_nc_16(g + 4 * i), mu = ldg_16(m + 4 * i), vu = ldg_16(v + 4 * i);
    float pp[4] = {__uint_as_float(pu.x), __uint_as_float(pu.y), __uint_as_float(pu.z), __uint_as_float(pu.w)};
    const float gg[4] = {__uint_as_float(gu.x), __uint_as_float(gu.y), __uint_as_float(gu.z), __uint_as_float(gu.w)};
    float mm[4] = {__uint_as_float(mu.x), __uint_as_float(mu.y), __uint_as_float(mu.z), __uint_as_float(mu.w)};
    float vv[4] = {__uint_as_float(vu.x), __uint_as_float(vu.y), __uint_as_float(vu.z), __uint_as_float(vu.w)};
#pragma unroll
    for (int k = 0; k < 4; ++k) adam_one(pp[k], gg[k], mm[k], vv[k], c, inv_scale);
    stg_16(p + 4 * i, make_uint4(__float_as_uint(pp[0]), __float_as_uint(pp[1]), __float_as_uint(pp[2]), __float_as_uint(pp[3])));
    stg_16(m + 4 * i, make_uint4(__float_as_uint(mm[0]), __float_as_uint(mm[1]), __float_as_uint(mm[2]), __float_as_uint(mm[3])));
    stg_16(v + 4 * i, make_uint4(__float_as_uint(vv[0]), __float_as_uint(vv[1]), __float_as_uint(vv[2]), __float_as_uint(vv[3])));
    if (p_bf16 != nullptr) {
      __nv_bfloat162 lo = f2bf(pp[0], pp[1]), hi = f2bf(pp[2], pp[3]);
      *reinterpret_cast<uint2*>(p_bf16 + 4 * i) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {  // tail of a bucket whose size is not a multiple of 4
    const int64_t i = (n4 << 2) + threadIdx.x;
    float pp = p[i], mm = m[i], vv = v[i];
    adam_one(pp, g[i], mm, vv, c, inv_scale);
    p[i] = pp, m[i] = mm, v[i] = vv;
    if (p_bf16 != nullptr) p_bf16[i] = __bfloat16_as_ushort(__float2bfloat16_rn(pp));
  }
}

}  // namespace vbx

using namespace vbx;

extern "C" int vbx_adam_step(float* p, const float* g, float* m, float* v, uint16_t* p_bf16, int64_t n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int decoupled, int64_t step, const float* grad_scale,
                             const float* found_inf, void* stream) {
  VBX_REQUIRE(p && g && m && v, VBX_E_NULL);
  VBX_REQUIRE(n > 0 && step >= 1 && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && lr >= 0.f && eps >= 0.f &&
                  weight_decay >= 0.f,
              VBX_E_SHAPE);
  VBX_REQUIRE(VBX_ALIGNED16(p) && VBX_ALIGNED16(g) && VBX_ALIGNED16(m) && VBX_ALIGNED16(v) &&
                  (!p_bf16 || (reinterpret_cast<uintptr_t>(p_bf16) & 7) == 0),
              VBX_E_ALIGN);
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  AdamConsts c;
  c.beta1 = beta1, c.beta2 = beta2, c.one_minus_beta1 = 1.0f - beta1, c.one_minus_beta2 = 1.0f - beta2;
  c.step_size = (float)((double)lr / bc1);
  c.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  c.eps = eps;
  c.l2 = decoupled ? 0.f : weight_decay;
  c.decay = decoupled ? 1.0f - lr * weight_decay : 1.0f;
  adam_flat_kernel<<<grid_for(n / 4 + 1, 256, 8), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, p_bf16, n, c, grad_scale, found_inf);
  return VBX_LAUNCH_RC();
}
