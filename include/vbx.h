/*
 * vbx.h -- C ABI of libvbx_sm100a.so: the B200-native (sm_100a) kernels of the Voicebox conditional-flow-matching
 * hot path (lucidrains/voicebox-pytorch @ v0.5.0, `vp.py` = voicebox_pytorch/voicebox_pytorch.py).
 *
 * The reference has no operator registry / FFI: its seams are Python methods.  Each entry point below replaces the
 * chain of ATen ops named in its comment (reference file:line) and is what a `torch.autograd.Function` inside the
 * replaced `forward` binds through ctypes (see INTEGRATION.md for the reference-side stub).
 *
 * Conventions
 *   - plain pointers + int64 sizes only; every pointer is DEVICE memory owned by the caller (torch allocator);
 *     the library never allocates, frees or retains device memory.
 *   - `stream` is a cudaStream_t passed as void*; every launch goes on it; no device synchronisation inside.
 *   - return 0 on success, a negative VBX_E_* for argument errors detected before launch, or a positive
 *     cudaError_t from the launch.  vbx_strerror() maps both.
 *   - bf16 = __nv_bfloat16 bits (uint16_t), f32 = float, masks = uint8_t (torch.bool storage, 0/1).
 *   - "rows" are tokens; all row-major with the feature dimension contiguous unless a stride is given.
 */
#ifndef VBX_H_
#define VBX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VBX_VERSION 100

enum {
  VBX_OK = 0,
  VBX_E_NULL = -1,        /* required pointer is NULL */
  VBX_E_SHAPE = -2,       /* size is <= 0 or violates a documented divisibility rule */
  VBX_E_ALIGN = -3,       /* pointer or stride not aligned for vector / TMA access */
  VBX_E_UNSUPPORTED = -4, /* valid in the reference but outside what this kernel implements */
  VBX_E_DRIVER = -5       /* CUDA driver entry point (tensor-map encode) unavailable or failed */
};

int vbx_version(void);
const char* vbx_strerror(int rc);

/* ---------------------------------------------------------------------------------------------------------------
 * Fused residual add + (adaptive) RMSNorm                       replaces vp.py:237-247 (RMSNorm), 249-276 (AdaptiveRMSNorm)
 *                                                               and the `+ x` residual adds at vp.py:468, 471.
 *   x_out[b,r,:] = x_in[b,r,:] + branch[b,r,:]                 (branch may be NULL: x_out = x_in, x_out may be NULL)
 *   h[b,r,:]     = x_out / max(||x_out||_2, 1e-12) * sqrt(D) * gamma[b or 0,:] (+ beta[b,:])        -> bf16
 * x_in and branch are addressed as base + b*x_batch_stride + (row0+r)*D  (lets the final norm skip the register
 * tokens, vp.py:476-479); x_out/h are dense [B, rows, D].  gamma/beta: f32 [B,D] when per_batch!=0 (adaptive,
 * outputs of to_gamma/to_beta, vp.py:273) else gamma f32 [D] and beta NULL.   D % 8 == 0, D <= 2048.
 * x_out may alias x_in when x_batch_stride == rows*D and row0 == 0 (inference, in-place residual stream).
 * rstd (f32 [B*rows], may be NULL) receives 1/max(||x_out||,1e-12) for the backward. */
int vbx_adarms_fwd(const float* x_in, int64_t x_batch_stride, int64_t row0, const uint16_t* branch, const float* gamma,
                   const float* beta, int per_batch, float* x_out, uint16_t* h, float* rstd, int64_t B, int64_t rows,
                   int64_t D, void* stream);

/* Backward of the above.  Inputs: x (= x_out of the forward, same addressing as x_in), rstd, gamma, dh (bf16 [B,rows,D]),
 * dx_res (f32 [B,rows,D] gradient arriving on the residual stream, may be NULL).
 * Outputs: dx (f32, dense [B,rows,D]) = dx_res + J^T dh ; dbranch (bf16 copy of dx, may be NULL);
 *          dgamma/dbeta: f32 [B,D] (per_batch) or [D], ACCUMULATED with atomics -- caller zeroes them. */
int vbx_adarms_bwd(const float* x, int64_t x_batch_stride, int64_t row0, const float* rstd, const float* gamma, int per_batch,
                   const uint16_t* dh, const float* dx_res, float* dx, uint16_t* dbranch, float* dgamma, float* dbeta,
                   int64_t B, int64_t rows, int64_t D, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * GEGLU                                                          replaces vp.py:337-340 (chunk, F.gelu, mul)
 *   h: bf16 [T, 2*Fp], value = h[:, :Fp], gate = h[:, Fp:]  ->  out bf16 [T, Fp] = gelu_erf(gate) * value
 * Fp is the feed-forward inner width zero-padded to a multiple of 8 by the caller (exact: gelu(0)*0 = 0). */
int vbx_geglu_fwd(const uint16_t* h, uint16_t* out, int64_t T, int64_t Fp, void* stream);
/* dbias (f32 [2*Fp], may be NULL, ACCUMULATED: caller zeroes): column sums of dh, i.e. the bias gradient of the Linear that
 * produced h (vp.py:345) -- saves the separate reduction pass over dh. */
int vbx_geglu_bwd(const uint16_t* h, const uint16_t* dout, uint16_t* dh, float* dbias, int64_t T, int64_t Fp, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Conv positional embedding + residual + register-token pack      replaces vp.py:203-233 (+ caller's `+ x`, :826/:1080)
 *                                                                 and the register-token pack, vp.py:422-425.
 *   u[b,n,c]   = bias[c] + sum_j w[c,j] * (x[b,n+j-K/2,c] * m[b,n+j-K/2])       (depthwise, zero padded, K odd <= 31)
 *   y[b,R+n,c] = gelu_erf(u) * m[b,n] + x[b,n,c]                                   -> f32 residual stream [B, R+N, C]
 *   y[b,r,c]   = reg[r,c] for r < R                                               (reg may be NULL iff R == 0)
 * x bf16 [B,N,C]; w f32 [C,K]; bias f32 [C]; m uint8 [B,N] or NULL; pre (bf16 [B,N,C], may be NULL) receives u for
 * the backward.  C % 64 == 0. */
int vbx_convpos_fwd(const uint16_t* x, const float* w, const float* bias, const uint8_t* mask, const float* reg,
                    float* y, uint16_t* pre, int64_t B, int64_t N, int64_t C, int64_t K, int64_t R, void* stream);
/* dy f32 [B,R+N,C] -> dx bf16 [B,N,C] (= dy + conv^T(g)), dw f32 [C,K], dbias f32 [C], dreg f32 [R,C]
 * (dw/dbias/dreg ACCUMULATED with atomics; caller zeroes).  g = dy * gelu'(pre) * m. */
int vbx_convpos_bwd(const uint16_t* x, const uint16_t* pre, const float* w, const uint8_t* mask, const float* dy,
                    uint16_t* dx, float* dw, float* dbias, float* dreg, int64_t B, int64_t N, int64_t C, int64_t K,
                    int64_t R, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * CFM noise/interpolation + conditioning mask + concat           replaces vp.py:1408-1410 (w, flow), vp.py:1003
 *                                                                (cond = target), :1035 (cond * ~mask), :1075-1076 (cat)
 *   emb[b,n,0:D]  = bf16( (1-(1-sigma) t_b) x0 + t_b x1 )
 *   emb[b,n,D:2D] = bf16( (x1 - (1-sigma) x0) * !cond_mask[b,n] )
 * x0,x1 f32 [B,N,D]; times f32 [B]; cond_mask uint8 [B,N]; emb bf16 [B,N,2D].  D % 8 == 0. */
int vbx_cfm_embed(const float* x0, const float* x1, const float* times, const uint8_t* cond_mask, float sigma,
                  uint16_t* emb, int64_t B, int64_t N, int64_t D, void* stream);
/* General form used by VoiceBox.forward / sampling: emb[...,0:D] = bf16(x) (skipped if x NULL),
 * emb[...,D:2D] = bf16(cond * !cond_mask) (skipped if cond NULL; cond_mask NULL = keep all). */
int vbx_embed_concat(const float* x, const float* cond, const uint8_t* cond_mask, uint16_t* emb, int64_t B, int64_t N,
                     int64_t D, void* stream);

/* Masked-mean MSE against the flow target                        replaces vp.py:1099-1115
 *   num[b] += sum_n m[b,n] * mean_d (pred - target)^2            (num f32 [B], ACCUMULATED; caller zeroes)
 * target = tgt (f32 [B,N,D]) if tgt != NULL else x1 - (1-sigma) x0 recomputed on the fly.  pred bf16.
 * The caller finishes loss = mean_b(num[b] / max(den[b],1e-5)) on [B] scalars. */
int vbx_masked_mse_fwd(const uint16_t* pred, const float* tgt, const float* x0, const float* x1, float sigma,
                       const uint8_t* loss_mask, float* num, int64_t B, int64_t N, int64_t D, void* stream);
/* dpred[b,n,:] = coef[b] * m[b,n] * (pred - target)   with coef[b] = 2*gout / (D * max(den_b,1e-5) * B)  (f32 [B]) */
int vbx_masked_mse_bwd(const uint16_t* pred, const float* tgt, const float* x0, const float* x1, float sigma,
                       const uint8_t* loss_mask, const float* coef, uint16_t* dpred, int64_t B, int64_t N, int64_t D,
                       void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * One fused ODE stage combine                                     replaces torchdiffeq fixed-grid euler/midpoint
 *                                                                 stage arithmetic at the vp.py:1295 call site.
 *   dt = t[i1] - t[i0];  a = half ? 0.5*dt : dt;   y_out = y + a * f            (f bf16 = model prediction)
 *   emb[b,n,0:D] = bf16(y_out)   if emb != NULL (refreshes the x-half of the next evaluation's to_embed input)
 *   t_out[0]     = t[i0] + 0.5*dt if t_out != NULL (time of the midpoint evaluation, a device scalar)
 * t is the DEVICE linspace grid: no host value is baked in, so a CUDA graph of one solver step can be replayed. */
int vbx_ode_axpy(const float* y, const uint16_t* f, const float* t, int64_t i0, int64_t i1, int half, float* y_out,
                 uint16_t* emb, float* t_out, int64_t B, int64_t N, int64_t D, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * bf16 GEMM on tcgen05 tensor cores with fused epilogues          replaces F.linear under autocast (vp.py:320, 333, 348, 1078, 1092)
 *                                                                 and, for vbx_ff1_geglu, vp.py:337-346: Linear(D, 2F) + GEGLU.
 * a: bf16 [M, K] activations; w: bf16 [N, K] (nn.Linear layout, K contiguous); bias: bf16 [N] or NULL; c: bf16 [M, N].
 * fp32 accumulation in tensor memory, one rounding to bf16 on output.  K % 8 == 0, N % 8 == 0, N >= 32; any M. */
int vbx_gemm_bf16(const uint16_t* a, const uint16_t* w, const uint16_t* bias, uint16_t* c, int64_t M, int64_t N, int64_t K,
                  void* stream);
/* x: bf16 [M, K]; w1: bf16 [2Fp, K] = (value rows | gate rows), b1: bf16 [2Fp] (the zero-padded operand layout, Fp % 32 == 0);
 *   h = x w1^T + b1 (rounded to bf16, as the reference's autocast Linear output), g = gelu_erf(h[:, Fp:]) * h[:, :Fp]
 * h: bf16 [M, 2Fp] or NULL (inference: the intermediate is never written); g: bf16 [M, Fp]. */
int vbx_ff1_geglu(const uint16_t* x, const uint16_t* w1, const uint16_t* b1, uint16_t* h, uint16_t* g, int64_t M, int64_t Fp,
                  int64_t K, void* stream);

/* FF2 data gradient + GEGLU backward in one launch (vp.py:337-348, backward):
 *   dg = dy w2 (never written) ; dh[:, :Fp] = dg * gelu_erf(gate) ; dh[:, Fp:] = dg * value * gelu_erf'(gate) ; db1 += colsum(dh)
 * dy: bf16 [M, K] (gradient of the feed-forward output); w2t: bf16 [Fp, K] = the second Linear's weight TRANSPOSED and
 * zero-padded (rows >= F are zero); h: bf16 [M, 2Fp] saved by vbx_ff1_geglu; dh: bf16 [M, 2Fp]; db1: f32 [2Fp], accumulated
 * with atomics (caller zeroes).  Fp % 64 == 0, K % 8 == 0. */
int vbx_ff2_dgrad_geglu_bwd(const uint16_t* dy, const uint16_t* w2t, const uint16_t* h, uint16_t* dh, float* db1, int64_t M,
                            int64_t Fp, int64_t K, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Operand packing                                               replaces the per-use fp32 -> bf16 casts of every Linear weight
 *                                                               and bias under autocast (trainer.py:267; vp.py:259-260, 320,
 *                                                               333, 345, 348) -- ~14 cast kernels per layer per forward --
 *                                                               and the GEGLU zero-padding copies, with ONE launch.
 * segs: DEVICE table, n_seg entries of 6 x int64: { src (const float*), dst (uint16_t*), rows, cols, src_pitch, dst_pitch }
 * (pitches in elements); row_start: DEVICE int64 [n_seg+1], exclusive prefix sum of `rows`; total_rows = row_start[n_seg].
 * Every segment copies a [rows x cols] fp32 matrix into a bf16 matrix of a possibly larger pitch; bytes outside the copied
 * columns / rows are never written (operand buffers are allocated zero-filled once, which is what makes the padding exact). */
int vbx_pack_bf16(const void* segs, const int64_t* row_start, int64_t n_seg, int64_t total_rows, void* stream);
/* dst (f32 [rows, cols], row pitch dst_pitch) += src (bf16, row pitch src_pitch): a bf16 weight gradient accumulated into the fp32
 * master gradient in one pass (replaces autograd's cast + AccumulateGrad add for the weights behind packed operands). */
int vbx_accum_bf16_2d(float* dst, int64_t dst_pitch, const uint16_t* src, int64_t src_pitch, int64_t rows, int64_t cols, void* stream);
/* Table form of the same (segment layout of vbx_pack_bf16 with src = bf16, dst = f32): many blocks, one launch. */
int vbx_accum_bf16_table(const void* segs, const int64_t* row_start, int64_t n_seg, int64_t total_rows, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Fused gradient clip + Adam step over FLAT buffers              replaces accelerator.clip_grad_norm_ + optim.step()
 *                                                                 (trainer.py:274-278; optimizer.py:32-35: torch Adam,
 *                                                                 betas (0.9, 0.99), eps 1e-8; AdamW when wd > 0)
 * p, g, m, v: f32 [n], 16-byte aligned: every parameter / gradient / first / second moment of the model, each laid out
 * contiguously in the same order (voicebox-pytorch_b200/dist.py keeps the gradients that way; optim.py the rest).
 *   g' = g / grad_scale[0]         grad_scale: DEVICE scalar or NULL -- the clip: max(1, (||g||+1e-6)/max_norm)
 *   g' += weight_decay * p         (decoupled == 0, torch.optim.Adam)    |   p *= 1 - lr*weight_decay  (decoupled != 0, AdamW)
 *   m = m + (1-beta1)(g' - m);  v = beta2 v + (1-beta2) g'^2
 *   p -= lr/(1-beta1^step) * m / (sqrt(v)/sqrt(1-beta2^step) + eps)        step is 1-based
 * found_inf: DEVICE scalar or NULL; when non-zero the launch changes nothing (same contract as torch's fused optimizers).
 * p_bf16 (bf16 [n], 8-byte aligned, may be NULL) receives the updated parameters rounded to bf16: the tensor-core operand
 * copies of the next forward, for free.  One pass: 28 B per element (+2). */
int vbx_adam_step(float* p, const float* g, float* m, float* v, uint16_t* p_bf16, int64_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int decoupled, int64_t step, const float* grad_scale, const float* found_inf,
                  void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * qk-RMSNorm + rotary + head split (attention prologue)          replaces vp.py:323-328 (rearrange, q_norm/k_norm,
 *                                                                 apply_rotary_pos_emb) incl. vp.py:193-199, 280-287
 * qkv bf16 [B, N, 3*H*64] (q | k | v blocks, each h-major d-minor, vp.py:320-321);
 * cosv/sinv f32 [N,32] = cos/sin(pos (x) inv_freq) (half-split rotary: pairs (d, d+32));
 * gq/gk f32 [H,64] or NULL (no qk-norm).  Writes qh, kh bf16 [B,N,H,64] (token-major:
 * the rope kernels move whole-token contiguous runs, the attention kernels address tiles through tensor-map strides):
 *   q^ = rot( q/max(||q||,1e-12) * 8 * gq )     (norm skipped when gq NULL) */
int vbx_qkrope_fwd(const uint16_t* qkv, const float* cosv, const float* sinv, const float* gq, const float* gk,
                   uint16_t* qh, uint16_t* kh, int64_t B, int64_t N, int64_t H, void* stream);
/* dqh f32 [B,N,H,64] (accumulated by vbx_attn_bwd's TMA reduce-adds), dkh bf16 [B,N,H,64] -> writes the q and k blocks of
 * dqkv (bf16 [B,N,3*H*64]; the v block is written by vbx_attn_bwd) and accumulates dgq, dgk f32 [H,64]. */
int vbx_qkrope_bwd(const uint16_t* qkv, const float* cosv, const float* sinv, const float* gq, const float* gk,
                   const float* dqh, const uint16_t* dkh, uint16_t* dqkv, float* dgq, float* dgk, int64_t B, int64_t N,
                   int64_t H, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Attention core (tcgen05 + TMEM + TMA)                           replaces attend.py:100-137 (math path) /
 *                                                                 attend.py:71-98 (SDPA path), head merge vp.py:332
 *   O = softmax(scale * Q K^T, masked keys -> -FLT_MAX) V          dim_head = 64
 * q,k bf16 [B,N,H,64] (as written by vbx_qkrope_fwd); v bf16 addressed v + b*v_bs + n*v_ns + h*64 (element strides; lets V be read in place from
 * the qkv GEMM output); key_mask uint8 [B,N] or NULL; o bf16 [B,N,H*64]; lse f32 [B,H,N] (log2 domain, for bwd). */
int vbx_attn_fwd(const uint16_t* q, const uint16_t* k, const uint16_t* v, int64_t v_bs, int64_t v_ns,
                 const uint8_t* key_mask, float scale, uint16_t* o, float* lse, int64_t B, int64_t H, int64_t N,
                 void* stream);
/* delta f32 [B,H,N] workspace; dq f32 [B,N,H,64] ACCUMULATED (caller zeroes); dk bf16 [B,N,H,64];
 * dv bf16 addressed like v (dv_bs/dv_ns). */
int vbx_attn_bwd(const uint16_t* q, const uint16_t* k, const uint16_t* v, int64_t v_bs, int64_t v_ns,
                 const uint8_t* key_mask, float scale, const uint16_t* o, const uint16_t* dout, const float* lse,
                 float* delta, float* dq, uint16_t* dk, uint16_t* dv, int64_t dv_bs, int64_t dv_ns, int64_t B,
                 int64_t H, int64_t N, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * tcgen05 / TMA self-test: c (f32 [128,128]) = A B^T with K = 128 through the same shared-memory / instruction
 * descriptors the attention kernels use.  a, b: bf16 [128,128].  variant bit0: B is MN-major (b holds [K][N]),
 * bit1: A is MN-major (a holds [K][M]), bit2: stage operands with TMA instead of thread stores, bit3: A operand placed in
 * TMEM with tcgen05.st and consumed by a TS-mode MMA (not combinable with bit1). */
int vbx_umma_selftest(const uint16_t* a, const uint16_t* b, float* c, int variant, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VBX_H_ */
