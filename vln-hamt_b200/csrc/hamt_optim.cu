// Fused optimizer step on the flat parameter arena.
//
// Replaces the reference's per-parameter python loop (pretrain_src/optim/adamw.py:53-110, ~400 parameters x ~8 eager kernels)
// and torch.nn.utils.clip_grad_norm_ (main_r2r.py:271-274) by three launches over ONE contiguous buffer:
//   1. adamw_sqnorm_kernel   : per-block partial sums of squares of the gradients of the ACTIVE segments
//   2. adamw_prepare_kernel  : (1 block) deterministic reduction of the partials -> global norm, clip coefficient;
//                              per-segment step counters += active; per-segment bias-corrected step size (double math, as
//                              the reference computes it in python floats)
//   3. adamw_update_kernel   : m, v, p update in fp32 exactly in the reference's operation order (weight decay applied AFTER
//                              the Adam update with lr * wd, eps added OUTSIDE the sqrt), + bf16 shadow of the new weights
//                              (the GEMM operand, saves the separate cast pass) + optional gradient zeroing.
// A "segment" is one parameter tensor: [seg_off[s], seg_off[s] + seg_len[s]) in elements, 64-element aligned (arena.py); a
// parameter whose .grad is None this step (task did not touch it) is inactive: no state update, step counter unchanged
// (adamw.py:64-66).  HBM-bound: 16 B read + 18 B written per active element.
#include <stdio.h>
#include "hamt_common.cuh"
#include "hamt_kernels.h"

namespace hamt {

static constexpr int kChunk = 64;          // elements per chunk_seg entry (arena alignment)
static constexpr int kNormBlocks = 592;    // 148 SMs x 4

__global__ void __launch_bounds__(256) adamw_sqnorm_kernel(const float* __restrict__ g, const int* __restrict__ chunk_seg,
                                                           const unsigned char* __restrict__ active, const long long* __restrict__ seg_end,
                                                           long long n_vec, float* __restrict__ partials) {
  pdl_grid_sync();
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_vec; i += (long long)gridDim.x * blockDim.x) {
    const int s = chunk_seg[i >> 4];
    if (s < 0 || !active[s]) continue;
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    const long long left = seg_end[s] - 4 * i;          // elements of this vector that belong to the parameter (alignment tail excluded)
    if (left >= 4) acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    else {
      if (left > 0) acc += v.x * v.x;
      if (left > 1) acc += v.y * v.y;
      if (left > 2) acc += v.z * v.z;
    }
  }
  __shared__ float sh[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += sh[k];
    partials[blockIdx.x] = t;
  }
}

// scal[0] = global grad norm, scal[1] = clip coefficient (<= 1)
__global__ void __launch_bounds__(256) adamw_prepare_kernel(const float* __restrict__ partials, int n_partials, float max_norm, float* __restrict__ scal,
                                                            int* __restrict__ seg_step, const unsigned char* __restrict__ active,
                                                            float* __restrict__ seg_step_size, int nseg, const float* __restrict__ lr_ptr,
                                                            double beta1, double beta2, int correct_bias) {
  pdl_grid_sync();
  __shared__ double sh[256];
  double acc = 0.0;
  if (partials != nullptr)
    for (int i = threadIdx.x; i < n_partials; i += blockDim.x) acc += (double)partials[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = (float)sqrt(sh[0]);
    float coef = 1.0f;
    if (max_norm > 0.f) coef = fminf(max_norm / (norm + 1e-6f), 1.0f);     // torch.nn.utils.clip_grad_norm_: clamp(max_norm / (norm + 1e-6), max=1)
    scal[0] = norm;
    scal[1] = coef;
  }
  const double lr = (double)lr_ptr[0];
  for (int s = threadIdx.x; s < nseg; s += blockDim.x) {
    if (!active[s]) continue;
    const int t = seg_step[s] + 1;
    seg_step[s] = t;
    double step_size = lr;
    if (correct_bias) step_size = lr * sqrt(1.0 - pow(beta2, (double)t)) / (1.0 - pow(beta1, (double)t));   // adamw.py:92-96
    seg_step_size[s] = (float)step_size;
  }
}

struct AdamWParams {
  float* p; float* g; float* m; float* v; __nv_bfloat16* shadow;
  const int* chunk_seg; const long long* seg_end; const unsigned char* active; const float* seg_wd; const float* seg_step_size; const float* scal; const float* lr_ptr;
  long long n_vec;
  float beta1, beta2, omb1, omb2, eps;
  int zero_grad;
};

__global__ void __launch_bounds__(256) adamw_update_kernel(const AdamWParams a) {
  pdl_grid_sync();
  const float coef = a.scal[1];
  const float lr = a.lr_ptr[0];
  const float omb1 = a.omb1, omb2 = a.omb2;     // float(1 - beta) with the subtraction in double, as python passes alpha
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.n_vec; i += (long long)gridDim.x * blockDim.x) {
    const int s = a.chunk_seg[i >> 4];
    if (s < 0 || !a.active[s]) continue;
    const float step_size = a.seg_step_size[s];
    const float decay = lr * a.seg_wd[s];
    float4 g4 = reinterpret_cast<const float4*>(a.g)[i];
    const long long left = a.seg_end[s] - 4 * i;
    if (left < 4) {                                       // alignment tail of the parameter: padding carries no gradient
      if (left < 1) g4.x = 0.f;
      if (left < 2) g4.y = 0.f;
      if (left < 3) g4.z = 0.f;
      g4.w = 0.f;
    }
    float4 m4 = reinterpret_cast<const float4*>(a.m)[i];
    float4 v4 = reinterpret_cast<const float4*>(a.v)[i];
    float4 p4 = reinterpret_cast<const float4*>(a.p)[i];
    float gg[4] = {g4.x, g4.y, g4.z, g4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = gg[k] * coef;                                   // clip_grad_norm_ scales the gradient in place
      mm[k] = __fadd_rn(__fmul_rn(mm[k], a.beta1), __fmul_rn(gk, omb1));             // exp_avg.mul_(b1).add_(grad, alpha=1-b1)
      vv[k] = __fadd_rn(__fmul_rn(vv[k], a.beta2), __fmul_rn(__fmul_rn(gk, gk), omb2));   // exp_avg_sq.mul_(b2).addcmul_(grad, grad, value=1-b2)
      const float denom = __fadd_rn(__fsqrt_rn(vv[k]), a.eps);         // sqrt().add_(eps)
      pp[k] = __fadd_rn(pp[k], __fmul_rn(-step_size, __fdiv_rn(mm[k], denom)));     // addcdiv_(exp_avg, denom, value=-step_size)
      if (decay > 0.f) pp[k] = __fadd_rn(pp[k], __fmul_rn(pp[k], -decay));          // p.add_(p, alpha=-lr*wd)   (adamw.py:107-108)
    }
    reinterpret_cast<float4*>(a.m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(a.v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    reinterpret_cast<float4*>(a.p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    if (a.shadow != nullptr) reinterpret_cast<uint2*>(a.shadow)[i] = make_uint2(pack_bf16(pp[0], pp[1]), pack_bf16(pp[2], pp[3]));
    if (a.zero_grad) reinterpret_cast<float4*>(a.g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

int adamw_workspace_floats(void) { return kNormBlocks + 2; }

int adamw_step(const AdamWArgs& a, cudaStream_t st) {
  HAMT_REQUIRE(a.total > 0 && a.total % kChunk == 0, "adamw: the flat buffer length must be a positive multiple of 64 elements");
  HAMT_REQUIRE(a.nseg > 0 && a.chunk_seg && a.seg_end && a.seg_active && a.seg_wd && a.seg_step && a.seg_step_size, "adamw: segment tables missing");
  HAMT_REQUIRE(a.param && a.grad && a.exp_avg && a.exp_avg_sq && a.lr && a.workspace, "adamw: null buffer");
  HAMT_REQUIRE((((uintptr_t)a.param | (uintptr_t)a.grad | (uintptr_t)a.exp_avg | (uintptr_t)a.exp_avg_sq | (uintptr_t)a.shadow) & 15) == 0,
               "adamw: buffers must be 16-byte aligned");
  HAMT_REQUIRE(a.beta1 >= 0.0 && a.beta1 < 1.0 && a.beta2 >= 0.0 && a.beta2 < 1.0 && a.eps >= 0.0, "adamw: invalid beta / eps");   // adamw.py:42-49
  const long long n_vec = a.total / 4;
  float* partials = a.workspace + 2;
  const bool clip = a.max_grad_norm > 0.f || a.want_norm;
  if (clip) {
    launch_pdl(adamw_sqnorm_kernel, kNormBlocks, 256, 0, st, (const float*)a.grad, a.chunk_seg, a.seg_active, a.seg_end, n_vec, partials);
    if (int rc = check_launch("adamw_sqnorm_kernel")) return rc;
  }
  launch_pdl(adamw_prepare_kernel, 1, 256, 0, st, clip ? (const float*)partials : (const float*)nullptr, kNormBlocks, a.max_grad_norm, a.workspace,
             a.seg_step, a.seg_active, a.seg_step_size, a.nseg, a.lr, a.beta1, a.beta2, a.correct_bias);
  if (int rc = check_launch("adamw_prepare_kernel")) return rc;
  AdamWParams p{a.param, a.grad, a.exp_avg, a.exp_avg_sq, (__nv_bfloat16*)a.shadow, a.chunk_seg, a.seg_end, a.seg_active, a.seg_wd, a.seg_step_size,
                a.workspace, a.lr, n_vec, (float)a.beta1, (float)a.beta2, (float)(1.0 - a.beta1), (float)(1.0 - a.beta2), (float)a.eps, a.zero_grad};
  long long blocks = (n_vec + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_pdl(adamw_update_kernel, (int)blocks, 256, 0, st, p);
  return check_launch("adamw_update_kernel");
}

}  // namespace hamt
