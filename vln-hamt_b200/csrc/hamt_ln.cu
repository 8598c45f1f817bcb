// Row-wise bandwidth kernels: fused dropout + residual + LayerNorm forward/backward, plus the small
// streaming helpers (cast, column-sum, mean-pool, add, row-broadcast multiply).
//
// Replaces the eager sequences  dense -> Dropout -> (+residual) -> LayerNorm  of BertSelfOutput /
// BertOutput (pretrain_src/model/vilmodel.py:139-143, :181-185) and their autograd backward; eps is
// added in fp32 (layer_norm_eps = 1e-12 is below bf16 resolution, SURVEY.md 8a gotcha 5).
// One warp per row, 128-bit loads: lane l owns columns {c*256 + l*8 + j}.
#include <stdio.h>
#include "hamt_common.cuh"
#include "hamt_kernels.h"

namespace hamt {

template <int NCH>
struct Row {
  float v[NCH * 8];
};

template <int NCH>
__device__ __forceinline__ void load_row_bf16(const __nv_bfloat16* p, int lane, float (&v)[NCH * 8]) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint4 w = *reinterpret_cast<const uint4*>(p + c * 256 + lane * 8);
    float2 f0 = unpack_bf16(w.x), f1 = unpack_bf16(w.y), f2 = unpack_bf16(w.z), f3 = unpack_bf16(w.w);
    v[c * 8 + 0] = f0.x; v[c * 8 + 1] = f0.y; v[c * 8 + 2] = f1.x; v[c * 8 + 3] = f1.y;
    v[c * 8 + 4] = f2.x; v[c * 8 + 5] = f2.y; v[c * 8 + 6] = f3.x; v[c * 8 + 7] = f3.y;
  }
}
template <int NCH>
__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* p, int lane, const float (&v)[NCH * 8]) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint4 w = make_uint4(pack_bf16(v[c * 8 + 0], v[c * 8 + 1]), pack_bf16(v[c * 8 + 2], v[c * 8 + 3]),
                         pack_bf16(v[c * 8 + 4], v[c * 8 + 5]), pack_bf16(v[c * 8 + 6], v[c * 8 + 7]));
    *reinterpret_cast<uint4*>(p + c * 256 + lane * 8) = w;
  }
}
template <int NCH>
__device__ __forceinline__ void load_vec_f32(const float* p, int lane, float (&v)[NCH * 8]) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    float4 a = *reinterpret_cast<const float4*>(p + c * 256 + lane * 8);
    float4 b = *reinterpret_cast<const float4*>(p + c * 256 + lane * 8 + 4);
    v[c * 8 + 0] = a.x; v[c * 8 + 1] = a.y; v[c * 8 + 2] = a.z; v[c * 8 + 3] = a.w;
    v[c * 8 + 4] = b.x; v[c * 8 + 5] = b.y; v[c * 8 + 6] = b.z; v[c * 8 + 7] = b.w;
  }
}

template <int NCH>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const __nv_bfloat16* x /* may alias z_out */, const __nv_bfloat16* __restrict__ res,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     __nv_bfloat16* __restrict__ y, __nv_bfloat16* z_out, float* __restrict__ mean_out,
                                                     float* __restrict__ rstd_out, int M, float eps, DropCfg dc) {
  pdl_grid_sync();
  constexpr int H = NCH * 256;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const DropState ds = drop_init(dc);
  float g[NCH * 8], b[NCH * 8];
  load_vec_f32<NCH>(gamma, lane, g);
  load_vec_f32<NCH>(beta, lane, b);
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += gridDim.x * wpb) {
    float z[NCH * 8];
    load_row_bf16<NCH>(x + (long long)row * H, lane, z);
    if (ds.on) {
#pragma unroll
      for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) z[c * 8 + j] *= drop_mult(ds, (unsigned long long)row * H + c * 256 + lane * 8 + j);
    }
    if (res != nullptr) {
      float r[NCH * 8];
      load_row_bf16<NCH>(res + (long long)row * H, lane, r);
#pragma unroll
      for (int i = 0; i < NCH * 8; ++i) z[i] += r[i];
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCH * 8; ++i) s += z[i];
    const float mean = warp_sum(s) * (1.0f / H);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NCH * 8; ++i) { const float d = z[i] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / H) + eps);
    if (z_out != nullptr) store_row_bf16<NCH>(z_out + (long long)row * H, lane, z);
    float o[NCH * 8];
#pragma unroll
    for (int i = 0; i < NCH * 8; ++i) o[i] = (z[i] - mean) * rstd * g[i] + b[i];
    store_row_bf16<NCH>(y + (long long)row * H, lane, o);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
  }
}

template <int NCH>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
                                                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                     const float* __restrict__ gamma, const __nv_bfloat16* __restrict__ dres_in,
                                                     __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dres,
                                                     float* dgamma, float* dbeta, float* dbias, int M, DropCfg dc) {
  pdl_grid_sync();
  constexpr int H = NCH * 256;
  __shared__ float sacc[3][H];
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const DropState ds = drop_init(dc);
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) (&sacc[0][0])[i] = 0.f;
  __syncthreads();
  float g[NCH * 8];
  load_vec_f32<NCH>(gamma, lane, g);
  float acc_g[NCH * 8], acc_b[NCH * 8], acc_x[NCH * 8];
#pragma unroll
  for (int i = 0; i < NCH * 8; ++i) acc_g[i] = acc_b[i] = acc_x[i] = 0.f;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += gridDim.x * wpb) {
    float d[NCH * 8], zz[NCH * 8];
    load_row_bf16<NCH>(dy + (long long)row * H, lane, d);
    load_row_bf16<NCH>(z + (long long)row * H, lane, zz);
    const float mean = mean_in[row], rstd = rstd_in[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NCH * 8; ++i) {
      const float xh = (zz[i] - mean) * rstd;
      const float gd = d[i] * g[i];
      acc_g[i] += d[i] * xh;
      acc_b[i] += d[i];
      s1 += gd;
      s2 += gd * xh;
      zz[i] = xh;
      d[i] = gd;
    }
    s1 = warp_sum(s1) * (1.0f / H);
    s2 = warp_sum(s2) * (1.0f / H);
    float dz[NCH * 8];
#pragma unroll
    for (int i = 0; i < NCH * 8; ++i) dz[i] = rstd * (d[i] - s1 - zz[i] * s2);
    if (dres != nullptr) {
      float o[NCH * 8];
      if (dres_in != nullptr) {
        load_row_bf16<NCH>(dres_in + (long long)row * H, lane, o);
#pragma unroll
        for (int i = 0; i < NCH * 8; ++i) o[i] += dz[i];
      } else {
#pragma unroll
        for (int i = 0; i < NCH * 8; ++i) o[i] = dz[i];
      }
      store_row_bf16<NCH>(dres + (long long)row * H, lane, o);
    }
    if (dx != nullptr) {
      if (ds.on) {
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
          for (int j = 0; j < 8; ++j) dz[c * 8 + j] *= drop_mult(ds, (unsigned long long)row * H + c * 256 + lane * 8 + j);
      }
      // bias gradient is the column sum of the *rounded* dx the wgrad GEMM will also see
#pragma unroll
      for (int i = 0; i < NCH * 8; ++i) acc_x[i] += dz[i];
      store_row_bf16<NCH>(dx + (long long)row * H, lane, dz);
    }
  }
  // cross-warp reduction through shared memory, then one atomic per column per CTA
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = c * 256 + lane * 8 + j;
      atomicAdd(&sacc[0][col], acc_g[c * 8 + j]);
      atomicAdd(&sacc[1][col], acc_b[c * 8 + j]);
      atomicAdd(&sacc[2][col], acc_x[c * 8 + j]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + i, sacc[0][i]);
    if (dbeta) atomicAdd(dbeta + i, sacc[1][i]);
    if (dbias) atomicAdd(dbias + i, sacc[2][i]);
  }
}

// EXPERIMENTAL variant of ln_bwd_kernel (opt-in through hamt_ln_set_variant(1); default off, unmeasured).  Same arithmetic, two
// scheduling changes aimed at the small-M launches (M = 2.5 k .. 8.5 k rows, 26 us at 1.2 TB/s -- 40 of the 44 launches per step):
//   * the residual-path gradient row (dres_in) is requested together with dy and z instead of after the two warp reductions
//     (one exposed HBM latency per row less);
//   * the cross-warp reduction at the end goes through warp-private shared slabs [warp][3][NCH*8][32] (value index major, lane minor:
//     32 consecutive banks per store) that the CTA folds afterwards; ln_bwd_kernel uses 72 shared fp32 atomics per lane on a row-major
//     [3][H] array -- 8-way bank conflicts, and every shared fp32 atomic is a compare-and-swap spin loop on sm_100 (ATOMS.CAST.SPIN).
// OCC2: compile for two resident CTAs per SM (128 registers, some spills) instead of one (246 registers; the default kernel also needs
// 238 and therefore runs 8 warps per SM) -- which of the two wins is a measurement for round 2.
template <int NCH, bool OCC2>
__global__ void __launch_bounds__(256, OCC2 ? 2 : 1) ln_bwd_kernel_v2(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
                                                        const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                        const float* __restrict__ gamma, const __nv_bfloat16* __restrict__ dres_in,
                                                        __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dres,
                                                        float* dgamma, float* dbeta, float* dbias, int M, DropCfg dc) {
  pdl_grid_sync();
  constexpr int H = NCH * 256;
  constexpr int NV = NCH * 8;                 // values per lane
  extern __shared__ float slab_all[];                 // [warps][3][NV][32]: warp-private, written once at the end (no atomics: fp32
                                                      // shared-memory atomics are compare-and-swap loops on sm_100)
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const DropState ds = drop_init(dc);
  float g[NV];
  load_vec_f32<NCH>(gamma, lane, g);
  float acc_g[NV], acc_b[NV], acc_x[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc_g[i] = acc_b[i] = acc_x[i] = 0.f;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += gridDim.x * wpb) {
    float d[NV], zz[NV], o[NV];
    load_row_bf16<NCH>(dy + (long long)row * H, lane, d);
    load_row_bf16<NCH>(z + (long long)row * H, lane, zz);
    if (dres != nullptr && dres_in != nullptr) load_row_bf16<NCH>(dres_in + (long long)row * H, lane, o);
    else {
#pragma unroll
      for (int i = 0; i < NV; ++i) o[i] = 0.f;
    }
    const float mean = mean_in[row], rstd = rstd_in[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float xh = (zz[i] - mean) * rstd;
      const float gd = d[i] * g[i];
      acc_g[i] += d[i] * xh;
      acc_b[i] += d[i];
      s1 += gd;
      s2 += gd * xh;
      zz[i] = xh;
      d[i] = gd;
    }
    s1 = warp_sum(s1) * (1.0f / H);
    s2 = warp_sum(s2) * (1.0f / H);
    float dz[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) dz[i] = rstd * (d[i] - s1 - zz[i] * s2);
    if (dres != nullptr) {
#pragma unroll
      for (int i = 0; i < NV; ++i) o[i] += dz[i];
      store_row_bf16<NCH>(dres + (long long)row * H, lane, o);
    }
    if (dx != nullptr) {
      if (ds.on) {
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
          for (int j = 0; j < 8; ++j) dz[c * 8 + j] *= drop_mult(ds, (unsigned long long)row * H + c * 256 + lane * 8 + j);
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) acc_x[i] += dz[i];
      store_row_bf16<NCH>(dx + (long long)row * H, lane, dz);
    }
  }
  float* slab = slab_all + (threadIdx.x >> 5) * (3 * H) + lane;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    slab[(0 * NV + i) * 32] = acc_g[i];
    slab[(1 * NV + i) * 32] = acc_b[i];
    slab[(2 * NV + i) * 32] = acc_x[i];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 3 * H; idx += blockDim.x) {       // fold the warp-private slabs, one global red per column and CTA
    float v = slab_all[idx];
    for (int w = 1; w < wpb; ++w) v += slab_all[w * 3 * H + idx];
    const int a = idx / H, r = idx % H;
    const int i = r >> 5, l = r & 31;                         // value i of lane l  ->  column (i / 8) * 256 + l * 8 + (i % 8)
    const int col = (i >> 3) * 256 + l * 8 + (i & 7);
    float* dst = a == 0 ? dgamma : (a == 1 ? dbeta : dbias);
    if (dst) atomicAdd(dst + col, v);
  }
}

// EXPERIMENTAL variant 3: as v2, but the three column-sum accumulators are not kept in registers across rows (72 registers per lane,
// the reason the kernels above need ~240 registers and run one CTA = 8 warps per SM).  Every warp owns a private [3][NCH*8][32] slab
// in dynamic shared memory (lane-minor: 32 consecutive banks per access, no other warp touches it) and adds its row's contributions
// with plain load / add / store -- fp32 shared-memory atomics compile to a compare-and-swap spin loop on sm_100 (ATOMS.CAST.SPIN), so
// they are avoided.  The slabs are folded by the CTA at the end.  Aim: <= 128 registers without spills -> 2 CTAs per SM.
template <int NCH>
__global__ void __launch_bounds__(256, 2) ln_bwd_kernel_v3(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
                                                           const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                           const float* __restrict__ gamma, const __nv_bfloat16* __restrict__ dres_in,
                                                           __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dres,
                                                           float* dgamma, float* dbeta, float* dbias, int M, DropCfg dc) {
  pdl_grid_sync();
  constexpr int H = NCH * 256;
  constexpr int NV = NCH * 8;
  extern __shared__ float slab_all[];                 // [warps][3][NV][32]
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const DropState ds = drop_init(dc);
  for (int i = threadIdx.x; i < wpb * 3 * H; i += blockDim.x) slab_all[i] = 0.f;
  __syncthreads();
  float* slab = slab_all + (threadIdx.x >> 5) * (3 * H) + lane;      // element (a, i) of this lane: slab[(a * NV + i) * 32]
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += gridDim.x * wpb) {
    float d[NV], zz[NV], o[NV];
    load_row_bf16<NCH>(dy + (long long)row * H, lane, d);
    load_row_bf16<NCH>(z + (long long)row * H, lane, zz);
    const bool has_in = dres != nullptr && dres_in != nullptr;
    if (has_in) load_row_bf16<NCH>(dres_in + (long long)row * H, lane, o);
    const float mean = mean_in[row], rstd = rstd_in[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const float4 ga = *reinterpret_cast<const float4*>(gamma + c * 256 + lane * 8);
      const float4 gb = *reinterpret_cast<const float4*>(gamma + c * 256 + lane * 8 + 4);
      const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = c * 8 + j;
        const float xh = (zz[i] - mean) * rstd;
        const float gd = d[i] * gg[j];
        slab[(0 * NV + i) * 32] += d[i] * xh;
        slab[(1 * NV + i) * 32] += d[i];
        s1 += gd;
        s2 += gd * xh;
        zz[i] = xh;
        d[i] = gd;
      }
    }
    s1 = warp_sum(s1) * (1.0f / H);
    s2 = warp_sum(s2) * (1.0f / H);
#pragma unroll
    for (int i = 0; i < NV; ++i) d[i] = rstd * (d[i] - s1 - zz[i] * s2);      // d <- dz
    if (dres != nullptr) {
      if (has_in) {
#pragma unroll
        for (int i = 0; i < NV; ++i) o[i] += d[i];
        store_row_bf16<NCH>(dres + (long long)row * H, lane, o);
      } else {
        store_row_bf16<NCH>(dres + (long long)row * H, lane, d);
      }
    }
    if (dx != nullptr) {
      if (ds.on) {
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
          for (int j = 0; j < 8; ++j) d[c * 8 + j] *= drop_mult(ds, (unsigned long long)row * H + c * 256 + lane * 8 + j);
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) slab[(2 * NV + i) * 32] += d[i];
      store_row_bf16<NCH>(dx + (long long)row * H, lane, d);
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 3 * H; idx += blockDim.x) {       // fold the warp-private slabs, one global red per column and CTA
    float v = slab_all[idx];
    for (int w = 1; w < wpb; ++w) v += slab_all[w * 3 * H + idx];
    const int a = idx / H, r = idx % H;
    const int i = r >> 5, l = r & 31;
    const int col = (i >> 3) * 256 + l * 8 + (i & 7);
    float* dst = a == 0 ? dgamma : (a == 1 ? dbeta : dbias);
    if (dst) atomicAdd(dst + col, v);
  }
}

static int g_ln_variant = 0;       // hamt_ln_set_variant: 0 = ln_bwd_kernel (default), 1 / 2 = ln_bwd_kernel_v2 with 1 / 2 CTAs per SM, 3 = ln_bwd_kernel_v3 (all experimental)
void ln_set_variant(int v) { g_ln_variant = v; }

static int grid_for_rows(int M, int wpb, int max_ctas) {
  int g = (M + wpb - 1) / wpb;
  return g < max_ctas ? (g < 1 ? 1 : g) : max_ctas;
}

int ln_fwd(const void* x, const void* res, const float* gamma, const float* beta, void* y, void* z_out, float* mean, float* rstd, int M,
           int H, float eps, DropArgs drop, cudaStream_t st) {
  HAMT_REQUIRE(H == 512 || H == 768 || H == 1024, "ln_fwd: hidden size must be 512/768/1024");
  if (M <= 0) return 0;
  DropCfg dc{drop.seed_ptr, drop.site, drop.p};
  const int grid = grid_for_rows(M, 8, 148 * 8);
  auto X = (const __nv_bfloat16*)x; auto R = (const __nv_bfloat16*)res; auto Y = (__nv_bfloat16*)y; auto Z = (__nv_bfloat16*)z_out;
  if (H == 768) launch_pdl(ln_fwd_kernel<3>, grid, 256, 0, st, X, R, gamma, beta, Y, Z, mean, rstd, M, eps, dc);
  else if (H == 512) launch_pdl(ln_fwd_kernel<2>, grid, 256, 0, st, X, R, gamma, beta, Y, Z, mean, rstd, M, eps, dc);
  else launch_pdl(ln_fwd_kernel<4>, grid, 256, 0, st, X, R, gamma, beta, Y, Z, mean, rstd, M, eps, dc);
  return check_launch("ln_fwd_kernel");
}

int ln_bwd(const void* dy, const void* z, const float* mean, const float* rstd, const float* gamma, const void* dres_in, void* dx, void* dres,
           float* dgamma, float* dbeta, float* dbias, int M, int H, DropArgs drop, cudaStream_t st) {
  HAMT_REQUIRE(H == 512 || H == 768 || H == 1024, "ln_bwd: hidden size must be 512/768/1024");
  if (M <= 0) return 0;
  DropCfg dc{drop.seed_ptr, drop.site, drop.p};
  const int grid = grid_for_rows(M, 8 * 4, 148 * 2);
  auto DY = (const __nv_bfloat16*)dy; auto Z = (const __nv_bfloat16*)z; auto DRI = (const __nv_bfloat16*)dres_in;
  auto DX = (__nv_bfloat16*)dx; auto DR = (__nv_bfloat16*)dres;
  if (g_ln_variant == 3) {
    const size_t smem = (size_t)8 * 3 * H * sizeof(float);           // 8 warps x [3][H] fp32: 72 KB at H = 768
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(ln_bwd_kernel_v3<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 768 * 4);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(ln_bwd_kernel_v3<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 512 * 4);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(ln_bwd_kernel_v3<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 1024 * 4);
      if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); return -3; }
      attr_set = true;
    }
    if (H == 768) launch_pdl(ln_bwd_kernel_v3<3>, grid, 256, smem, st, DY, Z, mean, rstd, gamma, DRI, DX, DR, dgamma, dbeta, dbias, M, dc);
    else if (H == 512) launch_pdl(ln_bwd_kernel_v3<2>, grid, 256, smem, st, DY, Z, mean, rstd, gamma, DRI, DX, DR, dgamma, dbeta, dbias, M, dc);
    else launch_pdl(ln_bwd_kernel_v3<4>, grid, 256, smem, st, DY, Z, mean, rstd, gamma, DRI, DX, DR, dgamma, dbeta, dbias, M, dc);
    return check_launch("ln_bwd_kernel_v3");
  }
  if (g_ln_variant == 1 || g_ln_variant == 2) {
    const size_t smem2 = (size_t)8 * 3 * H * sizeof(float);
#define HAMT_LNV2(NCH_)                                                                                                        \
  {                                                                                                                            \
    static bool set_ = false;                                                                                                  \
    if (!set_) {                                                                                                               \
      cudaFuncSetAttribute(ln_bwd_kernel_v2<NCH_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * NCH_ * 256 * 4); \
      cudaFuncSetAttribute(ln_bwd_kernel_v2<NCH_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * NCH_ * 256 * 4);\
      set_ = true;                                                                                                             \
    }                                                                                                                          \
    if (g_ln_variant == 2) launch_pdl(ln_bwd_kernel_v2<NCH_, true>, grid, 256, smem2, st, DY, Z, mean, rstd, gamma, DRI, DX, DR, dgamma, dbeta, dbias, M, dc); \
    else launch_pdl(ln_bwd_kernel_v2<NCH_, false>, grid, 256, smem2, st, DY, Z, mean, rstd, gamma, DRI, DX, DR, dgamma, dbeta, dbias, M, dc);                  \
  }
    if (H == 768) HAMT_LNV2(3) else if (H == 512) HAMT_LNV2(2) else HAMT_LNV2(4)
#undef HAMT_LNV2
    return check_launch("ln_bwd_kernel_v2");
  }
  if (H == 768) launch_pdl(ln_bwd_kernel<3>, grid, 256, 0, st, DY, Z, mean, rstd, gamma, DRI, DX, DR, dgamma, dbeta, dbias, M, dc);
  else if (H == 512) launch_pdl(ln_bwd_kernel<2>, grid, 256, 0, st, DY, Z, mean, rstd, gamma, DRI, DX, DR, dgamma, dbeta, dbias, M, dc);
  else launch_pdl(ln_bwd_kernel<4>, grid, 256, 0, st, DY, Z, mean, rstd, gamma, DRI, DX, DR, dgamma, dbeta, dbias, M, dc);
  return check_launch("ln_bwd_kernel");
}

// ---------------------------------------------------------------------------------------------
// streaming helpers
// ---------------------------------------------------------------------------------------------
__global__ void cast_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  pdl_grid_sync();
  const long long n8 = n >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(in)[2 * i];
    const float4 b = reinterpret_cast<const float4*>(in)[2 * i + 1];
    reinterpret_cast<uint4*>(out)[i] = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
  if (blockIdx.x == 0)
    for (long long i = (n8 << 3) + threadIdx.x; i < n; i += blockDim.x) out[i] = __float2bfloat16_rn(in[i]);
}
int cast_f32_to_bf16(const float* in, void* out, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  HAMT_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0, "cast: pointers must be 16-byte aligned");
  long long blocks = (n / 8 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_pdl(cast_kernel, (int)blocks, 256, 0, st, in, (__nv_bfloat16*)out, n);
  return check_launch("cast_kernel");
}

// out[n] += sum_m x[m,n].  Block = 32 x 8 threads; each block owns 64 columns (2 per thread-x) and a slice of rows.
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ x, long long ld, float* out, int M, int N, int rows_per_block) {
  pdl_grid_sync();
  __shared__ float sh[8][64];
  const int c0 = blockIdx.x * 64 + threadIdx.x * 2;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float a0 = 0.f, a1 = 0.f;
  if (c0 + 1 < N) {
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      const float2 f = unpack_bf16(*reinterpret_cast<const uint32_t*>(x + (long long)r * ld + c0));
      a0 += f.x; a1 += f.y;
    }
  } else if (c0 < N) {
    for (int r = r0 + threadIdx.y; r < r1; r += 8) a0 += __bfloat162float(x[(long long)r * ld + c0]);
  }
  sh[threadIdx.y][threadIdx.x * 2] = a0;
  sh[threadIdx.y][threadIdx.x * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.y == 0) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { a0 += sh[k][threadIdx.x * 2]; a1 += sh[k][threadIdx.x * 2 + 1]; }
    if (c0 < N) atomicAdd(out + c0, a0);
    if (c0 + 1 < N) atomicAdd(out + c0 + 1, a1);
  }
}
int colsum_bf16(const void* x, long long ld, float* out, int M, int N, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  HAMT_REQUIRE((ld & 1) == 0 && ((uintptr_t)x & 3) == 0, "colsum: pitch must be even and base 4-byte aligned");
  const int gx = (N + 63) / 64;
  int gy = (148 * 4 + gx - 1) / gx;
  int rpb = (M + gy - 1) / gy;
  if (rpb < 64) rpb = 64;
  gy = (M + rpb - 1) / rpb;
  launch_pdl(colsum_kernel, dim3(gx, gy), dim3(32, 8), 0, st, (const __nv_bfloat16*)x, ld, out, M, N, rpb);
  return check_launch("colsum_kernel");
}

// mean over the P tokens of each panorama (reference: torch.mean(dim=2), vilmodel.py:563-564 -- no masking)
__global__ void mean_pool_fwd_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int N, int P, int H) {
  pdl_grid_sync();
  const int n = blockIdx.x;
  for (int c = threadIdx.x * 2; c < H; c += blockDim.x * 2) {
    float a0 = 0.f, a1 = 0.f;
    for (int p = 0; p < P; ++p) {
      const float2 f = unpack_bf16(*reinterpret_cast<const uint32_t*>(x + ((long long)n * P + p) * H + c));
      a0 += f.x; a1 += f.y;
    }
    out[(long long)n * H + c] = a0 / P;
    out[(long long)n * H + c + 1] = a1 / P;
  }
}
__global__ void mean_pool_bwd_kernel(const float* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int N, int P, int H) {
  pdl_grid_sync();
  const int n = blockIdx.x;
  const float inv = 1.0f / P;
  for (int c = threadIdx.x * 2; c < H; c += blockDim.x * 2) {
    const uint32_t w = pack_bf16(dy[(long long)n * H + c] * inv, dy[(long long)n * H + c + 1] * inv);
    for (int p = 0; p < P; ++p) *reinterpret_cast<uint32_t*>(dx + ((long long)n * P + p) * H + c) = w;
  }
}
int mean_pool_fwd(const void* x, float* out, int N, int P, int H, cudaStream_t st) {
  if (N <= 0) return 0;
  HAMT_REQUIRE((H & 1) == 0, "mean_pool: H must be even");
  launch_pdl(mean_pool_fwd_kernel, N, 128, 0, st, (const __nv_bfloat16*)x, out, N, P, H);
  return check_launch("mean_pool_fwd_kernel");
}
int mean_pool_bwd(const float* dy, void* dx, int N, int P, int H, cudaStream_t st) {
  if (N <= 0) return 0;
  HAMT_REQUIRE((H & 1) == 0, "mean_pool: H must be even");
  launch_pdl(mean_pool_bwd_kernel, N, 128, 0, st, dy, (__nv_bfloat16*)dx, N, P, H);
  return check_launch("mean_pool_bwd_kernel");
}

__global__ void add_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ out, long long n) {
  pdl_grid_sync();
  const long long n8 = n >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 x = reinterpret_cast<const uint4*>(a)[i], y = reinterpret_cast<const uint4*>(b)[i];
    const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16(xs[k]), g = unpack_bf16(ys[k]);
      o[k] = pack_bf16(f.x + g.x, f.y + g.y);
    }
    reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
  if (blockIdx.x == 0)
    for (long long i = (n8 << 3) + threadIdx.x; i < n; i += blockDim.x) out[i] = __float2bfloat16_rn(__bfloat162float(a[i]) + __bfloat162float(b[i]));
}
int add_bf16(const void* a, const void* b, void* out, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  HAMT_REQUIRE((((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0, "add: pointers must be 16-byte aligned");
  long long blocks = (n / 8 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_pdl(add_kernel, (int)blocks, 256, 0, st, (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)out, n);
  return check_launch("add_kernel");
}

// out[b,s,:] = a[b,s,:] * v[b,:]   (SAP fusion  ob_embeds * txt_embeds[:, :1], pretrain_cmt.py:176)
__global__ void mul_rows_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ out, int S, int H) {
  pdl_grid_sync();
  const long long row = blockIdx.x;
  const int b = (int)(row / S);
  for (int c = threadIdx.x * 2; c < H; c += blockDim.x * 2) {
    const float2 x = unpack_bf16(*reinterpret_cast<const uint32_t*>(a + row * H + c));
    const float2 y = unpack_bf16(*reinterpret_cast<const uint32_t*>(v + (long long)b * H + c));
    *reinterpret_cast<uint32_t*>(out + row * H + c) = pack_bf16(x.x * y.x, x.y * y.y);
  }
}
int mul_rows_bf16(const void* a, const void* b, void* out, int B, int S, int H, cudaStream_t st) {
  if (B * S <= 0) return 0;
  HAMT_REQUIRE((H & 1) == 0, "mul_rows: H must be even");
  launch_pdl(mul_rows_kernel, B * S, 128, 0, st, (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)out, S, H);
  return check_launch("mul_rows_kernel");
}

}  // namespace hamt
