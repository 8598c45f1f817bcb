// Fused embedding kernels (forward + backward) for the three embedders of the HAMT backbone.
//
//  * text:     drop(LN(word[ids] + pos[s] + type[0]))                         vilmodel.py:54-69
//  * features: s = LN_img(t) + LN_ang(ang W_ang^T + b_ang) [+ type] [+ nav] [+ pano-mean] [+ pos]
//              out = drop(LN(s)) (or s when there is no final LN: pano tokens)  vilmodel.py:496-505, :549-571
//    where t = img_linear(x) comes from the tcgen05 GEMM.  The K=4 angle projection is done here in
//    fp32 registers; nothing but t (bf16) and the output round-trips HBM.  The backward recomputes the
//    forward row statistics instead of saving them.
// One warp per token row; lane l owns columns {c*256 + l*8 + j}; 128-bit loads of the 768-d rows.
#include <stdio.h>
#include "hamt_common.cuh"
#include "hamt_kernels.h"

namespace hamt {

template <int NCH>
__device__ __forceinline__ void ld_bf16_row(const __nv_bfloat16* p, int lane, float (&v)[NCH * 8]) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint4 w = *reinterpret_cast<const uint4*>(p + c * 256 + lane * 8);
    float2 f0 = unpack_bf16(w.x), f1 = unpack_bf16(w.y), f2 = unpack_bf16(w.z), f3 = unpack_bf16(w.w);
    v[c * 8 + 0] = f0.x; v[c * 8 + 1] = f0.y; v[c * 8 + 2] = f1.x; v[c * 8 + 3] = f1.y;
    v[c * 8 + 4] = f2.x; v[c * 8 + 5] = f2.y; v[c * 8 + 6] = f3.x; v[c * 8 + 7] = f3.y;
  }
}
template <int NCH>
__device__ __forceinline__ void st_bf16_row(__nv_bfloat16* p, int lane, const float (&v)[NCH * 8]) {
#pragma unroll
  for (int c = 0; c < NCH; ++c)
    *reinterpret_cast<uint4*>(p + c * 256 + lane * 8) =
        make_uint4(pack_bf16(v[c * 8 + 0], v[c * 8 + 1]), pack_bf16(v[c * 8 + 2], v[c * 8 + 3]), pack_bf16(v[c * 8 + 4], v[c * 8 + 5]),
                   pack_bf16(v[c * 8 + 6], v[c * 8 + 7]));
}
// v[i] (+)= p[col(i)]
template <int NCH, bool kAdd>
__device__ __forceinline__ void ld_f32_vec(const float* p, int lane, float (&v)[NCH * 8]) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p + c * 256 + lane * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p + c * 256 + lane * 8 + 4));
    const float t[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) v[c * 8 + j] = kAdd ? v[c * 8 + j] + t[j] : t[j];
  }
}
// in-place normalise: v <- (v - mean) * rstd ; returns rstd
template <int NCH>
__device__ __forceinline__ float normalize(float (&v)[NCH * 8], float eps) {
  constexpr int H = NCH * 256;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NCH * 8; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / H);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NCH * 8; ++i) { v[i] -= mean; q += v[i] * v[i]; }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / H) + eps);
#pragma unroll
  for (int i = 0; i < NCH * 8; ++i) v[i] *= rstd;
  return rstd;
}
// LN backward on a normalised row: g = upstream * gamma ; returns rstd * (g - mean(g) - xh * mean(g*xh)) in g
template <int NCH>
__device__ __forceinline__ void ln_back(float (&g)[NCH * 8], const float (&xh)[NCH * 8], float rstd) {
  constexpr int H = NCH * 256;
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NCH * 8; ++i) { s1 += g[i]; s2 += g[i] * xh[i]; }
  s1 = warp_sum(s1) * (1.0f / H);
  s2 = warp_sum(s2) * (1.0f / H);
#pragma unroll
  for (int i = 0; i < NCH * 8; ++i) g[i] = rstd * (g[i] - s1 - xh[i] * s2);
}
// accumulate into a WARP-PRIVATE shared-memory row (plain 128-bit read-modify-write, no atomics: shared fp32 atomics cost
// ~2 cycles per lane and made the first version of the backward 10x slower than its HBM time)
template <int NCH>
__device__ __forceinline__ void smem_acc(float* wacc, int lane, const float (&v)[NCH * 8]) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    float4* q = reinterpret_cast<float4*>(wacc + c * 256 + lane * 8);
    float4 a = q[0], b = q[1];
    a.x += v[c * 8 + 0]; a.y += v[c * 8 + 1]; a.z += v[c * 8 + 2]; a.w += v[c * 8 + 3];
    b.x += v[c * 8 + 4]; b.y += v[c * 8 + 5]; b.z += v[c * 8 + 6]; b.w += v[c * 8 + 7];
    q[0] = a; q[1] = b;
  }
}
template <int NCH>
__device__ __forceinline__ void gmem_acc(float* dst, int lane, const float (&v)[NCH * 8]) {
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(dst + c * 256 + lane * 8 + j, v[c * 8 + j]);
}
template <int NCH>
__device__ __forceinline__ void apply_drop(const DropState& ds, long long row, int lane, float (&v)[NCH * 8]) {
  if (!ds.on) return;
  constexpr int H = NCH * 256;
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) v[c * 8 + j] *= drop_mult(ds, (unsigned long long)row * H + c * 256 + lane * 8 + j);
}

// ------------------------------------------------------------------------------------- text
template <int NCH>
__global__ void __launch_bounds__(256) embed_text_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ word,
                                                             const float* __restrict__ pos, const float* __restrict__ type0,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             __nv_bfloat16* __restrict__ out, int M, int L, float eps, DropCfg dc) {
  pdl_grid_sync();
  constexpr int H = NCH * 256;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const DropState ds = drop_init(dc);
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += gridDim.x * wpb) {
    float e[NCH * 8];
    ld_f32_vec<NCH, false>(word + ids[row] * H, lane, e);
    ld_f32_vec<NCH, true>(pos + (long long)(row % L) * H, lane, e);
    ld_f32_vec<NCH, true>(type0, lane, e);
    normalize<NCH>(e, eps);
    float g[NCH * 8], b[NCH * 8];
    ld_f32_vec<NCH, false>(gamma, lane, g);
    ld_f32_vec<NCH, false>(beta, lane, b);
#pragma unroll
    for (int i = 0; i < NCH * 8; ++i) e[i] = e[i] * g[i] + b[i];
    apply_drop<NCH>(ds, row, lane, e);
    st_bf16_row<NCH>(out + (long long)row * H, lane, e);
  }
}

template <int NCH>
__global__ void __launch_bounds__(256) embed_text_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const long long* __restrict__ ids,
                                                             const float* __restrict__ word, const float* __restrict__ pos,
                                                             const float* __restrict__ type0, const float* __restrict__ gamma,
                                                             float* dword, float* dpos, float* dtype0, float* dgamma, float* dbeta, int M,
                                                             int L, float eps, DropCfg dc) {
  pdl_grid_sync();
  constexpr int H = NCH * 256;
  extern __shared__ float sall[];   // [warps][3][H]: dgamma, dbeta, dtype0 (warp-private rows)
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const DropState ds = drop_init(dc);
  for (int i = threadIdx.x; i < wpb * 3 * H; i += blockDim.x) sall[i] = 0.f;
  __syncthreads();
  float* sacc = sall + (threadIdx.x >> 5) * 3 * H;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += gridDim.x * wpb) {
    float xh[NCH * 8], d[NCH * 8], tmp[NCH * 8];
    const long long id = ids[row];
    ld_f32_vec<NCH, false>(word + id * H, lane, xh);
    ld_f32_vec<NCH, true>(pos + (long long)(row % L) * H, lane, xh);
    ld_f32_vec<NCH, true>(type0, lane, xh);
    const float rstd = normalize<NCH>(xh, eps);
    ld_bf16_row<NCH>(dy + (long long)row * H, lane, d);
    apply_drop<NCH>(ds, row, lane, d);
#pragma unroll
    for (int i = 0; i < NCH * 8; ++i) tmp[i] = d[i] * xh[i];
    smem_acc<NCH>(sacc, lane, tmp);
    smem_acc<NCH>(sacc + H, lane, d);
    ld_f32_vec<NCH, false>(gamma, lane, tmp);
#pragma unroll
    for (int i = 0; i < NCH * 8; ++i) d[i] *= tmp[i];
    ln_back<NCH>(d, xh, rstd);
    smem_acc<NCH>(sacc + 2 * H, lane, d);
    gmem_acc<NCH>(dword + id * H, lane, d);
    gmem_acc<NCH>(dpos + (long long)(row % L) * H, lane, d);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    float a = 0.f, b = 0.f, c = 0.f;
    for (int w = 0; w < wpb; ++w) { a += sall[(w * 3 + 0) * H + i]; b += sall[(w * 3 + 1) * H + i]; c += sall[(w * 3 + 2) * H + i]; }
    atomicAdd(dgamma + i, a);
    atomicAdd(dbeta + i, b);
    atomicAdd(dtype0 + i, c);
  }
}

// ------------------------------------------------------------------------------------- features
struct EmbP {
  const __nv_bfloat16* t; const float* ang; int A;
  const float *w_ang, *b_ang, *g_img, *b_img, *g_ang, *be_ang, *add_vec, *nav_table; const long long* nav_ids;
  const float* extra; const float* pos_table; const long long* pos_ids; int pos_mod;
  const float *g_f, *b_f;
  __nv_bfloat16* out; int M; float eps; DropCfg drop;
  // backward
  const __nv_bfloat16* dy; __nv_bfloat16* dt;
  float *dw_ang, *db_ang, *dg_img, *db_img, *dg_ang, *dbe_ang, *dadd_vec, *dnav_table, *dextra, *dpos_table, *dg_f, *db_f, *db_lin;
};

static constexpr int kMaxA = 8;

// shared-memory loads of a per-column parameter row
template <int NCH, bool kAdd>
__device__ __forceinline__ void lds_f32_vec(const float* sp, int lane, float (&v)[NCH * 8]) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const float4 a = *reinterpret_cast<const float4*>(sp + c * 256 + lane * 8);
    const float4 b = *reinterpret_cast<const float4*>(sp + c * 256 + lane * 8 + 4);
    const float t[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) v[c * 8 + j] = kAdd ? v[c * 8 + j] + t[j] : t[j];
  }
}

// stage W_ang transposed ([A][H], so that a lane's 8 columns are contiguous) + b_ang in shared memory
template <int NCH>
__device__ __forceinline__ void stage_wang(const EmbP& p, float* sw) {
  constexpr int H = NCH * 256;
  for (int i = threadIdx.x; i < p.A * H; i += blockDim.x) {
    const int a = i / H, col = i - a * H;
    sw[i] = __ldg(p.w_ang + (long long)col * p.A + a);
  }
  for (int i = threadIdx.x; i < H; i += blockDim.x) sw[p.A * H + i] = __ldg(p.b_ang + i);
}

// forward pieces shared by fwd and bwd: x1 = normalised t, x2 = normalised angle projection, s = sum
template <int NCH>
__device__ __forceinline__ void embed_row_forward(const EmbP& p, const float* sw, int row, int lane, float (&x1)[NCH * 8], float& rstd1,
                                                  float (&x2)[NCH * 8], float& rstd2, float (&s)[NCH * 8]) {
  constexpr int H = NCH * 256;
  ld_bf16_row<NCH>(p.t + (long long)row * H, lane, x1);
  rstd1 = normalize<NCH>(x1, p.eps);
  lds_f32_vec<NCH, false>(sw + p.A * H, lane, x2);                 // b_ang
  for (int a = 0; a < p.A; ++a) {                                    // K = angle_feat_size (4) projection in fp32 registers
    const float av = __ldg(p.ang + (long long)row * p.A + a);
    float w[NCH * 8];
    lds_f32_vec<NCH, false>(sw + a * H, lane, w);
#pragma unroll
    for (int i = 0; i < NCH * 8; ++i) x2[i] += av * w[i];
  }
  rstd2 = normalize<NCH>(x2, p.eps);
  float g[NCH * 8];
  ld_f32_vec<NCH, false>(p.g_img, lane, g);
#pragma unroll
  for (int i = 0; i < NCH * 8; ++i) s[i] = x1[i] * g[i];
  ld_f32_vec<NCH, true>(p.b_img, lane, s);
  ld_f32_vec<NCH, false>(p.g_ang, lane, g);
#pragma unroll
  for (int i = 0; i < NCH * 8; ++i) s[i] += x2[i] * g[i];
  ld_f32_vec<NCH, true>(p.be_ang, lane, s);
  if (p.add_vec) ld_f32_vec<NCH, true>(p.add_vec, lane, s);
  if (p.nav_table) ld_f32_vec<NCH, true>(p.nav_table + p.nav_ids[row] * H, lane, s);
  if (p.extra) ld_f32_vec<NCH, true>(p.extra + (long long)row * H, lane, s);
  if (p.pos_table) {
    const long long pid = p.pos_ids ? p.pos_ids[row] : (long long)(row % p.pos_mod);
    ld_f32_vec<NCH, true>(p.pos_table + pid * H, lane, s);
  }
}

template <int NCH>
__global__ void __launch_bounds__(256) embed_feat_fwd_kernel(const EmbP p) {
  pdl_grid_sync();
  constexpr int H = NCH * 256;
  extern __shared__ float sw[];     // [A+1][H]
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const DropState ds = drop_init(p.drop);
  stage_wang<NCH>(p, sw);
  __syncthreads();
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < p.M; row += gridDim.x * wpb) {
    float x1[NCH * 8], x2[NCH * 8], s[NCH * 8], r1, r2;
    embed_row_forward<NCH>(p, sw, row, lane, x1, r1, x2, r2, s);
    if (p.g_f) {
      normalize<NCH>(s, p.eps);
      ld_f32_vec<NCH, false>(p.g_f, lane, x1);
#pragma unroll
      for (int i = 0; i < NCH * 8; ++i) s[i] *= x1[i];
      ld_f32_vec<NCH, true>(p.b_f, lane, s);
    }
    apply_drop<NCH>(ds, row, lane, s);
    st_bf16_row<NCH>(p.out + (long long)row * H, lane, s);
  }
}

// shared accumulator rows
// accumulator rows (compacted at run time): DS, G1, G2, DU, DT, W[0..A), then GF, BF when there is a final LN, then NAV[0..3)
enum { ACC_DS = 0, ACC_G1, ACC_G2, ACC_DU, ACC_DT, ACC_W0 };
__host__ __device__ inline int acc_rows(int A, bool has_f, bool has_nav) { return ACC_W0 + A + (has_f ? 2 : 0) + (has_nav ? 3 : 0); }

template <int NCH>
__global__ void __launch_bounds__(256) embed_feat_bwd_kernel(const EmbP p) {
  pdl_grid_sync();
  constexpr int H = NCH * 256;
  extern __shared__ float smem_f[];   // [A+1][H] staged angle weights, then [warps][rows][H] warp-private accumulators
  float* sw = smem_f;
  float* sall = smem_f + (p.A + 1) * H;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const DropState ds = drop_init(p.drop);
  stage_wang<NCH>(p, sw);
  const int ACC_GF = ACC_W0 + p.A, ACC_BF = ACC_GF + 1, ACC_NAV0 = p.g_f ? ACC_GF + 2 : ACC_GF;
  const int nrows = acc_rows(p.A, p.g_f != nullptr, p.dnav_table != nullptr);
  for (int i = threadIdx.x; i < wpb * nrows * H; i += blockDim.x) sall[i] = 0.f;
  __syncthreads();
  float* sacc = sall + (threadIdx.x >> 5) * nrows * H;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < p.M; row += gridDim.x * wpb) {
    float x1[NCH * 8], x2[NCH * 8], s[NCH * 8], r1, r2;
    embed_row_forward<NCH>(p, sw, row, lane, x1, r1, x2, r2, s);
    float d[NCH * 8], tmp[NCH * 8];
    ld_bf16_row<NCH>(p.dy + (long long)row * H, lane, d);
    apply_drop<NCH>(ds, row, lane, d);
    if (p.g_f) {
      const float rf = normalize<NCH>(s, p.eps);   // s <- xhat_f
#pragma unroll
      for (int i = 0; i < NCH * 8; ++i) tmp[i] = d[i] * s[i];
      smem_acc<NCH>(sacc + ACC_GF * H, lane, tmp);
      smem_acc<NCH>(sacc + ACC_BF * H, lane, d);
      ld_f32_vec<NCH, false>(p.g_f, lane, tmp);
#pragma unroll
      for (int i = 0; i < NCH * 8; ++i) d[i] *= tmp[i];
      ln_back<NCH>(d, s, rf);
    }
    // d == ds (gradient of the pre-final-LN sum)
    smem_acc<NCH>(sacc + ACC_DS * H, lane, d);
    if (p.dnav_table) smem_acc<NCH>(sacc + (ACC_NAV0 + (int)p.nav_ids[row]) * H, lane, d);
    if (p.dextra) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        float* o = p.dextra + (long long)row * H + c * 256 + lane * 8;
        *reinterpret_cast<float4*>(o) = make_float4(d[c * 8], d[c * 8 + 1], d[c * 8 + 2], d[c * 8 + 3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(d[c * 8 + 4], d[c * 8 + 5], d[c * 8 + 6], d[c * 8 + 7]);
      }
    }
    if (p.dpos_table) {
      const long long pid = p.pos_ids ? p.pos_ids[row] : (long long)(row % p.pos_mod);
      gmem_acc<NCH>(p.dpos_table + pid * H, lane, d);
    }
    // image branch
#pragma unroll
    for (int i = 0; i < NCH * 8; ++i) tmp[i] = d[i] * x1[i];
    smem_acc<NCH>(sacc + ACC_G1 * H, lane, tmp);
    ld_f32_vec<NCH, false>(p.g_img, lane, tmp);
#pragma unroll
    for (int i = 0; i < NCH * 8; ++i) tmp[i] *= d[i];
    ln_back<NCH>(tmp, x1, r1);
    // round like the stored gradient so that the bias gradient matches what the wgrad GEMM sees
    st_bf16_row<NCH>(p.dt + (long long)row * H, lane, tmp);
    smem_acc<NCH>(sacc + ACC_DT * H, lane, tmp);
    // angle branch
#pragma unroll
    for (int i = 0; i < NCH * 8; ++i) tmp[i] = d[i] * x2[i];
    smem_acc<NCH>(sacc + ACC_G2 * H, lane, tmp);
    ld_f32_vec<NCH, false>(p.g_ang, lane, tmp);
#pragma unroll
    for (int i = 0; i < NCH * 8; ++i) tmp[i] *= d[i];
    ln_back<NCH>(tmp, x2, r2);
    smem_acc<NCH>(sacc + ACC_DU * H, lane, tmp);
    for (int a = 0; a < p.A; ++a) {
      const float av = __ldg(p.ang + (long long)row * p.A + a);
      float w[NCH * 8];
#pragma unroll
      for (int i = 0; i < NCH * 8; ++i) w[i] = tmp[i] * av;
      smem_acc<NCH>(sacc + (ACC_W0 + a) * H, lane, w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nrows * H; i += blockDim.x) {   // fold the warp-private rows into warp 0's
    float a = sall[i];
    for (int w = 1; w < wpb; ++w) a += sall[w * nrows * H + i];
    sall[i] = a;
  }
  __syncthreads();
  sacc = sall;
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    const float dsum = sacc[ACC_DS * H + i];
    atomicAdd(p.db_img + i, dsum);
    atomicAdd(p.dbe_ang + i, dsum);
    if (p.dadd_vec) atomicAdd(p.dadd_vec + i, dsum);
    atomicAdd(p.dg_img + i, sacc[ACC_G1 * H + i]);
    atomicAdd(p.dg_ang + i, sacc[ACC_G2 * H + i]);
    atomicAdd(p.db_ang + i, sacc[ACC_DU * H + i]);
    if (p.db_lin) atomicAdd(p.db_lin + i, sacc[ACC_DT * H + i]);   // bias grad of the image linear = column sum of dt
    if (p.dg_f) { atomicAdd(p.dg_f + i, sacc[ACC_GF * H + i]); atomicAdd(p.db_f + i, sacc[ACC_BF * H + i]); }
    for (int a = 0; a < p.A; ++a) atomicAdd(p.dw_ang + (long long)i * p.A + a, sacc[(ACC_W0 + a) * H + i]);
    if (p.dnav_table)
      for (int k = 0; k < 3; ++k) atomicAdd(p.dnav_table + (long long)k * H + i, sacc[(ACC_NAV0 + k) * H + i]);
  }
}

static EmbP to_embp(const EmbedFeatArgs& a) {
  EmbP p{};
  p.t = (const __nv_bfloat16*)a.t; p.ang = a.ang; p.A = a.A; p.w_ang = a.w_ang; p.b_ang = a.b_ang; p.g_img = a.g_img; p.b_img = a.b_img;
  p.g_ang = a.g_ang; p.be_ang = a.be_ang; p.add_vec = a.add_vec; p.nav_table = a.nav_table; p.nav_ids = a.nav_ids; p.extra = a.extra;
  p.pos_table = a.pos_table; p.pos_ids = a.pos_ids; p.pos_mod = a.pos_mod > 0 ? a.pos_mod : 1; p.g_f = a.g_f; p.b_f = a.b_f;
  p.out = (__nv_bfloat16*)a.out; p.M = a.M; p.eps = a.eps; p.drop = DropCfg{a.drop.seed_ptr, a.drop.site, a.drop.p};
  return p;
}
static int rows_grid(int M, int cap) {
  int g = (M + 7) / 8;
  if (g < 1) g = 1;
  return g < cap ? g : cap;
}

int embed_feat_fwd(const EmbedFeatArgs& a, cudaStream_t st) {
  HAMT_REQUIRE(a.H == 768 || a.H == 512 || a.H == 1024, "embed_feat: hidden size must be 512/768/1024");
  HAMT_REQUIRE(a.A >= 1 && a.A <= kMaxA, "embed_feat: angle feature size must be 1..8");
  HAMT_REQUIRE((a.nav_table == nullptr) == (a.nav_ids == nullptr), "embed_feat: nav_table and nav_ids go together");
  if (a.M <= 0) return 0;
  EmbP p = to_embp(a);
  const int grid = rows_grid(a.M, 148 * 8);
  const size_t smem = (size_t)(a.A + 1) * a.H * sizeof(float);
  if (a.H == 768) launch_pdl(embed_feat_fwd_kernel<3>, grid, 256, smem, st, p);
  else if (a.H == 512) launch_pdl(embed_feat_fwd_kernel<2>, grid, 256, smem, st, p);
  else launch_pdl(embed_feat_fwd_kernel<4>, grid, 256, smem, st, p);
  return check_launch("embed_feat_fwd_kernel");
}

template <int NCH>
static int launch_feat_bwd(const EmbP& p, int grid, cudaStream_t st) {
  const int nrows = acc_rows(p.A, p.g_f != nullptr, p.dnav_table != nullptr);
  const size_t per_warp = (size_t)nrows * NCH * 256 * sizeof(float);
  const size_t fixed = (size_t)(p.A + 1) * NCH * 256 * sizeof(float);
  int warps = (int)((224 * 1024 - fixed) / per_warp);
  if (warps > 8) warps = 8;
  if (warps < 1) { set_last_error("embed_feat_bwd: accumulator rows do not fit in shared memory"); return -1; }
  const size_t smem = fixed + per_warp * warps;
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(embed_feat_bwd_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); return -3; }
    set = true;
  }
  launch_pdl(embed_feat_bwd_kernel<NCH>, grid, warps * 32, smem, st, p);
  return check_launch("embed_feat_bwd_kernel");
}

int embed_feat_bwd(const EmbedFeatBwdArgs& a, cudaStream_t st) {
  const EmbedFeatArgs& f = a.f;
  HAMT_REQUIRE(f.H == 768 || f.H == 512 || f.H == 1024, "embed_feat: hidden size must be 512/768/1024");
  HAMT_REQUIRE(f.A >= 1 && f.A <= kMaxA, "embed_feat: angle feature size must be 1..8");
  HAMT_REQUIRE(a.dt && a.dw_ang && a.db_ang && a.dg_img && a.db_img && a.dg_ang && a.dbe_ang, "embed_feat_bwd: missing gradient buffers");
  HAMT_REQUIRE((f.g_f == nullptr) == (a.dg_f == nullptr), "embed_feat_bwd: final-LN grads must match forward");
  HAMT_REQUIRE(a.dnav_table == nullptr || f.nav_ids != nullptr, "embed_feat_bwd: dnav without nav ids");
  if (f.M <= 0) return 0;
  EmbP p = to_embp(f);
  p.dy = (const __nv_bfloat16*)a.dy; p.dt = (__nv_bfloat16*)a.dt; p.dw_ang = a.dw_ang; p.db_ang = a.db_ang; p.dg_img = a.dg_img;
  p.db_img = a.db_img; p.dg_ang = a.dg_ang; p.dbe_ang = a.dbe_ang; p.dadd_vec = a.dadd_vec; p.dnav_table = a.dnav_table; p.dextra = a.dextra;
  p.dpos_table = a.dpos_table; p.dg_f = a.dg_f; p.db_f = a.db_f; p.db_lin = a.db_lin;
  const int grid = rows_grid((f.M + 3) / 4, 148);
  if (f.H == 768) return launch_feat_bwd<3>(p, grid, st);
  if (f.H == 512) return launch_feat_bwd<2>(p, grid, st);
  return launch_feat_bwd<4>(p, grid, st);
}

int embed_text_fwd(const long long* ids, const float* word, const float* pos, const float* type0, const float* gamma, const float* beta, void* out,
                   int B, int L, int H, float eps, DropArgs drop, cudaStream_t st) {
  HAMT_REQUIRE(H == 768 || H == 512 || H == 1024, "embed_text: hidden size must be 512/768/1024");
  const int M = B * L;
  if (M <= 0) return 0;
  DropCfg dc{drop.seed_ptr, drop.site, drop.p};
  const int grid = rows_grid(M, 148 * 8);
  auto O = (__nv_bfloat16*)out;
  if (H == 768) launch_pdl(embed_text_fwd_kernel<3>, grid, 256, 0, st, ids, word, pos, type0, gamma, beta, O, M, L, eps, dc);
  else if (H == 512) launch_pdl(embed_text_fwd_kernel<2>, grid, 256, 0, st, ids, word, pos, type0, gamma, beta, O, M, L, eps, dc);
  else launch_pdl(embed_text_fwd_kernel<4>, grid, 256, 0, st, ids, word, pos, type0, gamma, beta, O, M, L, eps, dc);
  return check_launch("embed_text_fwd_kernel");
}

int embed_text_bwd(const void* dy, const long long* ids, const float* word, const float* pos, const float* type0, const float* gamma, float* dword,
                   float* dpos, float* dtype0, float* dgamma, float* dbeta, int B, int L, int H, float eps, DropArgs drop, cudaStream_t st) {
  HAMT_REQUIRE(H == 768 || H == 512 || H == 1024, "embed_text: hidden size must be 512/768/1024");
  const int M = B * L;
  if (M <= 0) return 0;
  DropCfg dc{drop.seed_ptr, drop.site, drop.p};
  const int grid = rows_grid((M + 3) / 4, 148 * 2);
  const size_t smem = (size_t)8 * 3 * H * sizeof(float);
  static bool set = false;
  if (!set) {
    cudaFuncSetAttribute(embed_text_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(embed_text_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(embed_text_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    set = true;
  }
  auto DY = (const __nv_bfloat16*)dy;
  if (H == 768) launch_pdl(embed_text_bwd_kernel<3>, grid, 256, smem, st, DY, ids, word, pos, type0, gamma, dword, dpos, dtype0, dgamma, dbeta, M, L, eps, dc);
  else if (H == 512) launch_pdl(embed_text_bwd_kernel<2>, grid, 256, smem, st, DY, ids, word, pos, type0, gamma, dword, dpos, dtype0, dgamma, dbeta, M, L, eps, dc);
  else launch_pdl(embed_text_bwd_kernel<4>, grid, 256, smem, st, DY, ids, word, pos, type0, gamma, dword, dpos, dtype0, dgamma, dbeta, M, L, eps, dc);
  return check_launch("embed_text_bwd_kernel");
}

}  // namespace hamt
