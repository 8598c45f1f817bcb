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

// The residual stream may be carried in fp32 next to its bf16 copy: `res32` (fp32 [M,H]) replaces `res`, and `y32` receives the
// un-rounded LayerNorm output for the next block's residual add.  The GEMMs still read the bf16 `y`; only the skip connection keeps
// full precision, which is what torch.autocast does to the reference (layer_norm runs in fp32 there) -- with a bf16-only stream the
// logits sit 1.2 .. 1.7x further from the fp32 reference than the reference's own autocast run, with the fp32 stream 0.5 .. 1.0x
// (measured with the bf16-regime oracle).
template <int NCH>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const __nv_bfloat16* x /* may alias z_out */, const __nv_bfloat16* __restrict__ res,
                                                     const float* __restrict__ res32, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, float* __restrict__ y32,
                                                     __nv_bfloat16* z_out, float* __restrict__ z32, float* __restrict__ mean_out,
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
    if (x != nullptr) load_row_bf16<NCH>(x + (long long)row * H, lane, z);
    else {                    // pre-LN entry: the row is the fp32 stream itself (res32)
#pragma unroll
      for (int i = 0; i < NCH * 8; ++i) z[i] = 0.f;
    }
    if (ds.on && x != nullptr) {
#pragma unroll
      for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) z[c * 8 + j] *= drop_mult(ds, (unsigned long long)row * H + c * 256 + lane * 8 + j);
    }
    if (res32 != nullptr) {
      float r[NCH * 8];
      load_vec_f32<NCH>(res32 + (long long)row * H, lane, r);
#pragma unroll
      for (int i = 0; i < NCH * 8; ++i) z[i] += r[i];
    } else if (res != nullptr) {
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
    if (z32 != nullptr) {     // pre-LN blocks (ViT): the un-normalised sum IS the residual stream, carried in fp32
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        float* p32 = z32 + (long long)row * H + c * 256 + lane * 8;
        *reinterpret_cast<float4*>(p32) = make_float4(z[c * 8], z[c * 8 + 1], z[c * 8 + 2], z[c * 8 + 3]);
        *reinterpret_cast<float4*>(p32 + 4) = make_float4(z[c * 8 + 4], z[c * 8 + 5], z[c * 8 + 6], z[c * 8 + 7]);
      }
    }
    float o[NCH * 8];
#pragma unroll
    for (int i = 0; i < NCH * 8; ++i) o[i] = (z[i] - mean) * rstd * g[i] + b[i];
    store_row_bf16<NCH>(y + (long long)row * H, lane, o);
    if (y32 != nullptr) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        float* p32 = y32 + (long long)row * H + c * 256 + lane * 8;
        *reinterpret_cast<float4*>(p32) = make_float4(o[c * 8], o[c * 8 + 1], o[c * 8 + 2], o[c * 8 + 3]);
        *reinterpret_cast<float4*>(p32 + 4) = make_float4(o[c * 8 + 4], o[c * 8 + 5], o[c * 8 + 6], o[c * 8 + 7]);
      }
    }
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
  }
}

// Backward of y = LN(dropout(x) + res): persistent (one 8-warp CTA per SM, each warp walks rows with stride = total warps).
// Round-1 version of this kernel ran at 1.2 TB/s for the small-M launches that make up 40 of its 44 launches per step
// (profiles/r01_kbench_v5.txt; round-2 measurement of its variants in profiles/r02_kbench_variants.txt): it launched M/32 CTAs that
// need ~240 registers (one CTA per SM) -> 160 CTAs = two waves at M = 5120, each wave a serial chain of 4 rows x (load, reduce, load,
// store) with nothing in flight, and folded its column sums with 72 shared-memory fp32 atomics per lane (CAS loops on sm_100).  Here:
//   * the grid never exceeds the SM count (one wave), rows are distributed over all warps of the chip;
//   * the NEXT row's dy / z / residual-gradient chunks are requested (raw bf16, 36 registers) before the current row is reduced, so
//     every warp always has one row in flight (8 warps x 4.6 KB per SM);
//   * the second phase recomputes xhat / gamma*dy from the raw chunks instead of keeping 48 fp32 values alive;
//   * column sums (dgamma, dbeta, dbias) stay in registers across rows and are folded once per CTA through warp-private,
//     bank-conflict-free shared slabs (no shared atomics), then one red.global.add.v4.f32 per 4 columns and CTA.
template <int NCH>
__global__ void __launch_bounds__(256, 1) ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
                                                        const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                        const float* __restrict__ gamma, const __nv_bfloat16* __restrict__ dres_in,
                                                        __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dres,
                                                        float* dgamma, float* dbeta, float* dbias, int M, DropCfg dc, int prenorm) {
  pdl_grid_sync();
  constexpr int H = NCH * 256;
  constexpr int NV = NCH * 8;                          // values per lane
  extern __shared__ float slab_all[];                  // [warps][3][NV][32] warp-private, written once at the end
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const DropState ds = drop_init(dc);
  const bool has_in = dres != nullptr && dres_in != nullptr;
  float g[NV];
  load_vec_f32<NCH>(gamma, lane, g);
  float acc_g[NV], acc_b[NV], acc_x[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc_g[i] = acc_b[i] = acc_x[i] = 0.f;
  const int stride = gridDim.x * wpb;
  int row = blockIdx.x * wpb + warp;
  uint4 nd[NCH], nz[NCH], nr[NCH];
  float nmean = 0.f, nrstd = 0.f;
  auto issue = [&](int r) {
    const long long o = (long long)r * H + lane * 8;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      nd[c] = *reinterpret_cast<const uint4*>(dy + o + c * 256);
      nz[c] = *reinterpret_cast<const uint4*>(z + o + c * 256);
      if (has_in) nr[c] = *reinterpret_cast<const uint4*>(dres_in + o + c * 256);
    }
    nmean = mean_in[r];
    nrstd = rstd_in[r];
  };
  auto unpack8 = [](const uint4& w, float (&f)[8]) {
    const float2 a = unpack_bf16(w.x), b = unpack_bf16(w.y), c = unpack_bf16(w.z), d = unpack_bf16(w.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  };
  auto pack8 = [](const float (&f)[8]) {
    return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  };
  if (row < M) issue(row);
  for (; row < M; row += stride) {
    uint4 cd[NCH], cz[NCH], cr[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) { cd[c] = nd[c]; cz[c] = nz[c]; cr[c] = has_in ? nr[c] : make_uint4(0, 0, 0, 0); }
    const float mean = nmean, rstd = nrstd;
    if (row + stride < M) issue(row + stride);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      float d[8], zz[8];
      unpack8(cd[c], d);
      unpack8(cz[c], zz);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (zz[j] - mean) * rstd;
        const float gd = d[j] * g[c * 8 + j];
        acc_g[c * 8 + j] = fmaf(d[j], xh, acc_g[c * 8 + j]);
        acc_b[c * 8 + j] += d[j];
        s1 += gd;
        s2 = fmaf(gd, xh, s2);
      }
    }
    s1 = warp_sum(s1) * (1.0f / H);
    s2 = warp_sum(s2) * (1.0f / H);
    const long long o = (long long)row * H + lane * 8;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      float d[8], zz[8], dz[8];
      unpack8(cd[c], d);
      unpack8(cz[c], zz);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (zz[j] - mean) * rstd;
        dz[j] = rstd * (d[j] * g[c * 8 + j] - s1 - xh * s2);
      }
      if (dres != nullptr) {
        float o8[8];
        if (has_in) {
          unpack8(cr[c], o8);
#pragma unroll
          for (int j = 0; j < 8; ++j) o8[j] += dz[j];
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) o8[j] = dz[j];
        }
        *reinterpret_cast<uint4*>(dres + o + c * 256) = pack8(o8);
        // pre-LN blocks: z is the residual STREAM (consumed again further down), so the sublayer output's gradient is the total
        // gradient of z -- LayerNorm path + incoming stream gradient -- not the LayerNorm path alone
        if (prenorm) {
#pragma unroll
          for (int j = 0; j < 8; ++j) dz[j] = o8[j];
        }
      }
      if (dx != nullptr) {
        if (ds.on) {
#pragma unroll
          for (int j = 0; j < 8; ++j) dz[j] *= drop_mult(ds, (unsigned long long)row * H + c * 256 + lane * 8 + j);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc_x[c * 8 + j] += dz[j];
        *reinterpret_cast<uint4*>(dx + o + c * 256) = pack8(dz);
      }
    }
  }
  float* slab = slab_all + warp * (3 * H) + lane;       // value index major, lane minor: 32 consecutive banks per store
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    slab[(0 * NV + i) * 32] = acc_g[i];
    slab[(1 * NV + i) * 32] = acc_b[i];
    slab[(2 * NV + i) * 32] = acc_x[i];
  }
  __syncthreads();
  // fold the warp-private slabs: thread t owns 4 consecutive COLUMNS of one accumulator -> one vector red per thread and pass
  for (int q = threadIdx.x; q < 3 * H / 4; q += blockDim.x) {
    const int a = q / (H / 4), col = (q % (H / 4)) * 4;
    float* dst = a == 0 ? dgamma : (a == 1 ? dbeta : dbias);
    if (dst == nullptr) continue;
    // column col + k  <-  value i = (col / 256) * 8 + (col % 8) + k of lane l = (col % 256) / 8
    const int i0 = (col >> 8) * 8 + (col & 7), l = (col & 255) >> 3;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    for (int w = 0; w < wpb; ++w) {
      const float* sl = slab_all + w * (3 * H) + (a * NV + i0) * 32 + l;
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] += sl[k * 32];
    }
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + col), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
  }
}

static int grid_for_rows(int M, int wpb, int max_ctas) {
  int g = (M + wpb - 1) / wpb;
  return g < max_ctas ? (g < 1 ? 1 : g) : max_ctas;
}

int ln_fwd(const void* x, const void* res, const float* res32, const float* gamma, const float* beta, void* y, float* y32, void* z_out,
           float* z32, float* mean, float* rstd, int M, int H, float eps, DropArgs drop, cudaStream_t st) {
  HAMT_REQUIRE(H == 512 || H == 768 || H == 1024, "ln_fwd: hidden size must be 512/768/1024");
  HAMT_REQUIRE(x != nullptr || res32 != nullptr, "ln_fwd: no input (x and res32 both null)");
  HAMT_REQUIRE(((uintptr_t)z32 & 15) == 0, "ln_fwd: z32 must be 16-byte aligned");
  HAMT_REQUIRE(res == nullptr || res32 == nullptr, "ln_fwd: the residual comes either as bf16 or as fp32, not both");
  HAMT_REQUIRE((((uintptr_t)res32 | (uintptr_t)y32) & 15) == 0, "ln_fwd: fp32 residual / output must be 16-byte aligned");
  if (M <= 0) return 0;
  DropCfg dc{drop.seed_ptr, drop.site, drop.p};
  const int grid = grid_for_rows(M, 8, 148 * 8);
  auto X = (const __nv_bfloat16*)x; auto R = (const __nv_bfloat16*)res; auto Y = (__nv_bfloat16*)y; auto Z = (__nv_bfloat16*)z_out;
  if (H == 768) launch_pdl(ln_fwd_kernel<3>, grid, 256, 0, st, X, R, res32, gamma, beta, Y, y32, Z, z32, mean, rstd, M, eps, dc);
  else if (H == 512) launch_pdl(ln_fwd_kernel<2>, grid, 256, 0, st, X, R, res32, gamma, beta, Y, y32, Z, z32, mean, rstd, M, eps, dc);
  else launch_pdl(ln_fwd_kernel<4>, grid, 256, 0, st, X, R, res32, gamma, beta, Y, y32, Z, z32, mean, rstd, M, eps, dc);
  return check_launch("ln_fwd_kernel");
}

int ln_bwd(const void* dy, const void* z, const float* mean, const float* rstd, const float* gamma, const void* dres_in, void* dx, void* dres,
           float* dgamma, float* dbeta, float* dbias, int M, int H, DropArgs drop, int prenorm, cudaStream_t st) {
  HAMT_REQUIRE(!prenorm || dres != nullptr, "ln_bwd (pre-LN): the stream gradient output is required");
  HAMT_REQUIRE(H == 512 || H == 768 || H == 1024, "ln_bwd: hidden size must be 512/768/1024");
  HAMT_REQUIRE((((uintptr_t)dgamma | (uintptr_t)dbeta | (uintptr_t)dbias) & 15) == 0, "ln_bwd: column-sum outputs must be 16-byte aligned");
  if (M <= 0) return 0;
  DropCfg dc{drop.seed_ptr, drop.site, drop.p};
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int grid = grid_for_rows(M, 8, sms);           // one wave: the kernel holds one CTA per SM
  const size_t smem = (size_t)8 * 3 * H * sizeof(float);   // 8 warps x [3][H] fp32: 72 KB at H = 768
  auto DY = (const __nv_bfloat16*)dy; auto Z = (const __nv_bfloat16*)z; auto DRI = (const __nv_bfloat16*)dres_in;
  auto DX = (__nv_bfloat16*)dx; auto DR = (__nv_bfloat16*)dres;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(ln_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 768 * 4);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ln_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 512 * 4);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ln_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 1024 * 4);
    if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); return -3; }
    attr_set = true;
  }
  if (H == 768) launch_pdl(ln_bwd_kernel<3>, grid, 256, smem, st, DY, Z, mean, rstd, gamma, DRI, DX, DR, dgamma, dbeta, dbias, M, dc, prenorm);
  else if (H == 512) launch_pdl(ln_bwd_kernel<2>, grid, 256, smem, st, DY, Z, mean, rstd, gamma, DRI, DX, DR, dgamma, dbeta, dbias, M, dc, prenorm);
  else launch_pdl(ln_bwd_kernel<4>, grid, 256, smem, st, DY, Z, mean, rstd, gamma, DRI, DX, DR, dgamma, dbeta, dbias, M, dc, prenorm);
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
