// Small kernels of the proxy-task heads (pretrain_src/model/pretrain_cmt.py:13-71, :142-262):
//  * rowdot: the final Linear(768 -> 1|2|3) of the SAP / SAR / SPREL / ITM heads -- too narrow for a
//    tensor-core tile, done as warp-per-row dot products in fp32;
//  * cross-entropy forward / backward over fp32 logits (handles the -inf entries produced by
//    masked_fill_(nav_type == 0, -inf), pretrain_cmt.py:177; MLM rows of 30522 logits, :150-153);
//  * row gather / scatter for `hidden[mask]` (_compute_masked_hidden, pretrain_cmt.py:161-165).
#include <stdio.h>
#include "hamt_common.cuh"
#include "hamt_kernels.h"
#include "../../include/hamt_b200.h"

namespace hamt {

static constexpr int kMaxN = 4;

// y[m,n] = sum_k x[m,k] w[n,k] + b[n]
__global__ void __launch_bounds__(256) rowdot_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ b, float* __restrict__ y, int M, int N, int H) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += gridDim.x * wpb) {
    float acc[kMaxN] = {0.f, 0.f, 0.f, 0.f};
    for (int c = lane * 2; c < H; c += 64) {
      const float2 f = unpack_bf16(*reinterpret_cast<const uint32_t*>(x + (long long)row * H + c));
#pragma unroll
      for (int n = 0; n < kMaxN; ++n)
        if (n < N) acc[n] += f.x * __ldg(w + (long long)n * H + c) + f.y * __ldg(w + (long long)n * H + c + 1);
    }
#pragma unroll
    for (int n = 0; n < kMaxN; ++n) {
      if (n < N) {
        const float s = warp_sum(acc[n]);
        if (lane == 0) y[(long long)row * N + n] = s + (b ? b[n] : 0.f);
      }
    }
  }
}

// dx[m,:] = sum_n dy[m,n] w[n,:] ; dw[n,:] += sum_m dy[m,n] x[m,:] ; db[n] += sum_m dy[m,n]
__global__ void __launch_bounds__(256) rowdot_bwd_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                         const float* __restrict__ w, __nv_bfloat16* __restrict__ dx, float* dw, float* db,
                                                         int M, int N, int H) {
  pdl_grid_sync();
  extern __shared__ float sdw[];   // [N][H] + [N]
  for (int i = threadIdx.x; i < N * H + N; i += blockDim.x) sdw[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += gridDim.x * wpb) {
    float g[kMaxN];
#pragma unroll
    for (int n = 0; n < kMaxN; ++n) g[n] = n < N ? dy[(long long)row * N + n] : 0.f;
    for (int c = lane * 2; c < H; c += 64) {
      const float2 f = unpack_bf16(*reinterpret_cast<const uint32_t*>(x + (long long)row * H + c));
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int n = 0; n < kMaxN; ++n)
        if (n < N) {
          d0 += g[n] * __ldg(w + (long long)n * H + c);
          d1 += g[n] * __ldg(w + (long long)n * H + c + 1);
          atomicAdd(&sdw[n * H + c], g[n] * f.x);
          atomicAdd(&sdw[n * H + c + 1], g[n] * f.y);
        }
      if (dx) *reinterpret_cast<uint32_t*>(dx + (long long)row * H + c) = pack_bf16(d0, d1);
    }
    if (lane == 0)
      for (int n = 0; n < N; ++n) atomicAdd(&sdw[N * H + n], g[n]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N * H; i += blockDim.x) atomicAdd(dw + i, sdw[i]);
  if (db)
    for (int i = threadIdx.x; i < N; i += blockDim.x) atomicAdd(db + i, sdw[N * H + i]);
}

// block-per-row cross entropy over fp32 logits; -inf logits contribute exp() = 0
__device__ __forceinline__ float block_reduce(float v, bool is_max, float* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = is_max ? -INFINITY : 0.f;
  for (int i = 0; i < nw; ++i) r = is_max ? fmaxf(r, sh[i]) : r + sh[i];
  return r;
}
__global__ void __launch_bounds__(256) ce_fwd_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ labels,
                                                     float* __restrict__ loss, float* __restrict__ lse, int N) {
  pdl_grid_sync();
  __shared__ float sh[8];
  const long long row = blockIdx.x;
  const float* p = logits + row * ld;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < N; i += blockDim.x) mx = fmaxf(mx, p[i]);
  mx = block_reduce(mx, true, sh);
  float s = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s += __expf(p[i] - mx);
  s = block_reduce(s, false, sh);
  if (threadIdx.x == 0) {
    const float l = mx + __logf(s);
    lse[row] = l;
    loss[row] = l - p[labels[row]];
  }
}
__global__ void __launch_bounds__(256) ce_bwd_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ labels,
                                                     const float* __restrict__ lse, const float* __restrict__ gloss, float* dl_f32,
                                                     __nv_bfloat16* dl_bf16, long long ld_d, int N, int N_pad) {
  pdl_grid_sync();
  const long long row = blockIdx.x;
  const float* p = logits + row * ld;
  const float l = lse[row], g = gloss[row];
  const long long lab = labels[row];
  for (int i = threadIdx.x; i < N_pad; i += blockDim.x) {
    float d = 0.f;
    if (i < N) d = g * (__expf(p[i] - l) - (i == lab ? 1.f : 0.f));
    if (dl_f32) { if (i < N) dl_f32[row * ld_d + i] = d; }
    else dl_bf16[row * ld_d + i] = __float2bfloat16_rn(d);
  }
}

// out[i,:] = x[idx[i],:]   /   out[idx[i],:] = x[i,:]   (rows of H bf16, H % 8 == 0)
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ x, const long long* __restrict__ idx, __nv_bfloat16* __restrict__ out, int n,
                                   int H, int scatter) {
  pdl_grid_sync();
  const int per_row = H / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)n * per_row; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / per_row;
    const int c = (int)(i % per_row) * 8;
    const long long src = scatter ? r : idx[r], dst = scatter ? idx[r] : r;
    *reinterpret_cast<uint4*>(out + dst * H + c) = *reinterpret_cast<const uint4*>(x + src * H + c);
  }
}

// out[i,:] = idx[i] >= 0 ? x[idx[i],:] : 0   (rows of H bf16, H % 8 == 0).  Device-side batch assembly from a resident feature table
// (SURVEY 8 f4): replaces the host-side np.stack / pad_tensors of pretrain_src/data/r2r_data.py:264-329 + common.py:5-20; padded
// history steps and "killed" observations (r2r_tasks.py:322-324) are zero rows.
__global__ void gather_rows_pad_kernel(const __nv_bfloat16* __restrict__ x, const long long* __restrict__ idx, __nv_bfloat16* __restrict__ out,
                                       long long n, int H) {
  pdl_grid_sync();
  const int per_row = H / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n * per_row; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / per_row;
    const int c = (int)(i % per_row) * 8;
    const long long src = idx[r];
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (src >= 0) v = *reinterpret_cast<const uint4*>(x + src * H + c);
    *reinterpret_cast<uint4*>(out + r * H + c) = v;
  }
}

}  // namespace hamt

using namespace hamt;

extern "C" {

int hamt_rowdot_fwd(const void* x, const float* w, const float* b, float* y, int M, int N, int H, void* stream) {
  HAMT_REQUIRE(N >= 1 && N <= kMaxN && (H & 1) == 0, "rowdot: N must be 1..4 and H even");
  if (M <= 0) return 0;
  int grid = (M + 7) / 8;
  if (grid > 148 * 8) grid = 148 * 8;
  launch_pdl(rowdot_fwd_kernel, grid, 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, w, b, y, M, N, H);
  return check_launch("rowdot_fwd_kernel");
}
int hamt_rowdot_bwd(const float* dy, const void* x, const float* w, void* dx, float* dw, float* db, int M, int N, int H, void* stream) {
  HAMT_REQUIRE(N >= 1 && N <= kMaxN && (H & 1) == 0, "rowdot: N must be 1..4 and H even");
  HAMT_REQUIRE((size_t)(N * H + N) * 4 <= 48 * 1024, "rowdot_bwd: N*H too large");
  if (M <= 0) return 0;
  int grid = (M + 31) / 32;
  if (grid > 148) grid = 148;
  launch_pdl(rowdot_bwd_kernel, grid, 256, (size_t)(N * H + N) * 4, (cudaStream_t)stream, dy, (const __nv_bfloat16*)x, w, (__nv_bfloat16*)dx, dw, db, M, N, H);
  return check_launch("rowdot_bwd_kernel");
}
int hamt_ce_fwd(const float* logits, long long ld, const long long* labels, float* loss, float* lse, int M, int N, void* stream) {
  if (M <= 0) return 0;
  HAMT_REQUIRE(N > 0, "ce: empty class dimension");
  launch_pdl(ce_fwd_kernel, M, 256, 0, (cudaStream_t)stream, logits, ld, labels, loss, lse, N);
  return check_launch("ce_fwd_kernel");
}
int hamt_ce_bwd(const float* logits, long long ld, const long long* labels, const float* lse, const float* gloss, float* dl_f32, void* dl_bf16,
                long long ld_d, int M, int N, void* stream) {
  if (M <= 0) return 0;
  HAMT_REQUIRE((dl_f32 != nullptr) != (dl_bf16 != nullptr), "ce_bwd: exactly one of dl_f32 / dl_bf16");
  const int n_pad = dl_bf16 ? (int)ld_d : N;   // bf16 output is written out to the padded pitch (zeros) for the TMA GEMM
  launch_pdl(ce_bwd_kernel, M, 256, 0, (cudaStream_t)stream, logits, ld, labels, lse, gloss, dl_f32, (__nv_bfloat16*)dl_bf16, ld_d, N, n_pad);
  return check_launch("ce_bwd_kernel");
}
int hamt_gather_rows_bf16(const void* x, const long long* idx, void* out, int n, int H, void* stream) {
  if (n <= 0) return 0;
  HAMT_REQUIRE(H % 8 == 0, "gather_rows: H must be a multiple of 8");
  int grid = (int)(((long long)n * (H / 8) + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  launch_pdl(gather_rows_kernel, grid, 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, idx, (__nv_bfloat16*)out, n, H, 0);
  return check_launch("gather_rows_kernel");
}
int hamt_gather_rows_pad_bf16(const void* x, long long x_rows, const long long* idx, void* out, long long n, int H, void* stream) {
  if (n <= 0) return 0;
  HAMT_REQUIRE(H % 8 == 0 && x_rows > 0, "gather_rows_pad: H must be a multiple of 8 and the table non-empty");
  HAMT_REQUIRE((((uintptr_t)x | (uintptr_t)out) & 15) == 0, "gather_rows_pad: table and output must be 16-byte aligned");
  long long grid = (n * (H / 8) + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  launch_pdl(gather_rows_pad_kernel, (int)grid, 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, idx, (__nv_bfloat16*)out, n, H);
  return check_launch("gather_rows_pad_kernel");
}
int hamt_scatter_rows_bf16(const void* x, const long long* idx, void* out, int n, int H, void* stream) {
  if (n <= 0) return 0;
  HAMT_REQUIRE(H % 8 == 0, "scatter_rows: H must be a multiple of 8");
  int grid = (int)(((long long)n * (H / 8) + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  launch_pdl(gather_rows_kernel, grid, 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, idx, (__nv_bfloat16*)out, n, H, 1);
  return check_launch("scatter_rows_kernel");
}

}  // extern "C"
