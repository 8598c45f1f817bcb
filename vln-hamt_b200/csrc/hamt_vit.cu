// Row kernels of the end-to-end ViT-B/16 stage: everything around the patch-projection GEMM and the 12 pre-LN blocks
// that is not already a GEMM / attention / LayerNorm kernel of the feature-based path.
//
//   patchify      images fp32 [N, C, Hh, Ww] -> bf16 [N * gh * gw, C * ps * ps], columns ordered (channel, row, column) = the flattening
//                 of the Conv2d weight [E, C, ps, ps], so that PatchEmbed's Conv2d(k = s = 16) (pretrain_src/model/vision_transformer.py:
//                 213, :221) becomes one tcgen05 GEMM [N*196, 768] x [768, 768]; the fp32 -> bf16 cast of the pixels happens in this pass.
//   vit_embed     x = pos_drop([cls ; patch tokens] + pos_embed)   (vision_transformer.py:337-342): fp32 residual stream + its bf16 copy;
//                 backward: masked gradient (for the pos_embed / cls_token column sums) and the patch-token rows (for the GEMM wgrad).
// HBM-bound streaming kernels: 16-byte accesses, one warp per token row.
#include <stdio.h>
#include "hamt_common.cuh"
#include "hamt_kernels.h"

namespace hamt {

__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, long long chunks, int C,
                                                       int Hh, int Ww, int ps, int gh, int gw) {
  pdl_grid_sync();
  const int cpr = C * ps * ps / 8;              // 16-byte output chunks per patch row
  const int kx_chunks = ps / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += (long long)gridDim.x * blockDim.x) {
    const long long prow = i / cpr;             // patch index n * gh * gw + gy * gw + gx
    const int cc = (int)(i - prow * cpr);       // chunk inside the patch row: (c, ky, kx / 8)
    const int kxc = cc % kx_chunks, ky = (cc / kx_chunks) % ps, c = cc / (kx_chunks * ps);
    const int gx = (int)(prow % gw), gy = (int)((prow / gw) % gh);
    const long long n = prow / ((long long)gw * gh);
    const float* src = img + ((n * C + c) * Hh + (long long)gy * ps + ky) * Ww + gx * ps + kxc * 8;
    const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
    *reinterpret_cast<uint4*>(out + i * 8) = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
}

int patchify_bf16(const float* img, void* out, int N, int C, int Hh, int Ww, int ps, cudaStream_t st) {
  HAMT_REQUIRE(ps % 8 == 0 && Hh % ps == 0 && Ww % ps == 0, "patchify: the patch size must be a multiple of 8 and divide the image");
  HAMT_REQUIRE(Ww % 4 == 0 && (((uintptr_t)img | (uintptr_t)out) & 15) == 0, "patchify: 16-byte aligned rows required");
  if (N <= 0) return 0;
  const int gh = Hh / ps, gw = Ww / ps;
  const long long chunks = (long long)N * gh * gw * (C * ps * ps / 8);
  long long blocks = (chunks + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_pdl(patchify_kernel, (int)blocks, 256, 0, st, img, (__nv_bfloat16*)out, chunks, C, Hh, Ww, ps, gh, gw);
  return check_launch("patchify_kernel");
}

template <int NCH>
__global__ void __launch_bounds__(256) vit_embed_fwd_kernel(const __nv_bfloat16* __restrict__ t0, const float* __restrict__ cls,
                                                            const float* __restrict__ pos, float* __restrict__ x32, __nv_bfloat16* __restrict__ x16,
                                                            int N, int S, DropCfg dc) {
  pdl_grid_sync();
  constexpr int H = NCH * 256;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const DropState ds = drop_init(dc);
  const long long rows = (long long)N * S;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
    const long long n = row / S;
    const int s = (int)(row - n * S);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = c * 256 + lane * 8;
      float v[8];
      if (s == 0) {
        const float4 a = *reinterpret_cast<const float4*>(cls + col), b = *reinterpret_cast<const float4*>(cls + col + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
        const uint4 w = *reinterpret_cast<const uint4*>(t0 + (n * (S - 1) + (s - 1)) * H + col);
        const float2 f0 = unpack_bf16(w.x), f1 = unpack_bf16(w.y), f2 = unpack_bf16(w.z), f3 = unpack_bf16(w.w);
        v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y; v[4] = f2.x; v[5] = f2.y; v[6] = f3.x; v[7] = f3.y;
      }
      const float4 pa = *reinterpret_cast<const float4*>(pos + (long long)s * H + col), pb = *reinterpret_cast<const float4*>(pos + (long long)s * H + col + 4);
      v[0] += pa.x; v[1] += pa.y; v[2] += pa.z; v[3] += pa.w; v[4] += pb.x; v[5] += pb.y; v[6] += pb.z; v[7] += pb.w;
      if (ds.on) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= drop_mult(ds, (unsigned long long)row * H + col + j);
      }
      float* o32 = x32 + row * H + col;
      *reinterpret_cast<float4*>(o32) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(o32 + 4) = make_float4(v[4], v[5], v[6], v[7]);
      *reinterpret_cast<uint4*>(x16 + row * H + col) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
  }
}

template <int NCH>
__global__ void __launch_bounds__(256) vit_embed_bwd_kernel(const __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dfull,
                                                            __nv_bfloat16* __restrict__ dt0, int N, int S, DropCfg dc) {
  pdl_grid_sync();
  constexpr int H = NCH * 256;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const DropState ds = drop_init(dc);
  const long long rows = (long long)N * S;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
    const long long n = row / S;
    const int s = (int)(row - n * S);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = c * 256 + lane * 8;
      uint4 w = *reinterpret_cast<const uint4*>(dx + row * H + col);
      if (ds.on) {
        const float2 f0 = unpack_bf16(w.x), f1 = unpack_bf16(w.y), f2 = unpack_bf16(w.z), f3 = unpack_bf16(w.w);
        float v[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= drop_mult(ds, (unsigned long long)row * H + col + j);
        w = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      }
      *reinterpret_cast<uint4*>(dfull + row * H + col) = w;
      if (s > 0) *reinterpret_cast<uint4*>(dt0 + (n * (S - 1) + (s - 1)) * H + col) = w;
    }
  }
}

static int rows_grid(long long rows) {
  long long b = (rows + 7) / 8;
  return (int)(b < 148 * 8 ? (b < 1 ? 1 : b) : 148 * 8);
}

int vit_embed_fwd(const void* t0, const float* cls, const float* pos, float* x32, void* x16, int N, int S, int H, DropArgs drop, cudaStream_t st) {
  HAMT_REQUIRE(H == 512 || H == 768 || H == 1024, "vit_embed_fwd: hidden size must be 512/768/1024");
  HAMT_REQUIRE((((uintptr_t)t0 | (uintptr_t)cls | (uintptr_t)pos | (uintptr_t)x32 | (uintptr_t)x16) & 15) == 0, "vit_embed_fwd: 16-byte alignment");
  if (N <= 0) return 0;
  DropCfg dc{drop.seed_ptr, drop.site, drop.p};
  const int grid = rows_grid((long long)N * S);
  auto T = (const __nv_bfloat16*)t0; auto X = (__nv_bfloat16*)x16;
  if (H == 768) launch_pdl(vit_embed_fwd_kernel<3>, grid, 256, 0, st, T, cls, pos, x32, X, N, S, dc);
  else if (H == 512) launch_pdl(vit_embed_fwd_kernel<2>, grid, 256, 0, st, T, cls, pos, x32, X, N, S, dc);
  else launch_pdl(vit_embed_fwd_kernel<4>, grid, 256, 0, st, T, cls, pos, x32, X, N, S, dc);
  return check_launch("vit_embed_fwd_kernel");
}

int vit_embed_bwd(const void* dx, void* dfull, void* dt0, int N, int S, int H, DropArgs drop, cudaStream_t st) {
  HAMT_REQUIRE(H == 512 || H == 768 || H == 1024, "vit_embed_bwd: hidden size must be 512/768/1024");
  HAMT_REQUIRE((((uintptr_t)dx | (uintptr_t)dfull | (uintptr_t)dt0) & 15) == 0, "vit_embed_bwd: 16-byte alignment");
  if (N <= 0) return 0;
  DropCfg dc{drop.seed_ptr, drop.site, drop.p};
  const int grid = rows_grid((long long)N * S);
  auto DX = (const __nv_bfloat16*)dx; auto DF = (__nv_bfloat16*)dfull; auto DT = (__nv_bfloat16*)dt0;
  if (H == 768) launch_pdl(vit_embed_bwd_kernel<3>, grid, 256, 0, st, DX, DF, DT, N, S, dc);
  else if (H == 512) launch_pdl(vit_embed_bwd_kernel<2>, grid, 256, 0, st, DX, DF, DT, N, S, dc);
  else launch_pdl(vit_embed_bwd_kernel<4>, grid, 256, 0, st, DX, DF, DT, N, S, dc);
  return check_launch("vit_embed_bwd_kernel");
}

}  // namespace hamt
