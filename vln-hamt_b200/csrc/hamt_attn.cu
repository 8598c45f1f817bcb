// Fused multi-head attention forward / backward for the HAMT hot path (head_dim 64).
//
//   P = softmax(Q K^T * scale + mask)  -- divide-then-add order and additive -10000 masks exactly as
//   the reference (pretrain_src/model/vilmodel.py:106-116 self, :332-343 cross);  O = dropout(P) V.
//
// The problems are tiny (S = 36 pano views, 53..78 vision tokens, 80 text tokens):
// one CTA owns one (batch, head) pair, stages Q/K/V in shared memory once, and each warp runs 16
// query rows through tensor-core MMAs (bf16 inputs, fp32 accumulate) with an online softmax computed
// with warp shuffles; scores / probabilities never touch HBM (the eager path materialises
// [N,12,S,S] three times).  The backward recomputes P from the saved log-sum-exp.
// Q/K/V are read in place from the fused QKV projection output through (batch stride, row pitch).
#include <stdio.h>
#include "hamt_common.cuh"
#include "hamt_kernels.h"

namespace hamt {

static constexpr int D = 64;      // head dim
static constexpr int LDS = 72;    // smem row pitch in bf16 (144 B: 16-byte aligned rows, conflict-free ldmatrix)

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

struct AttnP {
  const __nv_bfloat16 *q, *k, *v;
  long long q_bs, kv_bs, ldq, ldkv;
  const float* mask;
  __nv_bfloat16* out; long long ldo, o_bs;
  float* lse;
  int B, heads, Sq, Sk;
  float scale;
  DropCfg drop;
  // backward only
  const __nv_bfloat16* dout; long long lddo, do_bs;
  __nv_bfloat16 *dq, *dk, *dv;
  float *dbq, *dbk, *dbv;   // [heads * 64] fp32 or null: += column sums of the (bf16-rounded) dq / dk / dv  (bias gradients of the projections)
};

// cooperative ASYNC copy (cp.async, 16 B per request, no register round trip: every request of the CTA is in flight at once)
// of `rows` x 64 bf16 (row pitch ld) into smem [rows_pad][LDS]; padding rows are zero-filled (src-size 0).
__device__ __forceinline__ void stage_rows(__nv_bfloat16* dst, const __nv_bfloat16* src, long long ld, int rows, int rows_pad) {
  const uint32_t base = smem_u32(dst);
  for (int i = threadIdx.x; i < rows_pad * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    const bool valid = r < rows;
    const __nv_bfloat16* g = valid ? src + (long long)r * ld + c : src;
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(base + 2u * (uint32_t)(r * LDS + c)), "l"(g), "r"(sz) : "memory");
  }
}
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__global__ void __launch_bounds__(256, 2) attn_fwd_kernel(const AttnP p) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t smem[];
  // keys are consumed in blocks of 64 (Sk_pad) but only the 16-key groups that hold real keys are ever read from shared memory,
  // so K / V are staged with 16-row granularity (Sk_rows): S = 80 -> 35 KB instead of 48 KB per CTA, 6 instead of 4 CTAs per SM
  const int Sq_pad = (p.Sq + 15) & ~15, Sk_pad = (p.Sk + 63) & ~63, Sk_rows = (p.Sk + 15) & ~15;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sK = sQ + Sq_pad * LDS;
  __nv_bfloat16* sV = sK + Sk_rows * LDS;
  float* sMask = reinterpret_cast<float*>(sV + Sk_rows * LDS);
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  stage_rows(sQ, p.q + b * p.q_bs + h * D, p.ldq, p.Sq, Sq_pad);
  stage_rows(sK, p.k + b * p.kv_bs + h * D, p.ldkv, p.Sk, Sk_rows);
  stage_rows(sV, p.v + b * p.kv_bs + h * D, p.ldkv, p.Sk, Sk_rows);
  for (int j = threadIdx.x; j < Sk_pad; j += blockDim.x)
    sMask[j] = j < p.Sk ? (p.mask ? p.mask[(long long)b * p.Sk + j] : 0.f) : -INFINITY;
  stage_wait();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const AttnDrop ds = attn_drop_init(p.drop);
  const uint32_t sQ_a = smem_u32(sQ), sK_a = smem_u32(sK), sV_a = smem_u32(sV);

  for (int qb = warp; qb < Sq_pad / 16; qb += nwarps) {
    uint32_t aq[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      ldsm_x4(aq[ks], sQ_a + 2u * ((qb * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LDS + ks * 16 + 8 * (lane >> 4)));
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    const int row0 = qb * 16 + g;   // rows row0 and row0 + 8

    for (int kb = 0; kb < Sk_pad / 64; ++kb) {
      const int ng = min(4, (p.Sk - kb * 64 + 15) >> 4);   // 16-key groups of this block that hold real keys (warp-uniform)
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          if (np >= ng) continue;
          uint32_t bk[4];
          ldsm_x4(bk, sK_a + 2u * ((kb * 64 + np * 16 + (lane & 7) + 8 * (lane >> 4)) * LDS + ks * 16 + 8 * ((lane >> 3) & 1)));
          mma16816(s[2 * np], aq[ks], bk[0], bk[1]);
          mma16816(s[2 * np + 1], aq[ks], bk[2], bk[3]);
        }
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = kb * 64 + nt * 8 + 2 * t + (e & 1);
          const float v = s[nt][e] * p.scale + sMask[key];   // padded keys carry -inf
          s[nt][e] = v;
          mx[e >> 1] = fmaxf(mx[e >> 1], v);
        }
      float corr[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float mn = fmaxf(m[r], quad_max(mx[r]));
        corr[r] = __expf(m[r] - mn);     // m = -inf on the first block -> 0
        m[r] = mn;
        l[r] *= corr[r];
      }
      float rs[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pe = __expf(s[nt][e] - m[e >> 1]);
          rs[e >> 1] += pe;
          float pd = pe;
          if (ds.on) {
            const int key = kb * 64 + nt * 8 + 2 * t + (e & 1);
            const int row = row0 + 8 * (e >> 1);
            pd *= attn_drop_mult(ds, attn_drop_rowkey(ds, (unsigned long long)blockIdx.x * p.Sq + row), key);
          }
          s[nt][e] = pd;
        }
      l[0] += quad_sum(rs[0]);
      l[1] += quad_sum(rs[1]);
#pragma unroll
      for (int i = 0; i < 8; ++i) { o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1]; }
      // O += P_drop (16 x 64 keys) * V (64 keys x 64 d)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (kk >= ng) continue;          // probabilities of fully padded groups are exactly zero
        uint32_t ap[4];
        ap[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
        ap[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
        ap[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        ap[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t bv[4];
          ldsm_x4_t(bv, sV_a + 2u * ((kb * 64 + kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LDS + dp * 16 + 8 * (lane >> 4)));
          mma16816(o[2 * dp], ap, bv[0], bv[1]);
          mma16816(o[2 * dp + 1], ap, bv[2], bv[3]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = row0 + 8 * r;
      if (row < p.Sq) {
        const float inv = 1.0f / l[r];
        __nv_bfloat16* op = p.out + b * p.o_bs + (long long)row * p.ldo + h * D;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
          *reinterpret_cast<uint32_t*>(op + nt * 8 + 2 * t) = pack_bf16(o[nt][2 * r] * inv, o[nt][2 * r + 1] * inv);
        if (t == 0 && p.lse) p.lse[((long long)b * p.heads + h) * p.Sq + row] = m[r] + __logf(l[r]);
      }
    }
  }
}

// Column sums of a 16 x 64 fp32 accumulator fragment (rows g, g+8 of every lane quad; columns nt*8 + 2t + {0,1}) added to 64
// shared-memory accumulators.  The 16 per-thread partial sums are reduced over the 8 row groups (lane bits 2..4) with a halving
// butterfly -- each step exchanges the half of the values the partner keeps -- 14 shuffles instead of 48; every lane ends up owning 2
// distinct columns.  (fp32 sums of the unrounded accumulators: the separate pass this replaces summed the bf16-rounded values; the
// difference is rounding noise.)
__device__ __forceinline__ void frag_colsum(const float (&acc)[8][4], float* sdst, int lane) {
  float v[16];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { v[2 * nt] = acc[nt][0] + acc[nt][2]; v[2 * nt + 1] = acc[nt][1] + acc[nt][3]; }
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
  float w[8], x[4], y[2];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = (h16 ? v[8 + i] : v[i]) + __shfl_xor_sync(0xffffffffu, h16 ? v[i] : v[8 + i], 16);
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = (h8 ? w[4 + i] : w[i]) + __shfl_xor_sync(0xffffffffu, h8 ? w[i] : w[4 + i], 8);
#pragma unroll
  for (int i = 0; i < 2; ++i) y[i] = (h4 ? x[2 + i] : x[i]) + __shfl_xor_sync(0xffffffffu, h4 ? x[i] : x[2 + i], 4);
  const int k = (h16 ? 8 : 0) + (h8 ? 4 : 0) + (h4 ? 2 : 0);    // index of y[0] in v: nt = k >> 1, y[1] is the odd column of the same pair
  float* d = sdst + (k >> 1) * 8 + 2 * (lane & 3);
  atomicAdd(d, y[0]);
  atomicAdd(d + 1, y[1]);
}

// ---------------------------------------------------------------------------------------------
// backward.  Phase 1: warp w owns keys [16w,16w+16): dK, dV in registers, dS^T -> smem.
//            Phase 2: warp w owns 16 queries: dQ = dS K.
// ---------------------------------------------------------------------------------------------
// LONG = false: Sq, Sk <= 128 (every R2R / R4R shape): dS goes through shared memory for the dQ pass, two CTAs per SM.
// LONG = true : long sequences (RxR instructions, L = 300): Q, dO, K, V of the (batch, head) pair fill shared memory, so the dQ
//               pass RECOMPUTES S and dP from registers (7 instead of 5 small matmuls) and needs no dS buffer; one CTA per SM.
template <bool LONG>
__global__ void __launch_bounds__(LONG ? 384 : 256, LONG ? 1 : 2) attn_bwd_kernel(const AttnP p) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t smem[];
  const int Sq_pad = (p.Sq + 15) & ~15, Sk_pad = (p.Sk + 15) & ~15;
  const int LDP = Sk_pad + 8;                       // dS row pitch (bf16); (Sk_pad+8)*2 B is a multiple of 16
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sDO = sQ + Sq_pad * LDS;
  __nv_bfloat16* sK = sDO + Sq_pad * LDS;
  __nv_bfloat16* sV = sK + Sk_pad * LDS;
  __nv_bfloat16* sDS = sV + Sk_pad * LDS;           // [Sq_pad][LDP]  (absent when LONG)
  float* sMask = reinterpret_cast<float*>(sDS + (LONG ? 0 : Sq_pad * LDP));
  float* sLse = sMask + Sk_pad;
  float* sDelta = sLse + Sq_pad;
  float* sBias = sDelta + Sq_pad;                   // [3][64] column sums of dq, dk, dv of this (batch, head)
  const bool want_bias = p.dbq != nullptr;
  if (want_bias)
    for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) sBias[i] = 0.f;
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  stage_rows(sQ, p.q + b * p.q_bs + h * D, p.ldq, p.Sq, Sq_pad);
  stage_rows(sDO, p.dout + b * p.do_bs + h * D, p.lddo, p.Sq, Sq_pad);
  stage_rows(sK, p.k + b * p.kv_bs + h * D, p.ldkv, p.Sk, Sk_pad);
  stage_rows(sV, p.v + b * p.kv_bs + h * D, p.ldkv, p.Sk, Sk_pad);
  for (int j = threadIdx.x; j < Sk_pad; j += blockDim.x) sMask[j] = (j < p.Sk && p.mask) ? p.mask[(long long)b * p.Sk + j] : 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // delta_i = sum_d dO[i,d] * O[i,d]  (O = saved forward output).  Item = (row, 16-byte chunk): the O chunks are requested
  // from global memory while the cp.async staging is still in flight; at most 4 items per thread (blockDim >= Sq_pad * 2).
  constexpr int kItems = LONG ? 1 : 4;
  uint4 o_reg[kItems];
  if (!LONG) {
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
      const int item = threadIdx.x + k * blockDim.x, r = item >> 3, c = (item & 7) * 8;
      o_reg[k] = make_uint4(0, 0, 0, 0);
      if (item < Sq_pad * 8 && r < p.Sq) o_reg[k] = *reinterpret_cast<const uint4*>(p.out + b * p.o_bs + (long long)r * p.ldo + h * D + c);
    }
  }
  for (int i = threadIdx.x; i < Sq_pad; i += blockDim.x) sLse[i] = i < p.Sq ? p.lse[((long long)b * p.heads + h) * p.Sq + i] : 0.f;
  stage_wait();
  __syncthreads();
  const int n_iter = LONG ? (Sq_pad * 8 + (int)blockDim.x - 1) / (int)blockDim.x : kItems;
#pragma unroll
  for (int k = 0; k < n_iter; ++k) {
    const int item = threadIdx.x + k * blockDim.x, r = item >> 3, c = (item & 7) * 8;
    if (item < Sq_pad * 8) {                     // whole warps are in or out (Sq_pad * 8 and blockDim are multiples of 32)
      if (LONG) {
        o_reg[0] = make_uint4(0, 0, 0, 0);
        if (r < p.Sq) o_reg[0] = *reinterpret_cast<const uint4*>(p.out + b * p.o_bs + (long long)r * p.ldo + h * D + c);
      }
      const uint4 o4 = o_reg[LONG ? 0 : k];
      const uint4 d4 = *reinterpret_cast<const uint4*>(sDO + r * LDS + c);
      const uint32_t dw[4] = {d4.x, d4.y, d4.z, d4.w}, ow[4] = {o4.x, o4.y, o4.z, o4.w};
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 a = unpack_bf16(dw[q]), o = unpack_bf16(ow[q]);
        acc += a.x * o.x + a.y * o.y;
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if ((item & 7) == 0) sDelta[r] = acc;
    }
  }
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  const AttnDrop ds = attn_drop_init(p.drop);
  const uint32_t sQ_a = smem_u32(sQ), sDO_a = smem_u32(sDO), sK_a = smem_u32(sK), sV_a = smem_u32(sV), sDS_a = smem_u32(sDS);

  for (int kw = warp; kw < Sk_pad / 16; kw += nwarps) {
    uint32_t ak[4][4], av[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t off = 2u * ((kw * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LDS + ks * 16 + 8 * (lane >> 4));
      ldsm_x4(ak[ks], sK_a + off);
      ldsm_x4(av[ks], sV_a + off);
    }
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    const int key0 = kw * 16 + g;   // keys key0, key0 + 8
    for (int qb = 0; qb < Sq_pad / 16; ++qb) {
      float st[2][4], dpt[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) st[i][0] = st[i][1] = st[i][2] = st[i][3] = dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bq[4], bd[4];
        const uint32_t off = 2u * ((qb * 16 + (lane & 7) + 8 * (lane >> 4)) * LDS + ks * 16 + 8 * ((lane >> 3) & 1));
        ldsm_x4(bq, sQ_a + off);
        ldsm_x4(bd, sDO_a + off);
        mma16816(st[0], ak[ks], bq[0], bq[1]);
        mma16816(st[1], ak[ks], bq[2], bq[3]);
        mma16816(dpt[0], av[ks], bd[0], bd[1]);
        mma16816(dpt[1], av[ks], bd[2], bd[3]);
      }
      // element (nt, e): key = key0 + 8*(e>>1), query = qb*16 + nt*8 + 2t + (e&1)
      float pd[2][4], dsv[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = key0 + 8 * (e >> 1);
          const int qi = qb * 16 + nt * 8 + 2 * t + (e & 1);
          float pe = 0.f, mult = 1.f;
          if (key < p.Sk && qi < p.Sq) {
            pe = __expf(st[nt][e] * p.scale + sMask[key] - sLse[qi]);
            if (ds.on) mult = attn_drop_mult(ds, attn_drop_rowkey(ds, (unsigned long long)blockIdx.x * p.Sq + qi), key);
          }
          pd[nt][e] = pe * mult;
          dsv[nt][e] = pe * (dpt[nt][e] * mult - sDelta[qi]) * p.scale;
          if (!LONG) sDS[qi * LDP + key] = __float2bfloat16_rn(dsv[nt][e]);
        }
      uint32_t ap[4], as_[4];
      ap[0] = pack_bf16(pd[0][0], pd[0][1]); ap[1] = pack_bf16(pd[0][2], pd[0][3]);
      ap[2] = pack_bf16(pd[1][0], pd[1][1]); ap[3] = pack_bf16(pd[1][2], pd[1][3]);
      as_[0] = pack_bf16(dsv[0][0], dsv[0][1]); as_[1] = pack_bf16(dsv[0][2], dsv[0][3]);
      as_[2] = pack_bf16(dsv[1][0], dsv[1][1]); as_[3] = pack_bf16(dsv[1][2], dsv[1][3]);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t bdo[4], bqq[4];
        const uint32_t off = 2u * ((qb * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LDS + dp * 16 + 8 * (lane >> 4));
        ldsm_x4_t(bdo, sDO_a + off);
        ldsm_x4_t(bqq, sQ_a + off);
        mma16816(dv[2 * dp], ap, bdo[0], bdo[1]);
        mma16816(dv[2 * dp + 1], ap, bdo[2], bdo[3]);
        mma16816(dk[2 * dp], as_, bqq[0], bqq[1]);
        mma16816(dk[2 * dp + 1], as_, bqq[2], bqq[3]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int key = key0 + 8 * r;
      if (key < p.Sk) {
        __nv_bfloat16* kp = p.dk + b * p.kv_bs + (long long)key * p.ldkv + h * D;
        __nv_bfloat16* vp = p.dv + b * p.kv_bs + (long long)key * p.ldkv + h * D;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          *reinterpret_cast<uint32_t*>(kp + nt * 8 + 2 * t) = pack_bf16(dk[nt][2 * r], dk[nt][2 * r + 1]);
          *reinterpret_cast<uint32_t*>(vp + nt * 8 + 2 * t) = pack_bf16(dv[nt][2 * r], dv[nt][2 * r + 1]);
        }
      }
    }
    if (want_bias) {      // rows of padded keys are exactly zero
      frag_colsum(dk, sBias + D, lane);
      frag_colsum(dv, sBias + 2 * D, lane);
    }
  }
  __syncthreads();
  // phase 2: dQ[16 x 64] = dS[16 x Sk] K[Sk x 64]
  for (int qb = warp; qb < Sq_pad / 16; qb += nwarps) {
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    if (!LONG) {
      for (int kk = 0; kk < Sk_pad / 16; ++kk) {
        uint32_t a[4];
        ldsm_x4(a, sDS_a + 2u * ((qb * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LDP + kk * 16 + 8 * (lane >> 4)));
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t bk[4];
          ldsm_x4_t(bk, sK_a + 2u * ((kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LDS + dp * 16 + 8 * (lane >> 4)));
          mma16816(dq[2 * dp], a, bk[0], bk[1]);
          mma16816(dq[2 * dp + 1], a, bk[2], bk[3]);
        }
      }
    } else {
      // recompute S = Q K^T and dP = dO V^T for this warp's 16 queries, 16 keys at a time (operand layouts as in the forward)
      uint32_t aq[4][4], ad[4][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t off = 2u * ((qb * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LDS + ks * 16 + 8 * (lane >> 4));
        ldsm_x4(aq[ks], sQ_a + off);
        ldsm_x4(ad[ks], sDO_a + off);
      }
      const int row0 = qb * 16 + g;       // rows row0, row0 + 8
      const float lse_r[2] = {sLse[row0], sLse[row0 + 8]}, del_r[2] = {sDelta[row0], sDelta[row0 + 8]};
      for (int kg = 0; kg < Sk_pad / 16; ++kg) {
        float sc[2][4], dp_[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i) sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = dp_[i][0] = dp_[i][1] = dp_[i][2] = dp_[i][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t bk[4], bv[4];
          const uint32_t off = 2u * ((kg * 16 + (lane & 7) + 8 * (lane >> 4)) * LDS + ks * 16 + 8 * ((lane >> 3) & 1));
          ldsm_x4(bk, sK_a + off);
          ldsm_x4(bv, sV_a + off);
          mma16816(sc[0], aq[ks], bk[0], bk[1]);
          mma16816(sc[1], aq[ks], bk[2], bk[3]);
          mma16816(dp_[0], ad[ks], bv[0], bv[1]);
          mma16816(dp_[1], ad[ks], bv[2], bv[3]);
        }
        float dsv[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int key = kg * 16 + nt * 8 + 2 * t + (e & 1);
            const int qi = row0 + 8 * (e >> 1);
            float pe = 0.f, mult = 1.f;
            if (key < p.Sk && qi < p.Sq) {
              pe = __expf(sc[nt][e] * p.scale + sMask[key] - lse_r[e >> 1]);
              if (ds.on) mult = attn_drop_mult(ds, attn_drop_rowkey(ds, (unsigned long long)blockIdx.x * p.Sq + qi), key);
            }
            dsv[nt][e] = pe * (dp_[nt][e] * mult - del_r[e >> 1]) * p.scale;
          }
        uint32_t a[4];
        a[0] = pack_bf16(dsv[0][0], dsv[0][1]); a[1] = pack_bf16(dsv[0][2], dsv[0][3]);
        a[2] = pack_bf16(dsv[1][0], dsv[1][1]); a[3] = pack_bf16(dsv[1][2], dsv[1][3]);
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t bk[4];
          ldsm_x4_t(bk, sK_a + 2u * ((kg * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LDS + dp * 16 + 8 * (lane >> 4)));
          mma16816(dq[2 * dp], a, bk[0], bk[1]);
          mma16816(dq[2 * dp + 1], a, bk[2], bk[3]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = qb * 16 + g + 8 * r;
      if (row < p.Sq) {
        __nv_bfloat16* qp = p.dq + b * p.q_bs + (long long)row * p.ldq + h * D;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<uint32_t*>(qp + nt * 8 + 2 * t) = pack_bf16(dq[nt][2 * r], dq[nt][2 * r + 1]);
      }
    }
    if (want_bias) frag_colsum(dq, sBias, lane);     // rows of padded queries are exactly zero (dS = 0 there)
  }
  if (want_bias) {
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) {
      float* dst = i < D ? p.dbq : (i < 2 * D ? p.dbk : p.dbv);
      atomicAdd(dst + h * D + (i & (D - 1)), sBias[i]);
    }
  }
}

static int check_common(const AttnArgs& a) {
  HAMT_REQUIRE(a.B > 0 && a.heads > 0 && a.Sq > 0 && a.Sk > 0, "attn: empty problem");
  HAMT_REQUIRE((((uintptr_t)a.q | (uintptr_t)a.k | (uintptr_t)a.v | (uintptr_t)a.out) & 15) == 0, "attn: q/k/v/out must be 16-byte aligned");
  HAMT_REQUIRE(a.ldq % 8 == 0 && a.ldkv % 8 == 0 && a.ldo % 8 == 0 && a.q_bstride % 8 == 0 && a.kv_bstride % 8 == 0 && a.o_bstride % 8 == 0,
               "attn: pitches / batch strides must be multiples of 8 elements");
  return 0;
}
static AttnP to_params(const AttnArgs& a) {
  AttnP p{};
  p.q = (const __nv_bfloat16*)a.q; p.k = (const __nv_bfloat16*)a.k; p.v = (const __nv_bfloat16*)a.v;
  p.q_bs = a.q_bstride; p.kv_bs = a.kv_bstride; p.ldq = a.ldq; p.ldkv = a.ldkv;
  p.mask = a.mask; p.out = (__nv_bfloat16*)a.out; p.ldo = a.ldo; p.o_bs = a.o_bstride; p.lse = a.lse;
  p.B = a.B; p.heads = a.heads; p.Sq = a.Sq; p.Sk = a.Sk; p.scale = a.scale;
  p.drop = DropCfg{a.drop.seed_ptr, a.drop.site, a.drop.p};
  return p;
}

int attn_fwd(const AttnArgs& a, cudaStream_t st) {
  if (int rc = check_common(a)) return rc;
  {
    int rc = 0;
    if (attn_fwd_tc(a, st, &rc)) return rc;      // TMA + tcgen05 packed-tile kernel; shapes outside its envelope fall through
  }
  AttnP p = to_params(a);
  const int Sq_pad = (a.Sq + 15) & ~15, Sk_pad = (a.Sk + 63) & ~63, Sk_rows = (a.Sk + 15) & ~15;
  const size_t smem = (size_t)(Sq_pad + 2 * Sk_rows) * LDS * 2 + Sk_pad * 4;
  HAMT_REQUIRE(smem <= 227 * 1024, "attn_fwd: sequence too long for the single-CTA kernel");
  static bool attr_set = false;
  if (!attr_set) {
    // without the carveout hint the driver sized shared memory for ONE resident CTA (ncu: occupancy_limit_shared_mem = 1)
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); return -3; }
    attr_set = true;
  }
  int nw = Sq_pad / 16;
  if (nw > 8) nw = 8;
  launch_pdl(attn_fwd_kernel, a.B * a.heads, nw * 32, smem, st, p);
  return check_launch("attn_fwd_kernel");
}

int attn_bwd(const AttnBwdArgs& a, cudaStream_t st) {
  if (int rc = check_common(a.f)) return rc;
  HAMT_REQUIRE(a.f.lse != nullptr, "attn_bwd: lse required");
  HAMT_REQUIRE((((uintptr_t)a.dout | (uintptr_t)a.dq | (uintptr_t)a.dk | (uintptr_t)a.dv) & 15) == 0, "attn_bwd: grads must be 16-byte aligned");
  AttnP p = to_params(a.f);
  p.dout = (const __nv_bfloat16*)a.dout; p.lddo = a.lddo; p.do_bs = a.do_bstride;
  p.dq = (__nv_bfloat16*)a.dq; p.dk = (__nv_bfloat16*)a.dk; p.dv = (__nv_bfloat16*)a.dv;
  p.dbq = a.dbq; p.dbk = a.dbk; p.dbv = a.dbv;
  HAMT_REQUIRE((a.dbq == nullptr) == (a.dbk == nullptr) && (a.dbq == nullptr) == (a.dbv == nullptr), "attn_bwd: bias-gradient pointers come as a triple");
  {
    int rc = 0;
    if (attn_bwd_tc(a, st, &rc)) return rc;      // TMA + tcgen05 packed-tile kernel; shapes outside its envelope fall through
  }
  const int Sq_pad = (a.f.Sq + 15) & ~15, Sk_pad = (a.f.Sk + 15) & ~15;
  const bool is_long = Sq_pad > 128 || Sk_pad > 128;
  const size_t smem = (size_t)(2 * Sq_pad + 2 * Sk_pad) * LDS * 2 + (is_long ? 0 : (size_t)Sq_pad * (Sk_pad + 8) * 2) + (Sk_pad + 2 * Sq_pad + 3 * D) * 4;
  HAMT_REQUIRE(smem <= 227 * 1024, "attn_bwd: sequence too long for the single-CTA backward kernel (Sq + Sk <= ~790)");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); return -3; }
    attr_set = true;
  }
  int nw = (Sk_pad > Sq_pad ? Sk_pad : Sq_pad) / 16;
  if (nw > 8) nw = 8;
  // LONG: one CTA per SM (Q, dO, K, V of the pair fill shared memory): 12 warps, so the 19 key / query tiles of an RxR instruction
  // (L = 300) take two rounds per phase instead of three
  if (is_long) launch_pdl(attn_bwd_kernel<true>, a.f.B * a.f.heads, 384, smem, st, p);
  else launch_pdl(attn_bwd_kernel<false>, a.f.B * a.f.heads, nw * 32, smem, st, p);
  return check_launch("attn_bwd_kernel");
}

}  // namespace hamt
