// Blackwell-native fused attention for the HAMT hot path (head_dim 64): TMA-staged tiles, tcgen05.mma with the scores and the
// output accumulator in TMEM, several (sequence, head) problems PACKED into one 128-row UMMA tile.
//
//   P = softmax(Q K^T * scale + mask);  O = dropout(P) V        (pretrain_src/model/vilmodel.py:96-129 self, :322-349 cross)
//
// Why packing: the problems of this model are tiny -- 36 x 36 (the 11 520 panorama problems of a batch), 53 x 53, 80 x 53, 16 x 80 --
// so one problem fills a fraction of the 128-row tile the tensor core works on.  A tile therefore holds P problems:
//
//   rows (TMEM lanes, one softmax thread each)        columns of the score tile S = Qp Kp^T (fp32, TMEM)
//   Sq <= 32 : 4 problems, one per warp (32-row slot)   problem p owns the key window [p*W, p*W + Sk),  W = Sk rounded up to 8
//   Sq <= 40 : 3 problems, rows 0..31 of problem p in   (only the diagonal blocks of S are meaningful; the off-diagonal products
//              warp p, the Sq-32 remaining rows of all   cost tensor-pipe time that is idle anyway: the kernel is bound by HBM
//              three in warp 3 (lanes 8p .. 8p+7)         traffic and by the softmax ALU work, not by MMA issue)
//   Sq <= 64 : 2 problems at rows 0 and 64
//   else     : 1 problem per tile, Sq > 128 -> several 128-row tiles per problem (K / V re-read through L2)
//
// O = Pd V is ONE MMA over the packed key axis with a block-diagonal Pd: every softmax thread writes its whole row of Pd (bf16, K-major,
// 128-byte swizzle) with zeros outside its own key window.
//
// Warp roles (384 threads, one persistent CTA per SM, static round-robin tile list):
//   warp 0      TMA producer: Q / K / V boxes of the tile's problems straight out of the fused [tokens, 3H] projection buffer through
//               3-D tensor maps (column, position, sequence) -- rows past the end of a sequence are zero-filled by the TMA unit, so the
//               padding of the packed tile needs no extra pass -- into a ring of NS shared-memory stages;
//   warp 1      TMEM owner and single-thread tcgen05.mma issuer: S = Q K^T as soon as a stage lands, O = Pd V as soon as the softmax
//               warps publish Pd; order QK(i+1) before PV(i) so the tensor pipe never waits for the softmax of the same tile;
//   warps 4-7 / 8-11  two softmax warpgroups working on alternate tiles (each has its own S / O TMEM columns and Pd buffer): tcgen05.ld of the
//               lane's key window -> scale + additive mask (divide-then-add, -10000 masks as in the reference) -> max / exp2 / sum with
//               one thread per row (no shuffles) -> dropout on the probabilities -> Pd to shared memory -> later the O tile: tcgen05.ld,
//               1/l scaling, bf16, swizzled staging, ONE TMA store per warp (rows past the sequence end are clipped by the TMA unit).
// The log-sum-exp of every row is saved for the backward kernel (attn_bwd_tc_kernel below).
#include <cuda.h>
#include <stdio.h>
#include "hamt_common.cuh"
#include "hamt_kernels.h"

namespace hamt {

namespace tc {

static constexpr int kThreads = 384;     // forward: warps 0..3: TMA producer, MMA issuer, 2 spare; warps 4..7 and 8..11: the two softmax warpgroups
static constexpr int kBwdSoftmaxWarps = 16;                           // backward: warps 4..19 softmax backward (4 per TMEM lane quarter), 20..23 epilogue
static constexpr int kBwdThreads = (4 + kBwdSoftmaxWarps + 4) * 32;   // 768

__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const void* tmap, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(tmap), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes (st.shared) -> visible to the async proxy (UMMA operand reads, TMA stores)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// Columns [0, NU * 8) of a TMEM row window into v: full 32-column chunks, then a tail of 8 / 16 / 24 columns.  The caller issues
// tmem_ld_wait() before reading v.
template <int NU>
__device__ __forceinline__ void load_window(uint32_t taddr, uint32_t (&v)[NU * 8]) {
  constexpr int W = NU * 8;
#pragma unroll
  for (int c = 0; c * 32 < W; ++c) {
    constexpr int dummy = 0; (void)dummy;
    const int rem = W - c * 32;          // compile-time after unrolling
    if (rem >= 32) tmem_ld_x32(taddr + c * 32, &v[c * 32]);
    else {
      if (rem >= 16) {
        tmem_ld_x16(taddr + c * 32, &v[c * 32]);
        if (rem >= 24) tmem_ld_x8(taddr + c * 32 + 16, &v[c * 32 + 16]);
      } else if (rem >= 8) tmem_ld_x8(taddr + c * 32, &v[c * 32]);
    }
  }
}

// tile geometry shared by the forward and backward kernels (filled on the host by plan())
struct Geom {
  int nprob, heads, Sq, Sk;
  int W;          // key-window stride of a problem inside the packed key axis (Sk rounded up to 8)
  int P;          // problems per tile
  int regime;     // 0: 32-row slots (Sq <= 32), 1: 3 problems + shared remainder warp (Sq <= 40), 2: 64-row slots, 3: one problem per tile
  int QT;         // 128-row query tiles per problem (regime 3)
  int ntiles;
  int n_total;    // packed key extent of the MMAs (P * W rounded up to 16)
  int nwin;       // key windows the remainder warp of regime 1 has to look at (= P), else 1
};

struct FwdParams {
  Geom g;
  int ns;                 // input stages
  int nwg;                // softmax warpgroups in use (2, or 1 when the score tile needs more than half of TMEM)
  uint32_t stage_bytes, off_k, off_v;     // per input stage: Q at 0, K at off_k, V at off_v
  uint32_t off_pd;                        // Pd inside the stage: on top of the Q / K tiles (dead after S = Q K^T)
  uint32_t off_stg;                       // output staging (16 KB per warpgroup), from the start of dynamic smem
  uint32_t off_mask, mask_floats;         // per softmax warp: nwin * NCH*32 floats
  uint32_t off_bar;
  uint32_t tx_q32, tx_q8, tx_kv;          // bytes of one Q box (32 rows), one remainder box (8 rows), one K or V box
  int kv_box_rows;
  float scale_log2;       // softmax scale * log2(e)
  const float* mask;      // additive fp32 [B, Sk] or null
  float* lse;             // fp32 [B, heads, Sq] or null
  DropCfg drop;
};

// lane -> (problem slot inside the tile, query row of that problem); slot = TMEM lane quarter owned by the warp
struct RowMap {
  int pi;      // problem slot (0 .. P-1)
  int qrow;    // query position inside the problem
  bool ok;     // the lane maps to a row of the layout at all
};
__device__ __forceinline__ RowMap row_map(const Geom& g, int slot, int lane, int qt) {
  RowMap r;
  if (g.regime == 0) { r.pi = slot; r.qrow = lane; r.ok = lane < g.Sq && slot < g.P; }
  else if (g.regime == 1) {
    if (slot < 3) { r.pi = slot; r.qrow = lane; r.ok = true; }
    else { r.pi = lane >> 3; r.qrow = 32 + (lane & 7); r.ok = r.pi < 3 && r.qrow < g.Sq; }
  } else if (g.regime == 2) { r.pi = slot >> 1; r.qrow = (slot & 1) * 32 + lane; r.ok = r.qrow < g.Sq; }
  else { r.pi = 0; r.qrow = qt * 128 + slot * 32 + lane; r.ok = r.qrow < g.Sq; }
  return r;
}

// ------------------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------------------
// NU = key-window width in groups of 8 keys (W = 8 NU); MULTI = the 3-problem layout whose fourth warp holds rows of all three problems.
// Both are compile-time so that every instantiation carries only its own straight-line softmax: the first version (runtime widths,
// both layouts in one kernel) was 4.5 k SASS instructions and lost 28 % of its issue slots to instruction fetch
// (profiles/r02_ncu_attn_fwd_v1.txt).
template <int NU, bool MULTI>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_qr, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_or,
                   const FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const Geom& g = p.g;
  constexpr int NCH = (NU + 3) / 4;           // 32-column chunks of a key window (mask staging granularity)
  constexpr int W = NU * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_base = smem_base + p.off_bar;
  // barriers: full[ns], empty[ns], s_full[2], p_full[2], o_full[2], o_empty[2], tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.ns + s); };
  auto sfull_bar = [&](int w) { return bar_base + 8u * (2 * p.ns + w); };
  auto pfull_bar = [&](int w) { return bar_base + 8u * (2 * p.ns + 2 + w); };
  auto ofull_bar = [&](int w) { return bar_base + 8u * (2 * p.ns + 4 + w); };
  auto oempty_bar = [&](int w) { return bar_base + 8u * (2 * p.ns + 6 + w); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * p.ns + 8);

  if (warp == 0 && lane == 0) {
    if (smem_base & 1023u) { printf("hamt attn: dynamic smem base not 1024-byte aligned\n"); __trap(); }
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_o);
    if (g.regime == 1) { tma_prefetch_desc(&tm_qr); tma_prefetch_desc(&tm_or); }
    for (int s = 0; s < p.ns; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int w = 0; w < 2; ++w) { mbar_init(sfull_bar(w), 1); mbar_init(pfull_bar(w), 4); mbar_init(ofull_bar(w), 1); mbar_init(oempty_bar(w), 4); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr_addr);
  // Rows of the packed key axis that no TMA box ever writes (alignment padding behind the last problem; whole problem windows when
  // the last tile is partial) must hold finite values: they meet zero probabilities in the PV product.
  if (g.nprob % g.P != 0) {
    for (uint32_t i = threadIdx.x * 16u; i < (uint32_t)p.ns * p.stage_bytes; i += kThreads * 16u)
      *reinterpret_cast<uint4*>(smem_raw + i) = make_uint4(0, 0, 0, 0);
  } else {
    const uint32_t r0 = (uint32_t)(g.P * W) * 128u, r1 = (uint32_t)g.n_total * 128u;       // byte range inside a V tile
    for (int s = 0; s < p.ns; ++s)
      for (uint32_t i = r0 + threadIdx.x * 16u; i < r1; i += kThreads * 16u)
        *reinterpret_cast<uint4*>(smem_raw + (uint32_t)s * p.stage_bytes + p.off_v + i) = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_grid_sync();   // everything above overlapped the previous kernel
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  const int heads = g.heads;
  const int n_my = ((int)blockIdx.x < g.ntiles) ? (g.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int kch = (g.n_total + 63) >> 6;         // 64-key chunks of Pd
  // TMEM columns: warpgroup w: S at w*256, O at w*256 + 192 (two warpgroups, n_total <= 192); single warpgroup: S at 0, O at 448
  auto s_col = [&](int w) { return (uint32_t)(p.nwg == 2 ? w * 256 : 0); };
  auto o_col = [&](int w) { return (uint32_t)(p.nwg == 2 ? w * 256 + 192 : 448); };

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      for (int i = 0; i < n_my; ++i) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int s = i % p.ns;
        const uint32_t ph = (uint32_t)(i / p.ns) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const int grp = tile / g.QT, qt = tile % g.QT;
        const int p0 = grp * g.P;
        const int np = min(g.P, g.nprob - p0);
        const uint32_t sq = smem_base + (uint32_t)s * p.stage_bytes, sk = sq + p.off_k, sv = sq + p.off_v;
        // byte count first (expect_tx), then the copies
        uint32_t bytes = 0;
        for (int pi = 0; pi < np; ++pi) {
          if (g.regime == 0) bytes += p.tx_q32;
          else if (g.regime == 1) bytes += p.tx_q32 + p.tx_q8;
          else if (g.regime == 2) bytes += 2 * p.tx_q32;
          else for (int k = 0; k < 4; ++k) if (qt * 128 + k * 32 < g.Sq) bytes += p.tx_q32;
          bytes += 2 * p.tx_kv;
        }
        mbar_expect_tx(full_bar(s), bytes);
        for (int pi = 0; pi < np; ++pi) {
          const int pr = p0 + pi, b = pr / heads, h = pr % heads;
          if (g.regime == 0) tma_load_3d(sq + pi * 4096u, &tm_q, h * 64, 0, b, full_bar(s));
          else if (g.regime == 1) {
            tma_load_3d(sq + pi * 4096u, &tm_q, h * 64, 0, b, full_bar(s));
            tma_load_3d(sq + 96u * 128u + pi * 1024u, &tm_qr, h * 64, 32, b, full_bar(s));
          } else if (g.regime == 2) {
            tma_load_3d(sq + pi * 8192u, &tm_q, h * 64, 0, b, full_bar(s));
            tma_load_3d(sq + pi * 8192u + 4096u, &tm_q, h * 64, 32, b, full_bar(s));
          } else {
            for (int k = 0; k < 4; ++k)
              if (qt * 128 + k * 32 < g.Sq) tma_load_3d(sq + k * 4096u, &tm_q, h * 64, qt * 128 + k * 32, b, full_bar(s));
          }
          tma_load_3d(sk + (uint32_t)(pi * g.W) * 128u, &tm_k, h * 64, 0, b, full_bar(s));
          tma_load_3d(sv + (uint32_t)(pi * g.W) * 128u, &tm_v, h * 64, 0, b, full_bar(s));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      // Two kinds of work, issued in whatever order their inputs become ready (non-blocking mbarrier tests): S = Q K^T of the next
      // tile as soon as its stage has landed, O = Pd V as soon as the softmax warps have published Pd.  (A fixed issue order makes
      // PV(i) wait for the TMA of tile i + 1 and serialises the softmax warpgroups on memory latency.)
      const uint32_t idesc_s = umma_idesc_bf16(128, g.n_total, false, false);     // S = Q (K-major) x K^T (K-major)
      const uint32_t idesc_o = umma_idesc_bf16(128, 64, false, true);             // O = Pd (K-major) x V (MN-major: d contiguous)
      int qk_i = 0, pv_j = 0;
      unsigned long long t0 = 0;
      uint32_t idle = 0;
      while (pv_j < n_my) {
        bool progressed = false;
        // S columns of warpgroup w are free once Pd of its previous tile was published (observed before PV of that tile was issued)
        if (qk_i < n_my && qk_i < pv_j + p.nwg) {
          const int w = qk_i % p.nwg, st = qk_i % p.ns;
          if (mbar_test_wait(full_bar(st), (uint32_t)(qk_i / p.ns) & 1u)) {
            tc_fence_after();
            const uint32_t sq = smem_base + (uint32_t)st * p.stage_bytes, sk = sq + p.off_k;
            const uint32_t d_tmem = tmem_base + s_col(w);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = umma_smem_desc(sq + k * 32, 16, 1024);
              const uint64_t db = umma_smem_desc(sk + k * 32, 16, 1024);
              umma_bf16(d_tmem, da, db, idesc_s, k > 0 ? 1u : 0u);
            }
            umma_commit(sfull_bar(w));
            ++qk_i;
            progressed = true;
          }
        }
        if (pv_j < qk_i) {
          const int w = pv_j % p.nwg, st = pv_j % p.ns;
          const uint32_t use = (uint32_t)(pv_j / p.nwg);       // tiles this warpgroup has handled before
          // Pd of tile pv_j is in shared memory (and its S has been read); the O accumulator of the warpgroup's previous tile is drained
          if (mbar_test_wait(pfull_bar(w), use & 1u) && mbar_test_wait(oempty_bar(w), (use & 1u) ^ 1u)) {
            tc_fence_after();
            const uint32_t sbase = smem_base + (uint32_t)st * p.stage_bytes;
            const uint32_t sv = sbase + p.off_v, sp = sbase + p.off_pd;
            const uint32_t d_tmem = tmem_base + o_col(w);
            const int ksteps = g.n_total >> 4;
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t da = umma_smem_desc(sp + (uint32_t)(k >> 2) * 16384u + (uint32_t)(k & 3) * 32u, 16, 1024);
              const uint64_t db = umma_smem_desc(sv + (uint32_t)k * 2048u, 8192, 1024);
              umma_bf16(d_tmem, da, db, idesc_o, k > 0 ? 1u : 0u);
            }
            umma_commit(ofull_bar(w));
            umma_commit(empty_bar(st));           // Q / K / V / Pd of this stage are no longer needed
            ++pv_j;
            progressed = true;
          }
        }
        if (progressed) idle = 0;
        else if ((++idle & 0xfffu) == 0) {        // bounded: a protocol bug traps instead of hanging the GPU
          if (t0 == 0) t0 = globaltimer_ns();
          else if (globaltimer_ns() - t0 > 4000000000ull) { printf("hamt attn: MMA issuer stalled (block %d qk %d pv %d of %d)\n", blockIdx.x, qk_i, pv_j, n_my); __trap(); }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax + output warps =====================
    const int wg = (warp - 4) >> 2;              // warpgroup 0: warps 4..7, warpgroup 1: warps 8..11
    const int slot = warp & 3;                   // TMEM lane quarter this warp may access = 32-row slot of the tile
    if (wg < p.nwg) {
      const AttnDrop ds = attn_drop_init(p.drop);
      const bool multi = MULTI && slot == 3;
      float* smask = reinterpret_cast<float*>(smem_raw + p.off_mask) + (uint32_t)((warp - 4) * p.mask_floats);
      uint8_t* stg = smem_raw + p.off_stg + (uint32_t)wg * 16384u;       // this warpgroup's output staging tile
      const uint32_t stg_a = smem_base + p.off_stg + (uint32_t)wg * 16384u;
      const int row = slot * 32 + lane;          // row of the tile = TMEM lane
      const uint32_t lane_field = (uint32_t)(slot * 32) << 16;
      int use = 0;
      for (int i = wg; i < n_my; i += p.nwg, ++use) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int grp = tile / g.QT, qt = tile % g.QT;
        const int p0 = grp * g.P;
        const RowMap rm = row_map(g, slot, lane, qt);
        const int pr = p0 + rm.pi;
        const bool valid = rm.ok && pr < g.nprob;
        const int b = valid ? pr / heads : 0, h = valid ? pr % heads : 0;
        // ---- additive mask rows of the key window(s) this warp looks at, in log2 units; padded keys carry -inf
        if (p.mask_floats != 0) {
          const int nw = multi ? g.nwin : 1;
          for (int wdx = 0; wdx < nw; ++wdx) {
            const int prw = multi ? p0 + wdx : __shfl_sync(0xffffffffu, pr, 0);
            const bool okw = multi ? (prw < g.nprob) : __shfl_sync(0xffffffffu, valid ? 1 : 0, 0) != 0;
            const int bw = okw ? prw / heads : 0;
            for (int j = lane; j < NCH * 32; j += 32) {
              float mv = -INFINITY;
              if (j < g.Sk) mv = (p.mask != nullptr && okw) ? p.mask[(long long)bw * g.Sk + j] * 1.4426950408889634f : 0.f;
              smask[wdx * NCH * 32 + j] = mv;
            }
          }
          __syncwarp();
        }
        mbar_wait(sfull_bar(wg), (uint32_t)use & 1u);
        tc_fence_after();
        // ---- scores of this lane's key window
        uint32_t sv[NU * 8];
        const uint32_t t_s = tmem_base + lane_field + s_col(wg);
        if (!MULTI || !multi) {
          // (warps whose slot holds no problem still take part in the warp-collective load: window 0)
          const int wpi = __shfl_sync(0xffffffffu, valid ? rm.pi : 0, 0);
          load_window<NU>(t_s + (uint32_t)(wpi * W), sv);
          tmem_ld_wait();
        } else if constexpr (MULTI) {
          // remainder warp: its lanes belong to different problems -> fetch the three windows 16 columns at a time and keep the lane's own
#pragma unroll
          for (int c = 0; c < NU * 8; c += 16) {
            uint32_t t0[16], t1[16], t2[16];
            if (c + 16 <= NU * 8) {
              tmem_ld_x16(t_s + 0 * W + c, t0); tmem_ld_x16(t_s + 1 * W + c, t1); tmem_ld_x16(t_s + 2 * W + c, t2);
            } else {
              tmem_ld_x8(t_s + 0 * W + c, t0); tmem_ld_x8(t_s + 1 * W + c, t1); tmem_ld_x8(t_s + 2 * W + c, t2);
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c + j < NU * 8) sv[c + j] = rm.pi == 0 ? t0[j] : (rm.pi == 1 ? t1[j] : t2[j]);
          }
        }
        float mx = -INFINITY;
        if (p.mask_floats != 0) {
          const float* mrow = smask + (multi ? rm.pi * NCH * 32 : 0);
#pragma unroll
          for (int kk = 0; kk < NU; ++kk) {
            const float4 m0 = *reinterpret_cast<const float4*>(mrow + kk * 8), m1 = *reinterpret_cast<const float4*>(mrow + kk * 8 + 4);
            const float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float t = fmaf(__uint_as_float(sv[kk * 8 + e]), p.scale_log2, mm[e]);      // padded keys: 0 * c + (-inf)
              sv[kk * 8 + e] = __float_as_uint(t);
              mx = fmaxf(mx, t);
            }
          }
        } else {        // no mask (the panorama encoder attends to all 36 views): only the padded keys of the last group are excluded
#pragma unroll
          for (int j = 0; j < NU * 8; ++j) {
            float t = __uint_as_float(sv[j]) * p.scale_log2;
            if (j >= (NU - 1) * 8 && j >= g.Sk) t = -INFINITY;
            sv[j] = __float_as_uint(t);
            mx = fmaxf(mx, t);
          }
        }
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int j = 0; j < NU * 8; j += 2) {
          const float pa = ex2_approx(__uint_as_float(sv[j]) - mx), pb = ex2_approx(__uint_as_float(sv[j + 1]) - mx);
          l0 += pa; l1 += pb;
          sv[j] = __float_as_uint(pa); sv[j + 1] = __float_as_uint(pb);
        }
        const float l = l0 + l1;
        if (ds.on) {          // dropout on the probabilities: one hash per pair of keys
          const uint32_t rowkey = attn_drop_rowkey(ds, (unsigned long long)pr * g.Sq + rm.qrow);
#pragma unroll
          for (int j = 0; j < NU * 8; j += 2) {
            const uint32_t bits = attn_drop_bits(rowkey, (uint32_t)(j >> 1));
            const float ma = (bits & 0xffffu) < ds.thresh16 ? 0.f : ds.scale, mb = (bits >> 16) < ds.thresh16 ? 0.f : ds.scale;
            sv[j] = __float_as_uint(__uint_as_float(sv[j]) * ma);
            sv[j + 1] = __float_as_uint(__uint_as_float(sv[j + 1]) * mb);
          }
        }
        // ---- this lane's row of Pd: zeros outside its key window (16-byte units of 8 keys; unit u of row r lives at chunk u>>3,
        //      byte r*128 + ((u&7) ^ (r&7))*16: the 128-byte swizzle the UMMA descriptor expects).  Pd sits on top of the stage's
        //      Q / K tiles: Q K^T has completed (s_full), nobody reads them any more.
        {
          uint8_t* prow = smem_raw + (uint32_t)(i % p.ns) * p.stage_bytes + p.off_pd + (uint32_t)row * 128u;
          const uint32_t rx = (uint32_t)(row & 7);
          const int units = kch * 8;
          const int u0 = valid ? rm.pi * NU : units;        // rows that belong to no problem: all zeros
          auto unit_ptr = [&](int u) { return reinterpret_cast<uint4*>(prow + (uint32_t)(u >> 3) * 16384u + ((((uint32_t)u & 7u) ^ rx) << 4)); };
          for (int u = 0; u < u0; ++u) *unit_ptr(u) = make_uint4(0, 0, 0, 0);
          for (int u = u0 + NU; u < units; ++u) *unit_ptr(u) = make_uint4(0, 0, 0, 0);
          if (valid) {
#pragma unroll
            for (int kk = 0; kk < NU; ++kk)
                *unit_ptr(u0 + kk) = make_uint4(pack_bf16(__uint_as_float(sv[kk * 8 + 0]), __uint_as_float(sv[kk * 8 + 1])),
                                                pack_bf16(__uint_as_float(sv[kk * 8 + 2]), __uint_as_float(sv[kk * 8 + 3])),
                                                pack_bf16(__uint_as_float(sv[kk * 8 + 4]), __uint_as_float(sv[kk * 8 + 5])),
                                                pack_bf16(__uint_as_float(sv[kk * 8 + 6]), __uint_as_float(sv[kk * 8 + 7])));
          }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pfull_bar(wg));
        if (valid && p.lse != nullptr) p.lse[(long long)pr * g.Sq + rm.qrow] = (mx + __log2f(l)) * 0.6931471805599453f;
        // ---- output tile
        mbar_wait(ofull_bar(wg), (uint32_t)use & 1u);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        const uint32_t t_o = tmem_base + lane_field + o_col(wg);
        tmem_ld_x32(t_o, o0);
        tmem_ld_x32(t_o + 32, o1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(oempty_bar(wg));
        const float inv = 1.0f / l;
        // the previous tile's output store must have finished READING this warp's staging rows
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        uint8_t* srow = stg + (uint32_t)row * 128u;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t* src = c < 4 ? &o0[c * 8] : &o1[(c - 4) * 8];
          *reinterpret_cast<uint4*>(srow + (((uint32_t)c ^ (uint32_t)(row & 7)) << 4)) =
              make_uint4(pack_bf16(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv), pack_bf16(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv),
                         pack_bf16(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv), pack_bf16(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          const uint32_t sst = stg_a + (uint32_t)(slot * 32) * 128u;
          if (!multi) {
            // lane 0 holds the first row of the warp's slot: if it maps to no problem row, neither does any other lane
            if (valid) tma_store_3d(&tm_o, sst, (pr % heads) * 64, rm.qrow, pr / heads);
          } else {
            for (int wdx = 0; wdx < 3; ++wdx)
              if (p0 + wdx < g.nprob) tma_store_3d(&tm_or, sst + (uint32_t)wdx * 1024u, ((p0 + wdx) % heads) * 64, 32, (p0 + wdx) / heads);
          }
          tma_store_commit();
        }
        (void)b; (void)h;
      }
      if (lane == 0) tma_store_wait_all();
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------------------------
// backward
//
//   S = Q K^T, dP = dO V^T                       (two MMAs into TMEM, same packed-tile layout as the forward)
//   P = exp2(S c + mask - lse), Pd = P o drop, delta_i = sum_j Pd_ij dP_ij, dS = P o (dP o drop - delta) * scale     (one thread per query row)
//   dV = Pd^T dO, dK = dS^T Q, dQ = dS K         (three MMAs; Pd / dS are written ONCE to shared memory as bf16 [query][key] tiles and
//                                                 serve as K-major A operand for dQ and as MN-major A operand for dV / dK)
// delta is computed from the recomputed probabilities (sum_j Pd_ij dP_ij = dO_i . O_i): the saved forward output is not read at all.
// Warps 4..7 do the softmax backward of tile i + 1 while warps 8..11 drain / store the dQ, dK, dV accumulators of tile i.
// Bias gradients of the projections (column sums of dQ / dK / dV per head):
//   * value bias: sum_k dV[k,:] = sum_q (sum_k Pd[q,k]) dO[q,:] -- the row sums of Pd are thread-local, so they are written (as a bf16
//     hi + lo pair) into two spare key slots of the packed key axis and the dV MMA itself produces the sums as two extra rows;
//   * query bias: halving butterfly (62 shuffles per warp) over the fp32 dQ accumulator rows in the epilogue;
//   * key bias: identically zero (rows of dS sum to zero: sum_j P_ij (dP_ij drop_ij - delta_i) = delta_i - delta_i): not accumulated.
// dQ / dK / dV leave through plain 16-byte global stores, one full 128-byte row segment per lane (TMEM lane = row): a staged TMA
// store needs shared memory the input ring cannot spare (two 64 KB stages + 64 KB of Pd / dS) or ties the stage release to the
// store's completion -- measured 239 us against 203 us for the panorama shape (profiles/r02_attn_bwd_notes.txt).
// Envelope: Sq <= 128 (one query tile per problem), packed key axis <= 128.
// ------------------------------------------------------------------------------------------------------------------------------
struct BwdParams {
  Geom g;
  int ns;
  uint32_t stage_bytes, off_do, off_k, off_v;     // per input stage: Q at 0, dO, K, V
  uint32_t off_pd, off_ds;                        // Pd / dS tiles (single), from the start of dynamic smem
  uint32_t off_end;                               // end of the zero-initialised region (stages + Pd + dS)
  int sum_slot;                                   // first of the 2 P spare key slots that carry the Pd row sums (hi, lo per problem)
  uint32_t off_mask, mask_floats;                 // per softmax warp
  uint32_t off_xch;                               // fp32 [2][4][128]: per-part partial delta / Pd row sums of the tile rows
  uint32_t off_dbias;                             // fp32 [5][heads * 64]: per-CTA bias-gradient sums, 4 epilogue warps x query bias + value bias (0 = not used: global atomics)
  uint32_t off_bar;
  uint32_t tx_q32, tx_q8, tx_kv;
  float scale_log2, scale;
  const float* mask;
  const float* lse;
  float* dbq; float* dbv;                         // [heads * 64] fp32 or null
  __nv_bfloat16 *dq, *dk, *dv;                    // outputs, q / k / v strides
  long long q_bs, ldq, kv_bs, ldkv;
  DropCfg drop;
};

// calls f(smem_byte_offset, small_box, first_query_row) for every row box of problem slot pi (the boxes the producer loads for Q / dO
// and the epilogue stores for dQ)
template <typename F>
__device__ __forceinline__ void for_each_row_box(const Geom& g, int pi, F&& f) {
  if (g.regime == 0) f((uint32_t)pi * 4096u, false, 0);
  else if (g.regime == 1) { f((uint32_t)pi * 4096u, false, 0); f(96u * 128u + (uint32_t)pi * 1024u, true, 32); }
  else if (g.regime == 2) { f((uint32_t)pi * 8192u, false, 0); if (g.Sq > 32) f((uint32_t)pi * 8192u + 4096u, false, 32); }
  else { for (int k = 0; k < 4; ++k) if (k * 32 < g.Sq) f((uint32_t)k * 4096u, false, k * 32); }
}

template <int NU, bool MULTI>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_qr, const __grid_constant__ CUtensorMap tm_do,
                   const __grid_constant__ CUtensorMap tm_dor, const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                   const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const Geom& g = p.g;
  constexpr int NCH = (NU + 3) / 4;
  constexpr int W = NU * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_base = smem_base + p.off_bar;
  // barriers: full[ns], empty[ns], sdp_full, pds_full (16 softmax warps), acc_full, acc_empty (4 epilogue warps), tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.ns + s); };
  const uint32_t sdp_full = bar_base + 8u * (2 * p.ns), pds_full = sdp_full + 8, acc_full = sdp_full + 16, acc_empty = sdp_full + 24;
  const uint32_t tmem_ptr_addr = sdp_full + 32;
  if (warp == 0 && lane == 0) {
    if (smem_base & 1023u) { printf("hamt attn bwd: dynamic smem base not 1024-byte aligned\n"); __trap(); }
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
    for (int s = 0; s < p.ns; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(sdp_full, 1); mbar_init(pds_full, kBwdSoftmaxWarps); mbar_init(acc_full, 1); mbar_init(acc_empty, 4);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr_addr);
  // everything that a TMA box may leave unwritten starts from zeros (finite garbage in unused operand rows / columns is harmless, NaN is not)
  for (uint32_t i = threadIdx.x * 16u; i < p.off_end; i += kBwdThreads * 16u) *reinterpret_cast<uint4*>(smem_raw + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_grid_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
  const int heads = g.heads;
  const int n_my = ((int)blockIdx.x < g.ntiles) ? (g.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int kch = (g.n_total + 63) >> 6;
  // TMEM columns: S 0, dP 128, dQ 256, dK 320, dV 384
  constexpr uint32_t C_S = 0, C_DP = 128, C_DQ = 256, C_DK = 320, C_DV = 384;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      for (int i = 0; i < n_my; ++i) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int s = i % p.ns;
        mbar_wait(empty_bar(s), ((uint32_t)(i / p.ns) & 1u) ^ 1u);
        const int p0 = tile * g.P, np = min(g.P, g.nprob - p0);
        const uint32_t sq = smem_base + (uint32_t)s * p.stage_bytes, sdo = sq + p.off_do, sk = sq + p.off_k, sv = sq + p.off_v;
        uint32_t bytes = 0;
        for (int pi = 0; pi < np; ++pi) {
          for_each_row_box(g, pi, [&](uint32_t, bool small, int) { bytes += 2 * (small ? p.tx_q8 : p.tx_q32); });
          bytes += 2 * p.tx_kv;
        }
        mbar_expect_tx(full_bar(s), bytes);
        for (int pi = 0; pi < np; ++pi) {
          const int pr = p0 + pi, b = pr / heads, h = pr % heads;
          for_each_row_box(g, pi, [&](uint32_t off, bool small, int q0) {
            tma_load_3d(sq + off, small ? &tm_qr : &tm_q, h * 64, q0, b, full_bar(s));
            tma_load_3d(sdo + off, small ? &tm_dor : &tm_do, h * 64, q0, b, full_bar(s));
          });
          tma_load_3d(sk + (uint32_t)(pi * W) * 128u, &tm_k, h * 64, 0, b, full_bar(s));
          tma_load_3d(sv + (uint32_t)(pi * W) * 128u, &tm_v, h * 64, 0, b, full_bar(s));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      const uint32_t idesc_s = umma_idesc_bf16(128, g.n_total, false, false);    // S / dP: [128 q] x [n_total keys], both K-major (d contiguous)
      const uint32_t idesc_kv = umma_idesc_bf16(128, 64, true, true);            // dV / dK: A = Pd / dS as [keys x q] (MN-major), B = dO / Q (MN-major)
      const uint32_t idesc_q = umma_idesc_bf16(128, 64, false, true);            // dQ: A = dS [q x keys] K-major, B = K (MN-major)
      const uint32_t spd = smem_base + p.off_pd, sds = smem_base + p.off_ds;
      int a_i = 0, b_i = 0;            // next tile for phase A (S, dP) / phase B (dV, dK, dQ)
      unsigned long long t0 = 0;
      uint32_t idle = 0;
      while (b_i < n_my) {
        bool progressed = false;
        // phase A of tile a_i: at most one tile ahead of phase B (letting it overtake phase B of the previous tile measured slower:
        // 333 vs 279 us, profiles/r02_attn_bwd_notes.txt); the S / dP columns are free once Pd / dS of the previous tile exist
        if (a_i < n_my && a_i <= b_i) {
          const int st = a_i % p.ns;
          if ((a_i == 0 || mbar_test_wait(pds_full, (uint32_t)(a_i - 1) & 1u)) && mbar_test_wait(full_bar(st), (uint32_t)(a_i / p.ns) & 1u)) {
            tc_fence_after();
            const uint32_t sq = smem_base + (uint32_t)st * p.stage_bytes, sdo = sq + p.off_do, sk = sq + p.off_k, sv = sq + p.off_v;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + C_S, umma_smem_desc(sq + k * 32, 16, 1024), umma_smem_desc(sk + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + C_DP, umma_smem_desc(sdo + k * 32, 16, 1024), umma_smem_desc(sv + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
            umma_commit(sdp_full);
            ++a_i;
            progressed = true;
          }
        }
        // phase B of tile b_i: Pd / dS published (and S / dP read), accumulators of the previous tile drained
        if (b_i < a_i && mbar_test_wait(pds_full, (uint32_t)b_i & 1u) && mbar_test_wait(acc_empty, ((uint32_t)b_i & 1u) ^ 1u)) {
          tc_fence_after();
          const int st = b_i % p.ns;
          const uint32_t sq = smem_base + (uint32_t)st * p.stage_bytes, sdo = sq + p.off_do, sk = sq + p.off_k;
#pragma unroll
          for (int k = 0; k < 8; ++k) {          // reduction over the 128 query rows, 16 per step
            const uint64_t da = umma_smem_desc(spd + (uint32_t)k * 2048u, 16384, 1024);
            const uint64_t db = umma_smem_desc(sdo + (uint32_t)k * 2048u, 8192, 1024);
            umma_bf16(tmem_base + C_DV, da, db, idesc_kv, k > 0 ? 1u : 0u);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t da = umma_smem_desc(sds + (uint32_t)k * 2048u, 16384, 1024);
            const uint64_t db = umma_smem_desc(sq + (uint32_t)k * 2048u, 8192, 1024);
            umma_bf16(tmem_base + C_DK, da, db, idesc_kv, k > 0 ? 1u : 0u);
          }
          const int ksteps = g.n_total >> 4;
          for (int k = 0; k < ksteps; ++k) {     // reduction over the packed key axis
            const uint64_t da = umma_smem_desc(sds + (uint32_t)(k >> 2) * 16384u + (uint32_t)(k & 3) * 32u, 16, 1024);
            const uint64_t db = umma_smem_desc(sk + (uint32_t)k * 2048u, 8192, 1024);
            umma_bf16(tmem_base + C_DQ, da, db, idesc_q, k > 0 ? 1u : 0u);
          }
          umma_commit(acc_full);
          umma_commit(empty_bar(st));          // Q / dO / K / V of this stage are no longer needed
          ++b_i;
          progressed = true;
        }
        if (progressed) idle = 0;
        else if ((++idle & 0xfffu) == 0) {
          if (t0 == 0) t0 = globaltimer_ns();
          else if (globaltimer_ns() - t0 > 4000000000ull) { printf("hamt attn bwd: MMA issuer stalled (block %d a %d b %d of %d)\n", blockIdx.x, a_i, b_i, n_my); __trap(); }
        }
      }
    }
  } else if (warp >= 4 && warp < 4 + kBwdSoftmaxWarps) {
    // ===================== softmax backward: Pd, dS =====================
    // 16 warps: warp = 4 + 4 * part + quarter.  The four warps of a quarter own the same 32 tile rows (TMEM lanes) and split the
    // key window of a row into four runs of 8-key units, so a thread handles at most MU = ceil(NU / 4) units.  With one thread per
    // whole row (4 warps, one per scheduler) the ~2300 dependent instructions of a row ran at 0.2 IPC and bounded the kernel
    // (profiles/r02_attn_bwd_notes.txt); the row-wide quantities (delta, the Pd row sum) now cross the four warps once per tile
    // through shared memory and a 128-thread named barrier.
    constexpr int MU = (NU + 3) / 4;
    constexpr bool KEEP_DP = false;             // keeping dP in registers for pass 2 measured SLOWER (155 -> 165 us, panorama shape): 80-register budget
    const int slot = warp & 3;
    const int part = (warp - 4) >> 2;
    const int ub = (part * NU) >> 2, ue = ((part + 1) * NU) >> 2;       // this thread's units of the key window
    const AttnDrop ds = attn_drop_init(p.drop);
    const bool multi = MULTI && slot == 3;
    float* smask = reinterpret_cast<float*>(smem_raw + p.off_mask) + (uint32_t)((warp - 4) * p.mask_floats);
    float* xch = reinterpret_cast<float*>(smem_raw + p.off_xch);       // [2][4 parts][128 rows]: partial delta, partial Pd row sum
    const int row = slot * 32 + lane;
    const uint32_t lane_field = (uint32_t)(slot * 32) << 16;
    uint8_t* pd_row = smem_raw + p.off_pd + (uint32_t)row * 128u;
    uint8_t* ds_row = smem_raw + p.off_ds + (uint32_t)row * 128u;
    const uint32_t rx = (uint32_t)(row & 7);
    for (int i = 0; i < n_my; ++i) {
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      const int p0 = tile * g.P;
      const RowMap rm = row_map(g, slot, lane, 0);
      const int pr = p0 + rm.pi;
      const bool valid = rm.ok && pr < g.nprob;
      const float lse2 = valid ? p.lse[(long long)pr * g.Sq + rm.qrow] * 1.4426950408889634f : 0.f;
      if (p.mask_floats != 0) {
        const int nw = multi ? g.nwin : 1;
        for (int wdx = 0; wdx < nw; ++wdx) {
          const int prw = multi ? p0 + wdx : __shfl_sync(0xffffffffu, pr, 0);
          const bool okw = multi ? (prw < g.nprob) : __shfl_sync(0xffffffffu, valid ? 1 : 0, 0) != 0;
          const int bw = okw ? prw / heads : 0;
          for (int j = ub * 8 + lane; j < ue * 8; j += 32) {
            float mv = -INFINITY;
            if (j < g.Sk) mv = okw ? p.mask[(long long)bw * g.Sk + j] * 1.4426950408889634f : 0.f;
            smask[wdx * NCH * 32 + j] = mv;
          }
        }
        __syncwarp();
      }
      mbar_wait(sdp_full, (uint32_t)i & 1u);
      tc_fence_after();
      const uint32_t t_s = tmem_base + lane_field + C_S, t_dp = tmem_base + lane_field + C_DP;
      const int wpi = __shfl_sync(0xffffffffu, valid ? rm.pi : 0, 0);
      // unit loader: 8 columns of S or dP at column c of this lane's key window (remainder warp: pick the lane's window out of three)
      auto load8 = [&](uint32_t tbase, int c, uint32_t (&out)[8]) {
        if (!MULTI || !multi) {
          tmem_ld_x8(tbase + (uint32_t)(wpi * W + c), out);
        } else {
          uint32_t t0[8], t1[8], t2[8];
          tmem_ld_x8(tbase + 0 * W + c, t0); tmem_ld_x8(tbase + 1 * W + c, t1); tmem_ld_x8(tbase + 2 * W + c, t2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) out[j] = rm.pi == 0 ? t0[j] : (rm.pi == 1 ? t1[j] : t2[j]);
        }
      };
      // ---- pass 1: probabilities (kept in registers; the sign bit marks a dropped entry), partial delta and partial Pd row sum
      float pv[MU * 8];
      uint32_t dpk[KEEP_DP ? MU * 8 : 1];        // dP of the thread's units, kept for pass 2 where the register budget allows
      float delta = 0.f, rs = 0.f;
      const uint32_t rowkey = attn_drop_rowkey(ds, (unsigned long long)pr * g.Sq + rm.qrow);
      const float* mrow = smask + (multi ? rm.pi * NCH * 32 : 0);
      const float dscale = ds.scale;
#pragma unroll
      for (int k = 0; k < MU; ++k) {
        const int u = ub + k;
        if (u < ue) {             // warp-uniform
          const int c = u * 8;
          uint32_t s8[8], d8[8];
          load8(t_s, c, s8);
          load8(t_dp, c, d8);
          tmem_ld_wait();
          if constexpr (KEEP_DP) {
#pragma unroll
            for (int j = 0; j < 8; ++j) dpk[k * 8 + j] = d8[j];
          }
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            float ta, tb;
            if (p.mask_floats != 0) {
              ta = fmaf(__uint_as_float(s8[j]), p.scale_log2, mrow[c + j]) - lse2;
              tb = fmaf(__uint_as_float(s8[j + 1]), p.scale_log2, mrow[c + j + 1]) - lse2;
            } else {
              ta = (c + j < g.Sk) ? fmaf(__uint_as_float(s8[j]), p.scale_log2, -lse2) : -INFINITY;
              tb = (c + j + 1 < g.Sk) ? fmaf(__uint_as_float(s8[j + 1]), p.scale_log2, -lse2) : -INFINITY;
            }
            const float pa = ex2_approx(ta), pb = ex2_approx(tb);
            float ma = 1.f, mb = 1.f;
            if (ds.on) {
              const uint32_t bits = attn_drop_bits(rowkey, (uint32_t)((c + j) >> 1));
              ma = (bits & 0xffffu) < ds.thresh16 ? 0.f : dscale;
              mb = (bits >> 16) < ds.thresh16 ? 0.f : dscale;
            }
            const float pda = pa * ma, pdb = pb * mb;
            delta = fmaf(pda, __uint_as_float(d8[j]), delta);
            delta = fmaf(pdb, __uint_as_float(d8[j + 1]), delta);
            rs += bf16_round(pda) + bf16_round(pdb);        // sum of the bf16-rounded values the dV MMA reads
            pv[k * 8 + j] = ma == 0.f ? -pa : pa;
            pv[k * 8 + j + 1] = mb == 0.f ? -pb : pb;
          }
        }
      }
      // ---- the row's delta and Pd row sum: partials of the four column parts through shared memory
      xch[part * 128 + row] = delta;
      xch[512 + part * 128 + row] = rs;
      asm volatile("bar.sync %0, 128;" ::"r"(2 + slot) : "memory");
      delta = (xch[row] + xch[128 + row]) + (xch[256 + row] + xch[384 + row]);
      // Pd / dS tiles are read by the MMAs of the previous tile until acc_full
      if (i > 0) mbar_wait(acc_full, (uint32_t)(i - 1) & 1u);
      // ---- pass 2: dS, and both rows to shared memory (zeros outside the lane's key window)
      {
        const int units = kch * 8;
        const int u0 = valid ? rm.pi * NU : units;
        auto unit_off = [&](int u) { return (uint32_t)(u >> 3) * 16384u + ((((uint32_t)u & 7u) ^ rx) << 4); };
        const int us = p.sum_slot >> 3;                  // the 16-byte unit that holds the row-sum slots of all problems
        // zero fill of the units outside the row's own window, shared out over the four parts
        for (int u = part; u < units; u += 4) {
          if (u >= u0 && u < u0 + NU) continue;
          if (u != us || !valid) *reinterpret_cast<uint4*>(pd_row + unit_off(u)) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(ds_row + unit_off(u)) = make_uint4(0, 0, 0, 0);
        }
        // (the TMEM loads are warp-collective: every lane issues them, only the arithmetic and the stores depend on `valid`)
#pragma unroll
        for (int k = 0; k < MU; ++k) {
          const int u = ub + k;
          if (u < ue) {
            uint32_t d8[8];
            if constexpr (KEEP_DP) {
#pragma unroll
              for (int j = 0; j < 8; ++j) d8[j] = dpk[k * 8 + j];
            } else {
              load8(t_dp, u * 8, d8);
              tmem_ld_wait();
            }
            if (valid) {
              float pdv[8], dsv[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float pj = pv[k * 8 + e];
                const float pa = fabsf(pj);
                const float m = pj < 0.f ? 0.f : dscale;        // (exact zero probabilities carry no sign: pd = 0 either way)
                pdv[e] = pa * m;
                dsv[e] = pa * (__uint_as_float(d8[e]) * m - delta) * p.scale;
              }
              const uint32_t off = unit_off(u0 + u);
              *reinterpret_cast<uint4*>(pd_row + off) = make_uint4(pack_bf16(pdv[0], pdv[1]), pack_bf16(pdv[2], pdv[3]), pack_bf16(pdv[4], pdv[5]), pack_bf16(pdv[6], pdv[7]));
              *reinterpret_cast<uint4*>(ds_row + off) = make_uint4(pack_bf16(dsv[0], dsv[1]), pack_bf16(dsv[2], dsv[3]), pack_bf16(dsv[4], dsv[5]), pack_bf16(dsv[6], dsv[7]));
            }
          }
        }
        if (valid && part == 0) {
          // value-bias gradient through the dV MMA: Pd[row, sum_slot + 2 pi] = hi(rs), [.. + 1] = lo(rs); zeros for the other problems
          const float rsum = (xch[512 + row] + xch[640 + row]) + (xch[768 + row] + xch[896 + row]);
          const float hi = bf16_round(rsum), lo = rsum - hi;
          const uint32_t pair = pack_bf16(hi, lo);
          uint4 w = make_uint4(0, 0, 0, 0);
          if (rm.pi == 0) w.x = pair; else if (rm.pi == 1) w.y = pair; else if (rm.pi == 2) w.z = pair; else w.w = pair;
          *reinterpret_cast<uint4*>(pd_row + unit_off(us)) = w;
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
    }
  } else if (warp >= 4 + kBwdSoftmaxWarps) {
    // ===================== accumulators -> global: dQ, dK, dV (+ bias gradients) =====================
    // 32 columns at a time (the 768-thread CTA leaves 80 registers per thread)
    const int slot = warp & 3;
    const int row = slot * 32 + lane;
    const uint32_t lane_field = (uint32_t)(slot * 32) << 16;
    // Bias-gradient column sums are collected per CTA in shared memory and leave with ONE global atomic per column and CTA at the
    // end.  Per-tile global atomics put ~10^6 adds on 1536 addresses for the panorama shape (155 us without, 250 us with bias
    // gradients); shared-memory fp32 atomics are compare-and-swap loops on sm_100 and were worse still (433 us).  So there are NO
    // atomics per tile: every epilogue warp owns a private [heads * 64] query-bias array (after the butterfly its lanes hold
    // distinct columns; the problems of one tile have distinct heads), the value-bias array is only touched by the warp that
    // holds the spare key rows.  Plain read-modify-write.
    const int ew = warp - 4 - kBwdSoftmaxWarps;
    float* sdbq = reinterpret_cast<float*>(smem_raw + p.off_dbias) + ew * heads * 64;
    float* sdbv = reinterpret_cast<float*>(smem_raw + p.off_dbias) + 4 * heads * 64;
    const bool cta_sums = p.dbq != nullptr && p.off_dbias != 0;
    if (cta_sums) {
      float* all = reinterpret_cast<float*>(smem_raw + p.off_dbias);
      for (int j = ew * 32 + lane; j < 5 * heads * 64; j += 128) all[j] = 0.f;
      asm volatile("bar.sync 6, 128;" ::: "memory");
    }
    for (int i = 0; i < n_my; ++i) {
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      const int p0 = tile * g.P, np = min(g.P, g.nprob - p0);
      // destination rows of this lane: its query row (dQ) and its key of the packed key axis (dK, dV)
      const RowMap rm = row_map(g, slot, lane, 0);
      const bool q_ok = rm.ok && rm.pi < np;
      const int kpi = row / W, key = row - kpi * W;
      const bool k_ok = kpi < np && key < g.Sk;
      __nv_bfloat16* gq = nullptr; __nv_bfloat16* gk = nullptr; __nv_bfloat16* gv = nullptr;
      if (q_ok) { const int pr = p0 + rm.pi; gq = p.dq + (long long)(pr / heads) * p.q_bs + (long long)rm.qrow * p.ldq + (pr % heads) * 64; }
      if (k_ok) {
        const int pr = p0 + kpi;
        const long long off = (long long)(pr / heads) * p.kv_bs + (long long)key * p.ldkv + (pr % heads) * 64;
        gk = p.dk + off; gv = p.dv + off;
      }
      mbar_wait(acc_full, (uint32_t)i & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int mh = 0; mh < 6; ++mh) {
        const int m = mh >> 1, hf = mh & 1;           // accumulator (dQ, dK, dV), 32-column half
        uint32_t acc[32];
        const uint32_t col = (m == 0 ? C_DQ : (m == 1 ? C_DK : C_DV)) + (uint32_t)hf * 32u;
        tmem_ld_x32(tmem_base + lane_field + col, acc);
        tmem_ld_wait();
        if (mh == 5) {                                // the last piece is in registers: the next tile's MMAs may overwrite the accumulators
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty);
        }
        __nv_bfloat16* dst = m == 0 ? gq : (m == 1 ? gk : gv);
        if (dst != nullptr) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(dst + hf * 32 + c * 8) =
                make_uint4(pack_bf16(__uint_as_float(acc[c * 8 + 0]), __uint_as_float(acc[c * 8 + 1])), pack_bf16(__uint_as_float(acc[c * 8 + 2]), __uint_as_float(acc[c * 8 + 3])),
                           pack_bf16(__uint_as_float(acc[c * 8 + 4]), __uint_as_float(acc[c * 8 + 5])), pack_bf16(__uint_as_float(acc[c * 8 + 6]), __uint_as_float(acc[c * 8 + 7])));
        }
        if (p.dbq != nullptr) {
          if (m == 0) {
            // query-bias gradient: column sums of this warp's 32 dQ rows (rows of no problem are exactly zero) by a halving
            // butterfly; the remainder warp of the 3-problem layout stops after the 8-lane groups (one problem each)
            const bool h1 = lane & 1, h2 = lane & 2, h4 = lane & 4, h8 = lane & 8, h16 = lane & 16;
            float w16[16], w8[8], w4[4];
#pragma unroll
            for (int j = 0; j < 16; ++j)
              w16[j] = __uint_as_float(h1 ? acc[16 + j] : acc[j]) + __shfl_xor_sync(0xffffffffu, __uint_as_float(h1 ? acc[j] : acc[16 + j]), 1);
#pragma unroll
            for (int j = 0; j < 8; ++j) w8[j] = (h2 ? w16[8 + j] : w16[j]) + __shfl_xor_sync(0xffffffffu, h2 ? w16[j] : w16[8 + j], 2);
#pragma unroll
            for (int j = 0; j < 4; ++j) w4[j] = (h4 ? w8[4 + j] : w8[j]) + __shfl_xor_sync(0xffffffffu, h4 ? w8[j] : w8[4 + j], 4);
            // lane now holds 4 columns starting at c4, summed over its group of 8 rows
            const int c4 = hf * 32 + (h1 ? 16 : 0) + (h2 ? 8 : 0) + (h4 ? 4 : 0);
            if (MULTI && slot == 3) {
              const int pi = lane >> 3;
              if (pi < np) {
                float* bd = (cta_sums ? sdbq : p.dbq) + ((p0 + pi) % heads) * 64 + c4;
                if (cta_sums) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) bd[j] += w4[j];
                } else {
#pragma unroll
                  for (int j = 0; j < 4; ++j) atomicAdd(bd + j, w4[j]);
                }
              }
            } else {
              float w2[2];
#pragma unroll
              for (int j = 0; j < 2; ++j) w2[j] = (h8 ? w4[2 + j] : w4[j]) + __shfl_xor_sync(0xffffffffu, h8 ? w4[j] : w4[2 + j], 8);
              const float w1 = (h16 ? w2[1] : w2[0]) + __shfl_xor_sync(0xffffffffu, h16 ? w2[0] : w2[1], 16);
              int pi;
              if (g.regime == 0 || g.regime == 1) pi = slot; else if (g.regime == 2) pi = slot >> 1; else pi = 0;
              if (pi < np) {
                float* bd = (cta_sums ? sdbq : p.dbq) + ((p0 + pi) % heads) * 64 + c4 + (h8 ? 2 : 0) + (h16 ? 1 : 0);
                if (cta_sums) *bd += w1; else atomicAdd(bd, w1);
              }
            }
          } else if (m == 2) {
            // value-bias gradient: the dV rows of the spare key slots hold sum_q rs_q dO[q,:] (hi and lo part)
            const int k = row - p.sum_slot;
            const bool mine = k >= 0 && k < 2 * np;
            if (cta_sums) {
              // (the spare rows of one tile sit in ONE warp: sum_slot is a multiple of 8 and 2 P <= 8) hi + lo row of a problem are
              // neighbouring lanes: fold them, then the even lane adds into the CTA's value-bias array
              if (((p.sum_slot >> 5) & 3) == slot) {      // warp-uniform
                float* bd = sdbv + ((p0 + (max(k, 0) >> 1)) % heads) * 64 + hf * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const float v = __uint_as_float(acc[j]) + __shfl_xor_sync(0xffffffffu, __uint_as_float(acc[j]), 1);
                  if (mine && !(k & 1)) bd[j] += v;
                }
              }
            } else if (mine) {
              float* bd = p.dbv + ((p0 + (k >> 1)) % heads) * 64 + hf * 32;
#pragma unroll
              for (int j = 0; j < 32; ++j) atomicAdd(bd + j, __uint_as_float(acc[j]));
            }
          }
        }
      }
    }
    if (cta_sums) {
      asm volatile("bar.sync 6, 128;" ::: "memory");
      const float* all = reinterpret_cast<const float*>(smem_raw + p.off_dbias);
      const int hc = heads * 64;
      for (int j = ew * 32 + lane; j < hc; j += 128) {
        const float vq = (all[j] + all[hc + j]) + (all[2 * hc + j] + all[3 * hc + j]), vv = all[4 * hc + j];
        if (vq != 0.f) atomicAdd(p.dbq + j, vq);
        if (vv != 0.f) atomicAdd(p.dbv + j, vv);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  return fn;
}

// bf16 tensor seen as [B][S][heads*64] with row pitch ld and sequence stride bs (elements); box = box_rows x 64 columns of one sequence
static int make_map3(CUtensorMap* tm, const void* ptr, int B, int S, int heads, long long ld, long long bs, int box_rows) {
  PFN_encodeTiled enc = encode_fn();
  HAMT_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[3] = {(cuuint64_t)heads * 64, (cuuint64_t)S, (cuuint64_t)B};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)bs * 2};
  cuuint32_t box[3] = {64u, (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  if (B == 1) gstr[1] = (cuuint64_t)ld * 2 * (cuuint64_t)S;      // unused stride must still be a multiple of 16 bytes
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[200];
    snprintf(buf, sizeof buf, "attn: cuTensorMapEncodeTiled failed (%d) B=%d S=%d heads=%d ld=%lld bs=%lld box=%d", (int)r, B, S, heads, ld, bs, box_rows);
    set_last_error(buf);
    return -2;
  }
  return 0;
}

// packing plan for (Sq, Sk); max_ntotal bounds the packed key axis (forward: 192 with two warpgroups)
static bool plan(Geom& g, int B, int heads, int Sq, int Sk, int max_ntotal, int spare_per_problem = 0) {
  g.nprob = B * heads; g.heads = heads; g.Sq = Sq; g.Sk = Sk;
  g.W = (Sk + 7) & ~7;
  if (g.W > 128) return false;                 // longer key axes: legacy kernel (RxR instructions)
  if (Sq <= 32) { g.regime = 0; g.P = 4; }
  else if (Sq <= 40 && g.W <= 64) { g.regime = 1; g.P = 3; }      // (the remainder warp holds three windows: compiled for W <= 64)
  else if (Sq <= 64) { g.regime = 2; g.P = 2; }
  else { g.regime = 3; g.P = 1; }
  // the packed key axis must fit the score tile; fall back to fewer problems per tile (regime 2 / 3 layouts)
  auto ntot = [&](int P) { return (P * g.W + P * spare_per_problem + 15) & ~15; };
  while (g.P > 1 && ntot(g.P) > max_ntotal) {
    if (g.regime == 1) { g.regime = 2; g.P = 2; }
    else if (g.regime == 0 && g.P == 4) { g.P = 2; }          // slots 0 and 1 only
    else if (g.regime == 0 && g.P == 2) { g.P = 1; }
    else { g.regime = 3; g.P = 1; }
  }
  if (ntot(g.P) > max_ntotal) return false;
  g.n_total = ntot(g.P);
  g.QT = g.regime == 3 ? (Sq + 127) / 128 : 1;
  g.ntiles = ((g.nprob + g.P - 1) / g.P) * g.QT;
  g.nwin = g.regime == 1 ? 3 : 1;
  return true;
}

static int g_attn_impl = 0;        // hamt_attn_set_impl: 0 = auto (tcgen05 kernels where they win), 1 = legacy mma.sync kernels only, 2 = tcgen05 wherever the shape fits

static int num_sms_cached() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int NU, bool MULTI>
static int launch_fwd(const AttnArgs& a, const Geom& g, cudaStream_t st) {
  constexpr int NCH = (NU + 3) / 4;
  FwdParams p{};
  p.g = g;
  p.nwg = 2;
  const uint32_t krows = (uint32_t)g.n_total;                       // rows of the packed K / V tiles
  const uint32_t kch = (uint32_t)(g.n_total + 63) / 64;             // 64-key chunks of Pd, 16 KB each
  const uint32_t pd_bytes = kch * 16384u;
  p.off_k = 16384u;
  // Pd (kch chunks of 16 KB) lies on top of Q | K: both are dead once S = Q K^T has completed.  The K region is widened to whole
  // chunks where necessary (at most 12 KB of slack) so that Pd never reaches V.
  uint32_t k_bytes = krows * 128u;
  if (k_bytes < (kch - 1) * 16384u) k_bytes = (kch - 1) * 16384u;
  p.off_v = p.off_k + k_bytes;
  p.off_pd = 0u;
  p.stage_bytes = p.off_v + krows * 128u;
  (void)pd_bytes;
  p.mask_floats = a.mask != nullptr ? (uint32_t)(g.nwin * NCH * 32) : 0u;      // no staging without a mask
  const uint32_t fixed = 2 * 16384u + 8u * p.mask_floats * 4u + 256u;
  int ns = (int)((232448u - 1024u - fixed) / p.stage_bytes);
  if (ns > 6) ns = 6;
  HAMT_REQUIRE(ns >= 2, "attn_fwd (tcgen05): shape does not fit two input stages");
  p.ns = ns;
  p.off_stg = (uint32_t)ns * p.stage_bytes;
  p.off_mask = p.off_stg + 2 * 16384u;
  p.off_bar = (p.off_mask + 8u * p.mask_floats * 4u + 15u) & ~15u;
  const size_t smem = p.off_bar + 256;
  p.tx_q32 = 32 * 128; p.tx_q8 = 8 * 128; p.tx_kv = (uint32_t)g.W * 128u; p.kv_box_rows = g.W;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.mask = a.mask; p.lse = a.lse;
  p.drop = DropCfg{a.drop.seed_ptr, a.drop.site, a.drop.p};
  CUtensorMap tq, tqr, tk, tv, to, tor;
  int rc;
  if ((rc = make_map3(&tq, a.q, a.B, a.Sq, a.heads, a.ldq, a.q_bstride, 32))) return rc;
  if ((rc = make_map3(&tqr, a.q, a.B, a.Sq, a.heads, a.ldq, a.q_bstride, 8))) return rc;
  if ((rc = make_map3(&tk, a.k, a.B, a.Sk, a.heads, a.ldkv, a.kv_bstride, g.W))) return rc;
  if ((rc = make_map3(&tv, a.v, a.B, a.Sk, a.heads, a.ldkv, a.kv_bstride, g.W))) return rc;
  if ((rc = make_map3(&to, a.out, a.B, a.Sq, a.heads, a.ldo, a.o_bstride, 32))) return rc;
  if ((rc = make_map3(&tor, a.out, a.B, a.Sq, a.heads, a.ldo, a.o_bstride, 8))) return rc;
  auto kern = attn_fwd_tc_kernel<NU, MULTI>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); return -3; }
    attr_set = true;
  }
  const int grid = g.ntiles < num_sms_cached() ? g.ntiles : num_sms_cached();
  launch_pdl(kern, grid, kThreads, smem, st, tq, tqr, tk, tv, to, tor, p);
  return check_launch("attn_fwd_tc_kernel");
}

template <int NU, bool MULTI>
static int launch_bwd(const AttnBwdArgs& b, const Geom& g, cudaStream_t st) {
  constexpr int NCH = (NU + 3) / 4;
  const AttnArgs& a = b.f;
  BwdParams p{};
  p.g = g;
  const uint32_t krows = (uint32_t)g.n_total, kch = (uint32_t)(g.n_total + 63) / 64;
  p.off_do = 16384u; p.off_k = 32768u; p.off_v = p.off_k + krows * 128u;
  p.stage_bytes = p.off_v + krows * 128u;
  p.mask_floats = a.mask != nullptr ? (uint32_t)(g.nwin * NCH * 32) : 0u;
  // (a 64-key chunk of Pd read as MN-major A operand with M = 128 reaches one chunk past a 64-key tile: keep 2 chunks per tile)
  const uint32_t tile_bytes = (kch < 2 ? 2u : kch) * 16384u;
  // (problems of one tile must have distinct heads: heads >= P; the spare key rows of a tile must sit in one warp: always true, see the kernel)
  const uint32_t dbias_bytes = (b.dbq != nullptr && b.dbv != nullptr && a.heads <= 16 && a.heads >= g.P) ? (uint32_t)(5 * a.heads * 64 * 4) : 0u;
  const uint32_t fixed = 2 * tile_bytes + (uint32_t)kBwdSoftmaxWarps * p.mask_floats * 4u + 4096u + dbias_bytes + 256u;
  int ns = (int)((232448u - 1024u - fixed) / p.stage_bytes);
  if (ns > 4) ns = 4;
  if (ns < 2) return 1;                                   // does not fit: the caller falls back to the legacy kernel
  p.ns = ns;
  p.off_pd = (uint32_t)ns * p.stage_bytes;
  p.off_ds = p.off_pd + tile_bytes;
  p.off_end = p.off_ds + tile_bytes;
  p.off_mask = p.off_end;
  p.off_xch = (p.off_mask + (uint32_t)kBwdSoftmaxWarps * p.mask_floats * 4u + 15u) & ~15u;
  p.off_dbias = dbias_bytes ? p.off_xch + 4096u : 0u;
  p.off_bar = p.off_xch + 4096u + dbias_bytes;
  p.sum_slot = g.P * g.W;
  const size_t smem = p.off_bar + 256;
  p.tx_q32 = 32 * 128; p.tx_q8 = 8 * 128; p.tx_kv = (uint32_t)g.W * 128u;
  p.scale = a.scale; p.scale_log2 = a.scale * 1.4426950408889634f;
  p.mask = a.mask; p.lse = a.lse; p.dbq = b.dbq; p.dbv = b.dbv;
  p.drop = DropCfg{a.drop.seed_ptr, a.drop.site, a.drop.p};
  p.dq = (__nv_bfloat16*)b.dq; p.dk = (__nv_bfloat16*)b.dk; p.dv = (__nv_bfloat16*)b.dv;
  p.q_bs = a.q_bstride; p.ldq = a.ldq; p.kv_bs = a.kv_bstride; p.ldkv = a.ldkv;
  CUtensorMap tq, tqr, tdo, tdor, tk, tv;
  int rc;
  if ((rc = make_map3(&tq, a.q, a.B, a.Sq, a.heads, a.ldq, a.q_bstride, 32))) return rc;
  if ((rc = make_map3(&tqr, a.q, a.B, a.Sq, a.heads, a.ldq, a.q_bstride, 8))) return rc;
  if ((rc = make_map3(&tdo, b.dout, a.B, a.Sq, a.heads, b.lddo, b.do_bstride, 32))) return rc;
  if ((rc = make_map3(&tdor, b.dout, a.B, a.Sq, a.heads, b.lddo, b.do_bstride, 8))) return rc;
  if ((rc = make_map3(&tk, a.k, a.B, a.Sk, a.heads, a.ldkv, a.kv_bstride, g.W))) return rc;
  if ((rc = make_map3(&tv, a.v, a.B, a.Sk, a.heads, a.ldkv, a.kv_bstride, g.W))) return rc;
  auto kern = attn_bwd_tc_kernel<NU, MULTI>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); return -3; }
    attr_set = true;
  }
  const int grid = g.ntiles < num_sms_cached() ? g.ntiles : num_sms_cached();
  launch_pdl(kern, grid, kBwdThreads, smem, st, tq, tqr, tdo, tdor, tk, tv, p);
  return check_launch("attn_bwd_tc_kernel");
}

}  // namespace tc

void attn_set_impl(int v) { tc::g_attn_impl = v; }

// returns 1 when the tcgen05 forward took the problem (status in *rc), 0 when the caller should run the legacy kernel
int attn_fwd_tc(const AttnArgs& a, cudaStream_t st, int* rc) {
  if (tc::g_attn_impl == 1) return 0;
  tc::Geom g;
  if (!tc::plan(g, a.B, a.heads, a.Sq, a.Sk, 192)) return 0;
  // Measured (profiles/r02_kbench_attn_fwd_v3.txt): with <= 32 query rows per problem (history-only vision stream of MLM / MRC / ITM:
  // 16 x 80, 16 x 16) the whole launch is a handful of tiles per SM and the fixed pipeline latency of this kernel (TMA -> MMA -> softmax
  // -> MMA -> TMA store) loses against the legacy kernel's 6 resident CTAs per SM (12.9 vs 9.1 us, 7.5 vs 4.6 us): those stay legacy
  // unless the tcgen05 path is forced (hamt_attn_set_impl(2)).
  if (g.regime == 0 && tc::g_attn_impl != 2) return 0;
  const int nu = g.W / 8;
  if (g.regime == 1) {
    switch (nu) {
#define HAMT_CASE(N_) case N_: *rc = tc::launch_fwd<N_, true>(a, g, st); break;
      HAMT_CASE(1) HAMT_CASE(2) HAMT_CASE(3) HAMT_CASE(4) HAMT_CASE(5) HAMT_CASE(6) HAMT_CASE(7) HAMT_CASE(8)
#undef HAMT_CASE
      default: return 0;
    }
  } else {
    switch (nu) {
#define HAMT_CASE(N_) case N_: *rc = tc::launch_fwd<N_, false>(a, g, st); break;
      HAMT_CASE(1) HAMT_CASE(2) HAMT_CASE(3) HAMT_CASE(4) HAMT_CASE(5) HAMT_CASE(6) HAMT_CASE(7) HAMT_CASE(8)
      HAMT_CASE(9) HAMT_CASE(10) HAMT_CASE(11) HAMT_CASE(12) HAMT_CASE(13) HAMT_CASE(14) HAMT_CASE(15) HAMT_CASE(16)
#undef HAMT_CASE
      default: return 0;
    }
  }
  return 1;
}

// backward: Sq <= 128 and a packed key axis of at most 128 keys (TMEM: S, dP, dQ, dK, dV = 448 of 512 columns)
int attn_bwd_tc(const AttnBwdArgs& b, cudaStream_t st, int* rc) {
  if (tc::g_attn_impl == 1) return 0;
  const AttnArgs& a = b.f;
  tc::Geom g;
  if (a.Sq > 128 || !tc::plan(g, a.B, a.heads, a.Sq, a.Sk, 128, 2) || g.QT != 1) return 0;     // + 2 spare key slots per problem (Pd row sums)
  // Dispatch by measurement (profiles/r02_kbench_attn_bwd_v6.txt, with fused bias gradients, us tcgen05 / legacy): panorama 36 x 36
  // 174 / 216 (3 problems per tile), 53 x 53 23.9 / 30.1 (2 per tile), 80 x 80 36.5 / 37.5, 80 x 53 31.8 / 33.4 -- and it loses where a
  // tile is mostly padding: 53 x 80 35.5 / 32.4 (one 53-row problem per 128-row tile, 2 x 80 keys exceed the 128-key accumulator),
  // 16 x 80 33.8 / 19.9, 80 x 16 28.4 / 26.7, 16 x 16 14.7 / 8.8.
  if (tc::g_attn_impl != 2) {
    const bool wins = g.regime == 1 || (g.regime == 2 && g.P == 2) || (g.regime == 3 && a.Sq > 64 && g.W >= 48);
    if (!wins) return 0;
  }
  const int nu = g.W / 8;
  int r = 1;
  if (g.regime == 1) {
    switch (nu) {
#define HAMT_CASE(N_) case N_: r = tc::launch_bwd<N_, true>(b, g, st); break;
      HAMT_CASE(1) HAMT_CASE(2) HAMT_CASE(3) HAMT_CASE(4) HAMT_CASE(5) HAMT_CASE(6) HAMT_CASE(7) HAMT_CASE(8)
#undef HAMT_CASE
      default: return 0;
    }
  } else {
    switch (nu) {
#define HAMT_CASE(N_) case N_: r = tc::launch_bwd<N_, false>(b, g, st); break;
      HAMT_CASE(1) HAMT_CASE(2) HAMT_CASE(3) HAMT_CASE(4) HAMT_CASE(5) HAMT_CASE(6) HAMT_CASE(7) HAMT_CASE(8)
      HAMT_CASE(9) HAMT_CASE(10) HAMT_CASE(11) HAMT_CASE(12) HAMT_CASE(13) HAMT_CASE(14) HAMT_CASE(15) HAMT_CASE(16)
#undef HAMT_CASE
      default: return 0;
    }
  }
  if (r == 1) return 0;       // shared memory does not fit two stages: legacy kernel
  *rc = r;
  return 1;
}

}  // namespace hamt
