// tcgen05 / TMEM / TMA GEMM for the HAMT hot path (sm_100a).
//
//   D[M,N] = sum_k A(m,k) * B(n,k)        bf16 operands, fp32 accumulation in TMEM
//
// replaces the eager nn.Linear forward (addmm), its dgrad (mm) and wgrad (mm) of the reference
// (pretrain_src/model/vilmodel.py:96-98,139-143,168-171,181-185; op census in SURVEY.md 8a).
//
// * Each operand is either K-major (reduction index contiguous: activations in forward, dY in
//   dgrad) or MN-major (output index contiguous: W in dgrad, dY^T and X in wgrad), selected through
//   the UMMA instruction-descriptor major bits, so no transposed copy of anything is ever made.
// * Persistent CTAs (one per SM) walk a static list of work units (m-tile, n-tile, k-split).
//   Warp 0 = TMA producer, warp 1 = TMEM owner + single-thread tcgen05.mma issuer, warps 2..9 =
//   epilogue.  smem ring of kStages {A,B} tiles (128-byte swizzle, written by TMA, read through UMMA
//   descriptors), two TMEM accumulator stages so the epilogue of unit i overlaps the main loop of
//   unit i+1.
// * Epilogue: TMEM -> registers (tcgen05.ld 32x32b) -> fused bias / GELU / ReLU / dGELU / alpha ->
//   XOR-swizzled per-warp shared-memory transpose -> fully coalesced 128-byte-row global stores
//   (bf16 store, bf16 read-modify-write, fp32 store / RMW, or red.global.add.v4.f32 for split-K).
//   The bias slice of the tile is staged once per tile in shared memory.
#include <cuda.h>
#include <stdio.h>
#include "hamt_common.cuh"
#include "hamt_kernels.h"

namespace hamt {

static constexpr int BM = 128;
static constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
static constexpr int kEpiWarps = 8;
static constexpr int kEpiThreads = kEpiWarps * 32;
static constexpr int kThreads = 64 + kEpiThreads;

struct GemmParams {
  int M, N, K;
  int tiles_m, tiles_n, splits, kb_per_split, kb_total;
  int m_fast;         // work-unit order: 0 = column tiles fastest (CTAs of a wave share the A rows; the usual case: few column tiles, weights
                      // resident in L2), 1 = row tiles fastest (few row tiles and a B operand larger than L2's share: the MLM decoder,
                      // 804 x 30522 x 768 -- with columns fastest its 47 MB weight was re-read from HBM once per row tile)
  void* out;
  long long ldo;
  int out_f32;        // 0: bf16 out, 1: fp32 out
  int out_mode;       // 0: store, 1: out += acc (plain RMW), 2: atomic add (split-K)
  const float* bias;  // [N] or null
  int act;            // 0 none, 1 gelu(erf), 2 relu
  int aux_mode;       // 0 none, 1 store pre-activation (bf16) to aux, 2 multiply by dgelu(aux), 3 multiply by (aux > 0),
                      // 4 store gelu'(pre-activation) to aux, 5 multiply by aux
  __nv_bfloat16* aux;
  long long ld_aux;
  float alpha;
  float* colsum;      // [N] fp32 or null: += column sums of the bf16 output (bias gradient of the producing Linear)
};

// PAIR = two CTAs of a cluster run one 256 x BN UMMA (cta_group::2): each stages 128 rows of A and BN/2 rows of B
template <int BN, bool PAIR = false>
struct SmemLayout {
  static constexpr int kBRows = PAIR ? BN / 2 : BN;
  static constexpr int kStages = (kBRows == 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingOffset = kStages * kStageBytes;        // 8 warps x 32 rows x 128 B
  static constexpr int kStagingBytes = kEpiWarps * 32 * 128;
  static constexpr int kBiasOffset = kStagingOffset + kStagingBytes;  // 2 x BN floats
  static constexpr int kBarOffset = kBiasOffset + 2 * BN * 4;
  static constexpr int kTotal = kBarOffset + 256;
  static_assert(kTotal <= 232448, "exceeds the 227 KB dynamic shared memory of sm_100");
};

template <int THREADS = kEpiThreads>
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); }

// 16-byte chunk `c` of row `r` inside a warp's [32][128 B] staging tile (XOR swizzle -> conflict-free both ways)
__device__ __forceinline__ uint32_t stage_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// epilogue specialisations (EPI template argument)
enum : int {
  EPI_GENERIC = 0,    // every option decided at run time (heads: ReLU / dReLU, odd combinations)
  EPI_STORE = 1,      // bf16 store (+ bias)
  EPI_GELU_PRE = 2,   // bf16 store of gelu(acc + bias), pre-activation stored to aux        (BertIntermediate forward)
  EPI_DGELU = 3,      // bf16 store of acc * gelu'(aux) (+ fused column sums)                 (BertOutput dgrad)
  EPI_ACCUM = 4,      // bf16 out += acc                                                      (dgrad onto the residual-path gradient)
  EPI_F32 = 5,        // fp32 store / RMW / split-K red                                       (wgrad, logits)
  EPI_GELU_DER = 6,   // bf16 store of gelu(acc + bias), gelu'(acc + bias) stored to aux      (BertIntermediate forward, round 2)
  EPI_MULAUX = 7      // bf16 store of acc * aux (+ fused column sums)                        (BertOutput dgrad on the saved derivative)
};
// Round 2: the forward GELU epilogue can save the DERIVATIVE gelu'(pre) instead of the pre-activation (one more ex2 on the pass that
// already evaluates the Gaussian tail), so the backward epilogue is a single multiply instead of 2 MUFU + ~20 ALU per element: the
// dGELU dgrad was the most expensive GEMM signature of the step (profiles/r01_ncu_ffn_gemms_v6.txt: 272 us at M = 34 560 against
// 117 us for the plain store; 183 us with the 16-warp epilogue).


// ---------------------------------------------------------------------------------------------------------------------------------
// Wide epilogue (EW = 16), used for the dGELU dgrad (measured 270 -> 183 us at M = 34 560, profiles/r02_kbench_variants.txt; it lost on
// the store / GELU / accumulate epilogues, which keep the 8-warp version).
// The GELU / dGELU epilogues are ALU-bound (2 MUFU + ~20 ALU per element; profiles/r01_ncu_ffn_gemms_v6.txt: 272 us against a
// 116 us tensor-pipe floor at M = 34 560) and with 8 warps only two warps per scheduler hide each other's latencies (issue slots
// 36 % busy).  Here 16 warps share a 128 x 256 accumulator: warp e owns TMEM lane quarter (warp & 3) and the 64-column slice e >> 2,
// processed as two 32-column half groups so that the live state fits the 96-register budget of 18 warps.  Staging tile per warp:
// [32 rows][64 B], chunk index XOR ((row >> 1) & 3): conflict-free for the row-owner accesses and for the coalesced phase (8 rows
// x 64 B per instruction).  Only fully aligned problems are dispatched here (M % tile rows == 0, N % 256 == 0, 16-byte aligned
// pitches), so there are no guards.
// ---------------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t stage_off32(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

template <int BN, bool PAIR, int EPI>
__device__ __forceinline__ void epilogue_wide(const GemmParams& p, uint8_t* staging, float* sbias, uint32_t tmem_base, uint32_t tfull0,
                                              uint32_t tempty0, uint32_t rank, int u_first, int u_stride, int units, int warp, int lane) {
  static_assert(BN == 256, "wide epilogue: 128 x 256 accumulators only");
  static_assert(EPI == EPI_STORE || EPI == EPI_GELU_PRE || EPI == EPI_DGELU || EPI == EPI_ACCUM, "wide epilogue: bf16 store paths only");
  constexpr int TM = PAIR ? 2 * BM : BM;
  constexpr int kET = 16 * 32;
  const int e = warp - 2;
  const int quarter = warp & 3;
  const int slice = e >> 2;                       // 64-column slice of the tile
  const int et = (int)threadIdx.x - 64;
  uint8_t* stg = staging + e * (32 * 64);
  const int r_co = lane >> 2, c_co = lane & 3;    // coalesced phase: 8 rows x 4 chunks of 16 B per instruction
  const bool has_bias = p.bias != nullptr;
  const bool do_colsum = EPI == EPI_DGELU && p.colsum != nullptr;
  const __nv_bfloat16* pre_src = EPI == EPI_DGELU ? p.aux : (EPI == EPI_ACCUM ? reinterpret_cast<const __nv_bfloat16*>(p.out) : nullptr);
  const long long pre_ld = EPI == EPI_DGELU ? p.ld_aux : p.ldo;
  int as = 0;
  uint32_t aphase = 0;
  for (int u = u_first; u < units; u += u_stride) {
    const int tn = p.m_fast ? (u / p.tiles_m) % p.tiles_n : u % p.tiles_n;
    const int tm = p.m_fast ? u % p.tiles_m : (u / p.tiles_n) % p.tiles_m;
    if (has_bias) {
      if (et < BN) sbias[as * BN + et] = __ldg(p.bias + tn * BN + et);
      epi_bar_sync<kET>();
    }
    const int row0 = tm * TM + (int)rank * BM + quarter * 32;
    uint4 pre[4];
    auto issue_pre = [&](int hg) {
      const __nv_bfloat16* ap = pre_src + (long long)(row0 + r_co) * pre_ld + tn * BN + slice * 64 + hg * 32 + c_co * 8;
#pragma unroll
      for (int it = 0; it < 4; ++it) pre[it] = *reinterpret_cast<const uint4*>(ap + (long long)(it * 8) * pre_ld);
    };
    if (EPI == EPI_DGELU || EPI == EPI_ACCUM) issue_pre(0);
    mbar_wait(tfull0 + 8u * as, aphase);
    tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + slice * 64);
    const float* bias_t = sbias + as * BN + slice * 64;
#pragma unroll 1
    for (int hg = 0; hg < 2; ++hg) {
      uint32_t r[32];
      tmem_ld_32x32(t_row + hg * 32, r);
      tmem_ld_wait();
      const int col0 = tn * BN + slice * 64 + hg * 32;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (has_bias) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_t + hg * 32 + j);
          v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
        }
      }
      auto stage_own_row = [&]() {       // this lane's row (= its TMEM lane) -> 4 chunks of 8 bf16
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(stg + stage_off32(lane, c)) =
              make_uint4(pack_bf16(v[c * 8], v[c * 8 + 1]), pack_bf16(v[c * 8 + 2], v[c * 8 + 3]), pack_bf16(v[c * 8 + 4], v[c * 8 + 5]),
                         pack_bf16(v[c * 8 + 6], v[c * 8 + 7]));
      };
      if (EPI == EPI_GELU_PRE) {
        // the pre-activation goes out first (saved for the dGELU backward), then the activation through the same staging tile
        stage_own_row();
        __syncwarp();
        __nv_bfloat16* ap = p.aux + (long long)(row0 + r_co) * p.ld_aux + col0 + c_co * 8;
#pragma unroll
        for (int it = 0; it < 4; ++it)
          *reinterpret_cast<uint4*>(ap + (long long)(it * 8) * p.ld_aux) = *reinterpret_cast<const uint4*>(stg + stage_off32(it * 8 + r_co, c_co));
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
      } else if (EPI == EPI_DGELU) {
        // saved pre-activation: fetched coalesced (issue_pre), transposed through the staging tile so that every lane gets its row
#pragma unroll
        for (int it = 0; it < 4; ++it) *reinterpret_cast<uint4*>(stg + stage_off32(it * 8 + r_co, c_co)) = pre[it];
        __syncwarp();
        if (hg == 0) issue_pre(1);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 w = *reinterpret_cast<const uint4*>(stg + stage_off32(lane, c));
          const float2 f0 = unpack_bf16(w.x), f1 = unpack_bf16(w.y), f2 = unpack_bf16(w.z), f3 = unpack_bf16(w.w);
          const float a[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
          for (int j = 0; j < 8; ++j) v[c * 8 + j] *= dgelu_erf(a[j]);
        }
        __syncwarp();
      }
      stage_own_row();
      __syncwarp();
      float cs[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) cs[j] = 0.f;
      __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)(row0 + r_co) * p.ldo + col0 + c_co * 8;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        uint4 w = *reinterpret_cast<const uint4*>(stg + stage_off32(it * 8 + r_co, c_co));
        if (EPI == EPI_DGELU) {
          if (do_colsum) {
            const float2 f0 = unpack_bf16(w.x), f1 = unpack_bf16(w.y), f2 = unpack_bf16(w.z), f3 = unpack_bf16(w.w);
            cs[0] += f0.x; cs[1] += f0.y; cs[2] += f1.x; cs[3] += f1.y; cs[4] += f2.x; cs[5] += f2.y; cs[6] += f3.x; cs[7] += f3.y;
          }
        } else if (EPI == EPI_ACCUM) {     // gradient accumulation onto the residual-path gradient (old values prefetched)
          const uint32_t ws[4] = {w.x, w.y, w.z, w.w}, os[4] = {pre[it].x, pre[it].y, pre[it].z, pre[it].w};
          uint32_t rs[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 x = unpack_bf16(ws[k]), y = unpack_bf16(os[k]);
            rs[k] = pack_bf16(x.x + y.x, x.y + y.y);
          }
          w = make_uint4(rs[0], rs[1], rs[2], rs[3]);
        }
        *reinterpret_cast<uint4*>(op + (long long)(it * 8) * p.ldo) = w;
      }
      if (do_colsum) {      // warp-uniform; rows: 4 per thread above, then across the 8 row sub-groups (lane bits 2..4)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 4);
          cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 8);
          cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 16);
        }
        if (r_co == 0) {
          float* cp = p.colsum + col0 + c_co * 8;
          red_add_v4(cp, cs[0], cs[1], cs[2], cs[3]);
          red_add_v4(cp + 4, cs[4], cs[5], cs[6], cs[7]);
        }
      }
      __syncwarp();
      if (EPI == EPI_ACCUM && hg == 0) issue_pre(1);      // old values of the second half group
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (!PAIR || rank == 0) mbar_arrive(tempty0 + 8u * as);
      else mbar_arrive_cluster(mapa_u32(tempty0 + 8u * as, 0));
    }
    if (++as == 2) { as = 0; aphase ^= 1u; }
  }
}

// Register budget: 10 warps land 3 + 3 + 2 + 2 on the four SM sub-partitions (16 K registers each), so __launch_bounds__(320, 1) caps
// ptxas at 168 registers; asking for more ("too many resources requested for launch") does not fit 3 warps x 32 lanes.
// EW = epilogue warps: 8 (each warp owns a 32-row x BN/2 slab, 64-column groups) or 16 (wide epilogue of the dGELU dgrad, see epilogue_wide).
template <int BN, bool A_MN, bool B_MN, bool PAIR, int EPI, int EW = kEpiWarps>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmParams p) {
  using L = SmemLayout<BN, PAIR>;
  constexpr int TM = PAIR ? 2 * BM : BM;            // rows of the output tile owned by one CTA (pair)
  constexpr int BNL = L::kBRows;                    // B rows staged by this CTA
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  constexpr int kStages = L::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t bar_base = smem_base + L::kBarOffset;
  // barrier layout (8 B each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    if (smem_base & 1023u) { printf("hamt gemm: dynamic smem base not 1024-byte aligned\n"); __trap(); }
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);                    // pair: the leader's expect_tx covers the bytes of BOTH CTAs' loads
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), PAIR ? 2 * EW : EW);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair<2 * BN>(tmem_ptr_addr);
    else tmem_alloc<2 * BN>(tmem_ptr_addr);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                     // peer barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  pdl_grid_sync();   // everything above overlapped the previous kernel; global memory is touched only below
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  const int units = p.tiles_m * p.tiles_n * p.splits;
  const int u_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int u_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int u = u_first; u < units; u += u_stride) {
        const int tn = p.m_fast ? (u / p.tiles_m) % p.tiles_n : u % p.tiles_n;
        const int tm = p.m_fast ? u % p.tiles_m : (u / p.tiles_n) % p.tiles_m;
        const int sp = u / (p.tiles_n * p.tiles_m);
        const int kb0 = sp * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * L::kStageBytes;
          const uint32_t sb = sa + L::kABytes;
          const int m_row = tm * TM + (int)rank * BM, n_row = tn * BN + (int)rank * BNL;
          if (!PAIR) {
            mbar_expect_tx(full_bar(stage), L::kStageBytes);
            if (!A_MN) {
              tma_load_2d(sa, &tma_a, kb * BK, m_row, full_bar(stage));
            } else {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * (BK * 128), &tma_a, m_row + c * 64, kb * BK, full_bar(stage));
            }
            if (!B_MN) {
              tma_load_2d(sb, &tma_b, kb * BK, n_row, full_bar(stage));
            } else {
#pragma unroll
              for (int c = 0; c < BNL / 64; ++c) tma_load_2d(sb + c * (BK * 128), &tma_b, n_row + c * 64, kb * BK, full_bar(stage));
            }
          } else {
            // both CTAs' bytes are counted on the LEADER's full barrier, whose single arrival is the leader's expect_tx for
            // 2 x stage bytes.  (A remote release-arrive from the peer per k-block stalled its producer ~1000 cycles and halved
            // the MMA rate -- profiles/r01_ncu_pair.txt.)  The peer cannot run a phase ahead: it waits on its own empty barrier,
            // which the leader's MMA commit signals only after the previous phase of this full barrier has completed.
            const uint32_t lbar = mapa_u32(full_bar(stage), 0);
            if (leader) mbar_expect_tx(full_bar(stage), 2 * L::kStageBytes);
            if (!A_MN) {
              tma_load_2d_pair(sa, &tma_a, kb * BK, m_row, lbar);
            } else {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c) tma_load_2d_pair(sa + c * (BK * 128), &tma_a, m_row + c * 64, kb * BK, lbar);
            }
            if (!B_MN) {
              tma_load_2d_pair(sb, &tma_b, kb * BK, n_row, lbar);
            } else {
#pragma unroll
              for (int c = 0; c < BNL / 64; ++c) tma_load_2d_pair(sb + c * (BK * 128), &tma_b, n_row + c * 64, kb * BK, lbar);
            }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===================== MMA issuer (pair: the leader CTA issues for both) =====================
      constexpr uint32_t idesc = umma_idesc_bf16(TM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int u = u_first; u < units; u += u_stride) {
        const int sp = u / (p.tiles_n * p.tiles_m);
        const int kb0 = sp * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * L::kStageBytes;
          const uint32_t sb = sa + L::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: 16 elements = 32 B further along the swizzled row; SBO = 8 rows * 128 B.
            // MN-major: 16 k-rows = 2048 B further; LBO = stride between 64-element MN chunks
            // (one TMA box = BK rows * 128 B), SBO = 8 k-rows * 128 B.
            const uint64_t da = A_MN ? umma_smem_desc(sa + k * 2048, BK * 128, 1024) : umma_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_smem_desc(sb + k * 2048, BK * 128, 1024) : umma_smem_desc(sb + k * 32, 16, 1024);
            if (PAIR) umma_bf16_pair(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_bf16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // smem slot reusable once these MMAs have read it (pair: signalled to both CTAs' producers)
          if (PAIR) umma_commit_pair(empty_bar(stage), 3); else umma_commit(empty_bar(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (PAIR) umma_commit_pair(tfull_bar(as), 3); else umma_commit(tfull_bar(as));  // accumulator complete -> epilogue(s)
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else if constexpr (EW == 16) {
    epilogue_wide<BN, PAIR, EPI>(p, smem_raw + L::kStagingOffset, reinterpret_cast<float*>(smem_raw + L::kBiasOffset), tmem_base,
                                 tfull_bar(0), tempty_bar(0), rank, u_first, u_stride, units, warp, lane);
  } else {
    // ===================== epilogue warps =====================
    // The epilogue is specialised at compile time (EPI) for the combinations the hot path runs thousands of times per step, so
    // each instantiation carries only its own straight-line code: the one-size-fits-all version was ~4 k SASS instructions and
    // lost a quarter of its issue slots to instruction-cache misses (profiles/r01_ncu_gelu_epilogue.txt, stall_no_inst 24 %).
    // Interior tiles (all 32 rows of the warp's slab < M, full 128-byte column group, 16-byte aligned pitches) take a
    // predicate-free path; edge tiles run compact rolled loops with scalar accesses.
    constexpr bool GEN = EPI == EPI_GENERIC;
    const int e = warp - 2;
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int half = e >> 2;               // which half of the BN columns
    const int et = threadIdx.x - 64;       // 0..255 within the epilogue group
    constexpr int kColsPerWarp = BN / 2;
    uint8_t* stg = smem_raw + L::kStagingOffset + e * (32 * 128);
    float* sbias = reinterpret_cast<float*>(smem_raw + L::kBiasOffset);
    int as = 0;
    uint32_t aphase = 0;
    const bool out_f32 = GEN ? (p.out_f32 != 0) : (EPI == EPI_F32);
    const int act = GEN ? p.act : ((EPI == EPI_GELU_PRE || EPI == EPI_GELU_DER) ? 1 : 0);
    const int aux_mode = GEN ? p.aux_mode : (EPI == EPI_GELU_PRE ? 1 : (EPI == EPI_DGELU ? 2 : (EPI == EPI_GELU_DER ? 4 : (EPI == EPI_MULAUX ? 5 : 0))));
    const bool aux_store = aux_mode == 1 || aux_mode == 4;                       // the epilogue WRITES aux
    const bool aux_read = aux_mode == 2 || aux_mode == 3 || aux_mode == 5;       // the epilogue READS aux
    // mode 5 on interior tiles multiplies in the COALESCED phase (the prefetched aux chunks already have that layout), like the
    // accumulate epilogue adds: no transposition of the aux tile through shared memory.  The accumulator is rounded to bf16 by the
    // staging tile first, so the product is rounded twice (2^-9 relative, the level of the saved derivative itself).
    const bool mul_co = EPI == EPI_MULAUX;
    const int out_mode = (GEN || EPI == EPI_F32) ? p.out_mode : (EPI == EPI_ACCUM ? 1 : 0);
    const bool do_colsum = (GEN || EPI == EPI_DGELU || EPI == EPI_MULAUX) && p.colsum != nullptr;
    const bool use_alpha = (GEN || EPI == EPI_F32) && p.alpha != 1.0f;
    const int esz = out_f32 ? 4 : 2;
    const bool out_vec_ok = (((uintptr_t)p.out & 15) == 0) && ((p.ldo * esz) % 16 == 0);
    const bool aux_vec_ok = p.aux != nullptr && (p.ld_aux & 7) == 0 && ((uintptr_t)p.aux & 15) == 0;
    const bool has_bias = p.bias != nullptr;
    const int r_sub = lane >> 3, c_sub = lane & 7;   // coalesced phase: 4 rows x 8 chunks of 16 B per instruction
    const __nv_bfloat16* pre_src = nullptr;          // bf16 tile the epilogue has to READ: saved pre-activation, or the old gradient
    long long pre_ld = 0;
    if (!out_f32) {
      if (aux_read) { pre_src = p.aux; pre_ld = p.ld_aux; }
      else if (out_mode == 1) { pre_src = reinterpret_cast<const __nv_bfloat16*>(p.out); pre_ld = p.ldo; }
    }
    const bool pre_vec_ok = pre_src != nullptr && (pre_ld & 7) == 0 && ((uintptr_t)pre_src & 15) == 0;
    constexpr int kGroups = (BN / 2) / 64;

    for (int u = u_first; u < units; u += u_stride) {
      const int tn = p.m_fast ? (u / p.tiles_m) % p.tiles_n : u % p.tiles_n;
      const int tm = p.m_fast ? u % p.tiles_m : (u / p.tiles_n) % p.tiles_m;
      if (has_bias) {   // stage this tile's bias slice (double-buffered by accumulator stage)
        if (et < BN) sbias[as * BN + et] = (tn * BN + et < p.N) ? __ldg(p.bias + tn * BN + et) : 0.f;
        epi_bar_sync();
      }
      const int row0 = tm * TM + (int)rank * BM + quarter * 32;
      const bool rows_full = row0 + 32 <= p.M;
      // guarded 16-byte chunk accessors for edge tiles (row / column bounds, unaligned pitches)
      auto load_chunk = [&](const __nv_bfloat16* base, long long ld, bool vec, int row, int col, int nvalid) -> uint4 {
        uint4 w = make_uint4(0, 0, 0, 0);
        if (row < p.M && nvalid > 0) {
          const __nv_bfloat16* ap = base + (long long)row * ld + col;
          if (vec && nvalid >= 8) w = *reinterpret_cast<const uint4*>(ap);
          else {
            __nv_bfloat16* hw = reinterpret_cast<__nv_bfloat16*>(&w);
            for (int j = 0; j < 8 && j < nvalid; ++j) hw[j] = ap[j];
          }
        }
        return w;
      };
      auto store_chunk = [&](__nv_bfloat16* base, long long ld, bool vec, int row, int col, int nvalid, uint4 w) {
        if (row < p.M && nvalid > 0) {
          __nv_bfloat16* ap = base + (long long)row * ld + col;
          if (vec && nvalid >= 8) *reinterpret_cast<uint4*>(ap) = w;
          else {
            const __nv_bfloat16* hw = reinterpret_cast<const __nv_bfloat16*>(&w);
            for (int j = 0; j < 8 && j < nvalid; ++j) ap[j] = hw[j];
          }
        }
      };
      // Operand tiles the epilogue has to READ are fetched -- coalesced, 4 rows x 128 B per instruction -- BEFORE waiting for the
      // accumulator, so their HBM latency hides behind the main loop of this tile; the next group's tile is requested while the
      // current one is processed.  (Interior tiles only; edge tiles load at the point of use.)
      uint4 pre[8];
      auto group_fast = [&](int gI) {
        const int col0 = tn * BN + half * kColsPerWarp + gI * 64;
        return rows_full && col0 + 64 <= p.N && out_vec_ok && (pre_src == nullptr || pre_vec_ok) && (!aux_store || aux_vec_ok);
      };
      auto issue_pre = [&](int gI) {
        if (!group_fast(gI)) return;
        const __nv_bfloat16* ap = pre_src + (long long)(row0 + r_sub) * pre_ld + tn * BN + half * kColsPerWarp + gI * 64 + c_sub * 8;
#pragma unroll
        for (int it = 0; it < 8; ++it) pre[it] = *reinterpret_cast<const uint4*>(ap + (long long)(it * 4) * pre_ld);
      };
      if (pre_src != nullptr) issue_pre(0);
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + half * kColsPerWarp);
      const float* bias_t = sbias + as * BN + half * kColsPerWarp;

      if (!out_f32) {
        // ---------------- bf16 output: groups of 64 columns (one 128-byte row segment) ----------------
#pragma unroll 1
        for (int gI = 0; gI < kGroups; ++gI) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(t_row + gI * 64, r0);
          tmem_ld_32x32(t_row + gI * 64 + 32, r1);
          tmem_ld_wait();
          const int col0 = tn * BN + half * kColsPerWarp + gI * 64;
          if (col0 >= p.N) continue;  // warp-uniform
          const bool fast = group_fast(gI);
          const int ncols = min(64, p.N - col0);
          float v[64];
#pragma unroll
          for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
          if (use_alpha) {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] *= p.alpha;
          }
          if (has_bias) {
#pragma unroll
            for (int j = 0; j < 64; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias_t + gI * 64 + j);
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
          if (aux_store) {
            // also emit the pre-activation (needed by the dGELU / dReLU backward) -- or, mode 4, the derivative itself
            if (aux_mode == 4) {
              // gelu and its derivative share the Gaussian tail: v becomes the activation here (the act pass below is skipped)
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                float d8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) gelu_and_derivative(v[c * 8 + j], v[c * 8 + j], d8[j]);
                *reinterpret_cast<uint4*>(stg + stage_off(lane, c)) =
                    make_uint4(pack_bf16(d8[0], d8[1]), pack_bf16(d8[2], d8[3]), pack_bf16(d8[4], d8[5]), pack_bf16(d8[6], d8[7]));
              }
            } else {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<uint4*>(stg + stage_off(lane, c)) =
                  make_uint4(pack_bf16(v[c * 8], v[c * 8 + 1]), pack_bf16(v[c * 8 + 2], v[c * 8 + 3]), pack_bf16(v[c * 8 + 4], v[c * 8 + 5]),
                             pack_bf16(v[c * 8 + 6], v[c * 8 + 7]));
            }
            __syncwarp();
            if (fast) {
              __nv_bfloat16* ap = p.aux + (long long)(row0 + r_sub) * p.ld_aux + col0 + c_sub * 8;
#pragma unroll
              for (int it = 0; it < 8; ++it)
                *reinterpret_cast<uint4*>(ap + (long long)(it * 4) * p.ld_aux) = *reinterpret_cast<const uint4*>(stg + stage_off(it * 4 + r_sub, c_sub));
            } else {
#pragma unroll 1
              for (int it = 0; it < 8; ++it) {
                const int r = it * 4 + r_sub;
                store_chunk(p.aux, p.ld_aux, aux_vec_ok, row0 + r, col0 + c_sub * 8, ncols - c_sub * 8, *reinterpret_cast<const uint4*>(stg + stage_off(r, c_sub)));
              }
            }
            __syncwarp();
          } else if (aux_read && !(mul_co && fast)) {
            // the aux tile (pre-activation / derivative saved by the forward) goes through smem so that every lane gets its own row
            if (fast) {
#pragma unroll
              for (int it = 0; it < 8; ++it) *reinterpret_cast<uint4*>(stg + stage_off(it * 4 + r_sub, c_sub)) = pre[it];
            } else {
#pragma unroll 1
              for (int it = 0; it < 8; ++it) {
                const int r = it * 4 + r_sub;
                *reinterpret_cast<uint4*>(stg + stage_off(r, c_sub)) = load_chunk(pre_src, pre_ld, pre_vec_ok, row0 + r, col0 + c_sub * 8, ncols - c_sub * 8);
              }
            }
            __syncwarp();
            if (gI + 1 < kGroups) issue_pre(gI + 1);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint4 w = *reinterpret_cast<const uint4*>(stg + stage_off(lane, c));
              const float2 f0 = unpack_bf16(w.x), f1 = unpack_bf16(w.y), f2 = unpack_bf16(w.z), f3 = unpack_bf16(w.w);
              const float a[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
              for (int j = 0; j < 8; ++j)
                v[c * 8 + j] = aux_mode == 2 ? v[c * 8 + j] * dgelu_erf(a[j]) : (aux_mode == 5 ? v[c * 8 + j] * a[j] : (a[j] > 0.f ? v[c * 8 + j] : 0.f));
            }
            __syncwarp();
          }
          if (act == 1 && aux_mode != 4) {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = gelu_erf(v[j]);
          } else if (act == 2) {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = fmaxf(v[j], 0.f);
          }
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(stg + stage_off(lane, c)) =
                make_uint4(pack_bf16(v[c * 8], v[c * 8 + 1]), pack_bf16(v[c * 8 + 2], v[c * 8 + 3]), pack_bf16(v[c * 8 + 4], v[c * 8 + 5]),
                           pack_bf16(v[c * 8 + 6], v[c * 8 + 7]));
          __syncwarp();
          float cs[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) cs[j] = 0.f;
          auto add_bf16x8 = [](uint4 a, uint4 b) {
            const uint32_t as_[4] = {a.x, a.y, a.z, a.w}, bs_[4] = {b.x, b.y, b.z, b.w};
            uint32_t rs[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 x = unpack_bf16(as_[k]), y = unpack_bf16(bs_[k]);
              rs[k] = pack_bf16(x.x + y.x, x.y + y.y);
            }
            return make_uint4(rs[0], rs[1], rs[2], rs[3]);
          };
          auto mul_bf16x8 = [](uint4 a, uint4 b) {
            const uint32_t as_[4] = {a.x, a.y, a.z, a.w}, bs_[4] = {b.x, b.y, b.z, b.w};
            uint32_t rs[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 x = unpack_bf16(as_[k]), y = unpack_bf16(bs_[k]);
              rs[k] = pack_bf16(x.x * y.x, x.y * y.y);
            }
            return make_uint4(rs[0], rs[1], rs[2], rs[3]);
          };
          auto acc_cs = [&](uint4 w) {   // column sums of exactly the bf16 values a separate pass over the output would read
            const float2 f0 = unpack_bf16(w.x), f1 = unpack_bf16(w.y), f2 = unpack_bf16(w.z), f3 = unpack_bf16(w.w);
            cs[0] += f0.x; cs[1] += f0.y; cs[2] += f1.x; cs[3] += f1.y; cs[4] += f2.x; cs[5] += f2.y; cs[6] += f3.x; cs[7] += f3.y;
          };
          if (fast) {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)(row0 + r_sub) * p.ldo + col0 + c_sub * 8;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              uint4 w = *reinterpret_cast<const uint4*>(stg + stage_off(it * 4 + r_sub, c_sub));
              if (mul_co) w = mul_bf16x8(w, pre[it]);          // dgrad * saved gelu' (aux chunks prefetched in this layout)
              if (do_colsum) acc_cs(w);
              if (out_mode == 1) w = add_bf16x8(w, pre[it]);   // gradient accumulation onto a residual-path gradient (old values prefetched)
              *reinterpret_cast<uint4*>(op + (long long)(it * 4) * p.ldo) = w;
            }
          } else {
#pragma unroll 1
            for (int it = 0; it < 8; ++it) {
              const int r = it * 4 + r_sub, row = row0 + r, nvalid = ncols - c_sub * 8;
              uint4 w = *reinterpret_cast<const uint4*>(stg + stage_off(r, c_sub));
              if (row < p.M && nvalid > 0) {
                if (do_colsum) {
                  if (nvalid < 8) {   // zero the lanes beyond N so the scalar tail below adds nothing for them
                    __nv_bfloat16* hw = reinterpret_cast<__nv_bfloat16*>(&w);
                    for (int j = nvalid; j < 8; ++j) hw[j] = __float2bfloat16_rn(0.f);
                  }
                  acc_cs(w);
                }
                if (out_mode == 1) w = add_bf16x8(w, load_chunk(reinterpret_cast<const __nv_bfloat16*>(p.out), p.ldo, out_vec_ok, row, col0 + c_sub * 8, nvalid));
                store_chunk(reinterpret_cast<__nv_bfloat16*>(p.out), p.ldo, out_vec_ok, row, col0 + c_sub * 8, nvalid, w);
              }
            }
          }
          if (do_colsum) {      // warp-uniform
            // rows of this warp's 32 x 64 slab: 8 per thread above, then across the 4 row sub-groups (lane bits 3, 4)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 8);
              cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 16);
            }
            if (r_sub == 0 && c_sub * 8 < ncols) {
              float* cp = p.colsum + col0 + c_sub * 8;
              if (c_sub * 8 + 8 <= ncols && (((uintptr_t)cp) & 15) == 0) {
                red_add_v4(cp, cs[0], cs[1], cs[2], cs[3]);
                red_add_v4(cp + 4, cs[4], cs[5], cs[6], cs[7]);
              } else {
                for (int j = 0; j < 8 && c_sub * 8 + j < ncols; ++j) atomicAdd(cp + j, cs[j]);
              }
            }
          }
          __syncwarp();
          if ((!aux_read || (mul_co && fast)) && pre_src != nullptr && gI + 1 < kGroups) issue_pre(gI + 1);   // RMW / multiply: operand tile of the next group
        }
      } else {
        // ---------------- fp32 output: groups of 32 columns (128-byte row segment) ----------------
#pragma unroll 1
        for (int gI = 0; gI < kColsPerWarp / 32; ++gI) {
          uint32_t r0[32];
          tmem_ld_32x32(t_row + gI * 32, r0);
          tmem_ld_wait();
          const int col0 = tn * BN + half * kColsPerWarp + gI * 32;
          if (col0 >= p.N) continue;
          const int ncols = min(32, p.N - col0);
          const bool fast = rows_full && ncols == 32 && out_vec_ok;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]);
          if (use_alpha) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
          }
          if (has_bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += bias_t[gI * 32 + j];
          }
          if (GEN || EPI == EPI_F32) {
            if (act == 1) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
            } else if (act == 2) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) *reinterpret_cast<float4*>(stg + stage_off(lane, c)) = make_float4(v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
          __syncwarp();
          if (fast) {
            float* op = reinterpret_cast<float*>(p.out) + (long long)(row0 + r_sub) * p.ldo + col0 + c_sub * 4;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              float4 w = *reinterpret_cast<const float4*>(stg + stage_off(it * 4 + r_sub, c_sub));
              float* o = op + (long long)(it * 4) * p.ldo;
              if (out_mode == 2) red_add_v4(o, w.x, w.y, w.z, w.w);
              else {
                if (out_mode == 1) { const float4 q = *reinterpret_cast<const float4*>(o); w.x += q.x; w.y += q.y; w.z += q.z; w.w += q.w; }
                *reinterpret_cast<float4*>(o) = w;
              }
            }
          } else {
#pragma unroll 1
            for (int it = 0; it < 8; ++it) {
              const int r = it * 4 + r_sub, row = row0 + r;
              if (row < p.M && c_sub * 4 < ncols) {
                const float4 w = *reinterpret_cast<const float4*>(stg + stage_off(r, c_sub));
                float* op = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col0 + c_sub * 4;
                if (out_vec_ok && c_sub * 4 + 4 <= ncols) {
                  if (out_mode == 2) red_add_v4(op, w.x, w.y, w.z, w.w);
                  else {
                    float4 q = w;
                    if (out_mode == 1) { const float4 o = *reinterpret_cast<const float4*>(op); q.x += o.x; q.y += o.y; q.z += o.z; q.w += o.w; }
                    *reinterpret_cast<float4*>(op) = q;
                  }
                } else {
                  const float ws[4] = {w.x, w.y, w.z, w.w};
                  for (int j = 0; j < 4 && c_sub * 4 + j < ncols; ++j) {
                    if (out_mode == 2) atomicAdd(op + j, ws[j]);
                    else op[j] = (out_mode == 1 ? op[j] : 0.f) + ws[j];
                  }
                }
              }
            }
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (!PAIR || leader) mbar_arrive(tempty_bar(as));
        else mbar_arrive_cluster(mapa_u32(tempty_bar(as), 0));   // the leader's MMA issuer owns the accumulator hand-off
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();     // the peer may still multicast into / arrive on this CTA's barriers until both are done
  tc_fence_after();
  if (warp == 1) {
    if (PAIR) tmem_dealloc_pair<2 * BN>(tmem_base);
    else tmem_dealloc<2 * BN>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// rows x cols bf16 matrix, cols contiguous, pitch `ld` elements; box = box_rows x 64 cols, 128B swizzle
static int make_tmap(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  HAMT_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  HAMT_REQUIRE(((uintptr_t)ptr & 15) == 0, "gemm operand base must be 16-byte aligned");
  HAMT_REQUIRE((ld * 2) % 16 == 0, "gemm operand pitch must be a multiple of 8 elements");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld);
    set_last_error(buf);
    return -2;
  }
  return 0;
}

static bool g_auto_pair = true;    // hamt_gemm_set_auto_pair(0) restricts the cost model to single-CTA tiles
void gemm_set_auto_pair(int on) { g_auto_pair = on != 0; }
static bool g_wide_epi = true;     // hamt_gemm_set_wide_epilogue(0): 8-warp epilogue also for the dGELU dgrad (A/B measurements)
void gemm_set_wide_epilogue(int on) { g_wide_epi = on != 0; }
static int g_num_sms = 0;
static int g_sm_limit = 0;        // hamt_gemm_set_sm_limit: persistent GEMM grids use at most this many SMs (0 = all)
void gemm_set_sm_limit(int n) { g_sm_limit = n > 0 ? n : 0; }
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  if (g_sm_limit > 0 && g_sm_limit < g_num_sms) return g_sm_limit < 2 ? 2 : g_sm_limit;
  return g_num_sms;
}

template <int BN, bool A_MN, bool B_MN, bool PAIR, int EPI, int EW = kEpiWarps>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  using L = SmemLayout<BN, PAIR>;
  constexpr int kThreads = 64 + 32 * EW;
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN, PAIR, EPI, EW>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); return -3; }
    attr_set = true;
  }
  const int units = p.tiles_m * p.tiles_n * p.splits;
  if (!PAIR) {
    const int grid = units < num_sms() ? units : num_sms();
    launch_pdl(kern, grid, kThreads, L::kTotal, st, ta, tb, p);
  } else {
    const int pairs = units < num_sms() / 2 ? units : num_sms() / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = L::kTotal; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    cudaLaunchKernelEx(&cfg, kern, ta, tb, p);
  }
  return check_launch("gemm_tcgen05_kernel");
}

int gemm_bf16(const GemmArgs& a, cudaStream_t st) {
  HAMT_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem");
  HAMT_REQUIRE(a.out_mode != 2 || a.out_f32, "gemm: split-K accumulation needs an fp32 output");
  HAMT_REQUIRE(a.aux_mode == 0 || !a.out_f32, "gemm: aux epilogues are only implemented for bf16 outputs");
  GemmParams p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.kb_total = (a.K + BK - 1) / BK;
  // Tile width / split-K selection by a small cost model (cycles on one SM; constants from the measured per-k-block
  // MMA time -- 4 x tcgen05.mma 128xBNx16: BN=256 is tensor-bound at 512 cyc, BN=128 is smem-operand-bound at ~330 cyc --
  // and the measured epilogue cost per tile).  waves * unit_time + exposed tail.
  // tile_n: 0 = auto, 128 / 256 = single-CTA tiles 128 x tile_n, 512 = CTA-pair tile 256 x 256 (cta_group::2)
  int bn = a.tile_n, splits = a.out_mode == 2 ? a.splits : 1;
  bool pair = false;
  {
    const int sms = num_sms();
    double best = 1e30;
    int best_bn = 128, best_s = 1;
    bool best_pair = false;
    for (int cand = 0; cand < 3; ++cand) {
      const int cbn = cand == 0 ? 256 : (cand == 1 ? 256 : 128);
      const bool cpair = cand == 0;
      const int code = cpair ? 512 : cbn;
      if (a.tile_n == 128 || a.tile_n == 256 || a.tile_n == 512) { if (code != a.tile_n) continue; }
      else {
        if (cbn == 256 && a.N < 192) continue;
        if (cpair && (a.M < 256 || !g_auto_pair)) continue;
      }
      const int tm = (a.M + (cpair ? 255 : 127)) / (cpair ? 256 : 128), tn = (a.N + cbn - 1) / cbn;
      const int slots = cpair ? sms / 2 : sms;
      // cycles per 64-deep k-block of one tile and per-tile epilogue cost: least-squares fit of this model to the measured K sweep
      // of all three tile modes (tools/gemm_sweep.py, profiles/r02_gemm_sweep_v1.txt; 90 points, M = 1024 ... 34560): the 128 x 128
      // tile is bound by the shared-memory operand bandwidth (2 x 16 KB per 256 MMA cycles) and runs at ~0.9-1.1 PF, 128 x 256 at
      // ~1.5 PF, the CTA pair at ~1.6 PF.  (The round-1 constants 340 / 600 / 540 under-estimated the 128 x 128 tile and sent the
      // K = 768 text-side GEMMs to it: 23 us instead of 17 us for 5120 x 2304 x 768.)
      const double t_kb = cpair ? 673.0 : (cbn == 256 ? 749.0 : 575.0);
      const double t_epi = (cbn == 256 ? 2367.0 : 1427.0) * (a.out_f32 ? 1.6 : 1.0);
      const bool forced = a.out_mode == 2 && a.splits > 0;
      for (int s = forced ? a.splits : 1; s <= (forced ? a.splits : 32); s *= 2) {
        if (a.out_mode != 2 && s > 1) break;
        const int kbs = (p.kb_total + s - 1) / s;
        if (!forced && s > 1 && kbs < 6) break;
        const int units = tm * tn * ((p.kb_total + kbs - 1) / kbs);
        const int waves = (units + slots - 1) / slots;
        const double unit = kbs * t_kb > t_epi ? kbs * t_kb : t_epi;
        const double cost = waves * unit + t_epi + (s > 1 ? 600.0 * waves : 0.0) + (cpair ? 1370.0 : 0.0);   // pair: cluster sync + remote hand-offs
        if (cost < best) { best = cost; best_bn = cbn; best_s = s; best_pair = cpair; }
      }
    }
    bn = best_bn;
    splits = best_s;
    pair = best_pair;
  }
  p.tiles_m = (a.M + (pair ? 2 * BM : BM) - 1) / (pair ? 2 * BM : BM);
  p.tiles_n = (a.N + bn - 1) / bn;
  if (splits > p.kb_total) splits = p.kb_total;
  if (splits < 1) splits = 1;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.m_fast = (p.tiles_m < p.tiles_n && p.tiles_m <= 16 && (long long)a.N * a.K * 2 > (32ll << 20)) ? 1 : 0;
  p.out = a.out; p.ldo = a.ldo; p.out_f32 = a.out_f32;
  // fp32 accumulation keeps the fire-and-forget red.global.add even without a K split: the read-modify-write epilogue waits for its loads
  // (tied MLM decoder wgrad, 30522 x 768 x 804: 124 us with RMW at splits = 1 against 76 us with red at splits = 2,
  // profiles/r02_kbench_decoder.txt)
  p.out_mode = a.out_mode;
  p.bias = a.bias; p.act = a.act; p.aux_mode = a.aux_mode; p.aux = (__nv_bfloat16*)a.aux; p.ld_aux = a.ld_aux;
  p.alpha = a.alpha;
  p.colsum = a.colsum;
  HAMT_REQUIRE(a.colsum == nullptr || (!a.out_f32 && a.out_mode == 0), "gemm: colsum needs a plainly stored bf16 output");
  HAMT_REQUIRE(p.aux_mode == 0 || p.aux != nullptr, "gemm: aux_mode set without aux buffer");
  HAMT_REQUIRE(p.aux_mode >= 0 && p.aux_mode <= 5, "gemm: unknown aux_mode");
  HAMT_REQUIRE(!((p.aux_mode == 2 || p.aux_mode == 3 || p.aux_mode == 5) && p.out_mode != 0), "gemm: the dGELU / dReLU / multiply epilogues store, they do not accumulate");
  HAMT_REQUIRE(p.aux_mode != 4 || p.act == 1, "gemm: aux_mode 4 stores gelu'(pre-activation) and needs act = gelu");

  CUtensorMap ta, tb;
  int rc;
  // K-major operand: matrix [rows = M or N, cols = K].  MN-major operand: matrix [rows = K, cols = M or N].
  rc = a.a_mn ? make_tmap(&ta, a.A, a.K, a.M, a.lda, BK) : make_tmap(&ta, a.A, a.M, a.K, a.lda, BM);
  if (rc) return rc;
  rc = a.b_mn ? make_tmap(&tb, a.B, a.K, a.N, a.ldb, BK) : make_tmap(&tb, a.B, a.N, a.K, a.ldb, pair ? bn / 2 : bn);
  if (rc) return rc;

  // epilogue specialisation: the combinations the transformer blocks run; everything else takes the generic kernel
  int epi = EPI_GENERIC;
  if (p.out_f32) { if (p.act == 0 && p.aux_mode == 0 && p.colsum == nullptr) epi = EPI_F32; }
  else if (p.alpha == 1.0f) {
    if (p.act == 0 && p.aux_mode == 0 && p.colsum == nullptr) epi = p.out_mode == 0 ? EPI_STORE : (p.out_mode == 1 ? EPI_ACCUM : EPI_GENERIC);
    else if (p.act == 1 && p.aux_mode == 1 && p.out_mode == 0 && p.colsum == nullptr) epi = EPI_GELU_PRE;
    else if (p.act == 1 && p.aux_mode == 4 && p.out_mode == 0 && p.colsum == nullptr) epi = EPI_GELU_DER;
    else if (p.act == 0 && p.aux_mode == 2 && p.out_mode == 0) epi = EPI_DGELU;
    else if (p.act == 0 && p.aux_mode == 5 && p.out_mode == 0) epi = EPI_MULAUX;
  }
  // 16-warp epilogue for the dGELU dgrad: fully aligned problems only (no guards in epilogue_wide).  (Measured for the GELU + derivative
  // forward too: 178.9 us against 178.3 us with 8 warps at M = 34 560 -- that epilogue is bound by its two output streams, not by issue slots.)
  if (g_wide_epi && bn == 256 && epi == EPI_DGELU) {
    const int tm_rows = pair ? 2 * BM : BM;
    const bool aligned = a.M % tm_rows == 0 && a.N % 256 == 0 && (((uintptr_t)p.out | (uintptr_t)p.aux | (uintptr_t)p.colsum | (uintptr_t)p.bias) & 15) == 0 &&
                         p.ldo % 8 == 0 && (p.aux == nullptr || p.ld_aux % 8 == 0);
    if (aligned) {
#define HAMT_WIDE(AMN_, BMN_, EPI_)                                                        \
  if (a.a_mn == AMN_ && a.b_mn == BMN_ && epi == EPI_) {                                   \
    if (pair) return launch<256, AMN_, BMN_, true, EPI_, 16>(ta, tb, p, st);               \
    return launch<256, AMN_, BMN_, false, EPI_, 16>(ta, tb, p, st);                        \
  }
      HAMT_WIDE(false, true, EPI_DGELU)
#undef HAMT_WIDE
    }
  }
#define HAMT_LAUNCH(BN_, AMN_, BMN_, PAIR_, EPI_) return launch<BN_, AMN_, BMN_, PAIR_, EPI_>(ta, tb, p, st);
#define HAMT_DISPATCH(BN_, PAIR_)                                                                    \
  if (!a.a_mn && !a.b_mn) {                                                                          \
    if (epi == EPI_STORE) HAMT_LAUNCH(BN_, false, false, PAIR_, EPI_STORE)                           \
    if (epi == EPI_GELU_PRE) HAMT_LAUNCH(BN_, false, false, PAIR_, EPI_GELU_PRE)                     \
    if (epi == EPI_GELU_DER) HAMT_LAUNCH(BN_, false, false, PAIR_, EPI_GELU_DER)                     \
    if (epi == EPI_F32) HAMT_LAUNCH(BN_, false, false, PAIR_, EPI_F32)                               \
    HAMT_LAUNCH(BN_, false, false, PAIR_, EPI_GENERIC)                                               \
  }                                                                                                  \
  if (!a.a_mn && a.b_mn) {                                                                           \
    if (epi == EPI_STORE) HAMT_LAUNCH(BN_, false, true, PAIR_, EPI_STORE)                            \
    if (epi == EPI_DGELU) HAMT_LAUNCH(BN_, false, true, PAIR_, EPI_DGELU)                            \
    if (epi == EPI_MULAUX) HAMT_LAUNCH(BN_, false, true, PAIR_, EPI_MULAUX)                          \
    if (epi == EPI_ACCUM) HAMT_LAUNCH(BN_, false, true, PAIR_, EPI_ACCUM)                            \
    HAMT_LAUNCH(BN_, false, true, PAIR_, EPI_GENERIC)                                                \
  }                                                                                                  \
  if (a.a_mn && a.b_mn) {                                                                            \
    if (epi == EPI_F32) HAMT_LAUNCH(BN_, true, true, PAIR_, EPI_F32)                                 \
    HAMT_LAUNCH(BN_, true, true, PAIR_, EPI_GENERIC)                                                 \
  }                                                                                                  \
  HAMT_LAUNCH(BN_, true, false, PAIR_, EPI_GENERIC)
  if (pair) { HAMT_DISPATCH(256, true) }
  if (bn == 256) { HAMT_DISPATCH(256, false) }
  HAMT_DISPATCH(128, false)
#undef HAMT_LAUNCH
#undef HAMT_DISPATCH
}

}  // namespace hamt
