// Shared device helpers for the HAMT sm_100a kernels: mbarrier / TMA / tcgen05 PTX wrappers,
// bf16 packing, warp reductions, counter-hash dropout RNG.  Hand-written inline PTX; no CUTLASS.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace hamt {

// ---------------------------------------------------------------------------------------------
// error plumbing shared by every translation unit (defined in hamt_abi.cu)
// ---------------------------------------------------------------------------------------------
void set_last_error(const char* msg);
int check_launch(const char* what);   // returns 0 or a negative status; records cudaGetLastError text

#define HAMT_REQUIRE(cond, msg)                 \
  do {                                          \
    if (!(cond)) {                              \
      ::hamt::set_last_error(msg);              \
      return -1;                                \
    }                                           \
  } while (0)

// ---------------------------------------------------------------------------------------------
// programmatic dependent launch (PDL): every kernel of the library is launched with the
// programmatic-stream-serialization attribute and starts with pdl_grid_sync(), so the launch latency /
// CTA scheduling / prologue of kernel i+1 overlaps the tail of kernel i (also inside captured CUDA graphs).
// griddepcontrol.wait blocks until the preceding grid has completed and flushed its memory, so placing it
// before the first global-memory access keeps the usual stream-order semantics.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_grid_sync() { pdl_wait(); pdl_trigger(); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
// small math
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(h);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// erf-based GELU (reference: pretrain_src/model/vilmodel.py:23-29, x * 0.5 * (1 + erf(x / sqrt(2)))).
// The GEMM epilogues that apply it are ALU-bound (128 x 256 outputs per tile on 8 warps while the tensor pipe needs only ~6 k
// cycles per tile), so the Gaussian cdf is evaluated with ONE special-function op:  Phi(-a) = 2^R(a), a = |x|, R = degree-5
// minimax polynomial with R(0) = -1 (Phi(0) = 0.5 exactly), fitted on [0, 6] against scipy log_ndtr: |Phi error| <= 1.5e-5,
// |gelu error| <= 4.3e-6 (1/500 of a bf16 half-ulp at |x| ~ 1); the leading coefficient is negative so the tail underflows
// to 0 for large |x| instead of oscillating.  libdevice erff() (and the rcp + ex2 Abramowitz-Stegun 7.1.26 form used before) made
// the GELU / dGELU epilogues 2x longer than the tensor-core main loop.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gauss_tail(float a) {      // Phi(-a) for a >= 0
  float r = fmaf(-0.0004371631075628102f, a, 0.006833082064986229f);
  r = fmaf(r, a, -0.05121844261884689f);
  r = fmaf(r, a, -0.46058332920074463f);
  r = fmaf(r, a, -1.1506280899047852f);
  r = fmaf(r, a, -1.0f);
  return ex2_approx(r);
}
// x * Phi(x) = relu(x) - |x| * Phi(-|x|)
__device__ __forceinline__ float gelu_erf(float x) {
  const float a = fabsf(x);
  return fmaf(-a, gauss_tail(a), fmaxf(x, 0.f));
}
// d/dx [x * Phi(x)] = Phi(x) + x * phi(x),  phi(x) = exp(-x^2 / 2) / sqrt(2 pi)
__device__ __forceinline__ float dgelu_erf(float x) {
  const float p = gauss_tail(fabsf(x));
  const float cdf = x >= 0.f ? 1.0f - p : p;
  const float e = ex2_approx(x * x * -0.72134752044448170368f);
  return fmaf(x * 0.39894228040143267794f, e, cdf);
}

// gelu(x) and gelu'(x) together: one Gaussian tail (ex2) shared by both, one more ex2 for the density
__device__ __forceinline__ void gelu_and_derivative(float x, float& g, float& d) {
  const float a = fabsf(x);
  const float p = gauss_tail(a);
  g = fmaf(-a, p, fmaxf(x, 0.f));
  const float cdf = x >= 0.f ? 1.0f - p : p;
  const float e = ex2_approx(x * x * -0.72134752044448170368f);
  d = fmaf(x * 0.39894228040143267794f, e, cdf);
}

// ---------------------------------------------------------------------------------------------
// dropout: stateless counter hash.  keep(idx) is a pure function of (seed, site, idx) so forward
// and backward regenerate the identical mask without storing it.  `seed` lives in device memory
// (so a captured CUDA graph can be replayed with a new seed); `site` distinguishes call sites.
// ---------------------------------------------------------------------------------------------
struct DropCfg {
  const unsigned long long* seed_ptr;  // device pointer, may be null when p == 0
  uint32_t site;                       // unique per dropout call-site within a step
  float p;                             // drop probability; 0 disables
};

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
struct DropState {
  uint32_t k0, k1, thresh;
  float scale;
  bool on;
};
__device__ __forceinline__ DropState drop_init(const DropCfg& c) {
  DropState s;
  s.on = c.p > 0.f;
  s.k0 = s.k1 = 0; s.thresh = 0; s.scale = 1.f;
  if (s.on) {
    unsigned long long seed = *c.seed_ptr;
    s.k0 = mix32((uint32_t)seed ^ (c.site * 0x9E3779B1u));
    s.k1 = mix32((uint32_t)(seed >> 32) + c.site * 0x85EBCA77u + 0x165667B1u);
    // drop when hash < p * 2^32
    s.thresh = (uint32_t)fminf(c.p * 4294967296.0f, 4294967040.0f);
    s.scale = 1.0f / (1.0f - c.p);
  }
  return s;
}
// returns the multiplier for element idx: 0 (dropped) or 1/(1-p)
__device__ __forceinline__ float drop_mult(const DropState& s, unsigned long long idx) {
  if (!s.on) return 1.f;
  uint32_t h = mix32(((uint32_t)idx ^ s.k0) * 0x9E3779B1u + (uint32_t)(idx >> 32) * 0xC2B2AE3Du + s.k1);
  return h < s.thresh ? 0.f : s.scale;
}

// Dropout on attention probabilities: one hash per PAIR of adjacent keys of a row, 16 bits each (drop when the 16-bit lane is below
// p * 65536: p = 0.1 -> 0.100006).  The per-element hash above costs ~14 integer instructions -- more than the rest of the softmax
// arithmetic of an element -- and the probabilities are the most numerous dropout site of the model (12 heads x S^2 per sequence).
// All four attention kernels (tcgen05 and legacy, forward and backward) use this function, so any forward / backward combination
// regenerates the same mask.  row_id = (sequence * heads + head) * Sq + query.
struct AttnDrop {
  uint32_t k0, k1, thresh16;
  float scale;
  bool on;
};
__device__ __forceinline__ AttnDrop attn_drop_init(const DropCfg& c) {
  AttnDrop s;
  s.on = c.p > 0.f;
  s.k0 = s.k1 = 0; s.thresh16 = 0; s.scale = 1.f;
  if (s.on) {
    unsigned long long seed = *c.seed_ptr;
    s.k0 = mix32((uint32_t)seed ^ (c.site * 0x9E3779B1u));
    s.k1 = mix32((uint32_t)(seed >> 32) + c.site * 0x85EBCA77u + 0x165667B1u);
    s.thresh16 = (uint32_t)fminf(c.p * 65536.0f + 0.5f, 65535.0f);
    s.scale = 1.0f / (1.0f - c.p);
  }
  return s;
}
__device__ __forceinline__ uint32_t attn_drop_rowkey(const AttnDrop& s, unsigned long long row_id) {
  return mix32(((uint32_t)row_id ^ s.k0) * 0x9E3779B1u + (uint32_t)(row_id >> 32) * 0xC2B2AE3Du + s.k1);
}
// 2 x 16 random bits for keys (2 * pair, 2 * pair + 1) of the row
__device__ __forceinline__ uint32_t attn_drop_bits(uint32_t rowkey, uint32_t pair) { return mix32(rowkey + pair * 0x9E3779B1u); }
// multiplier (0 or 1/(1-p)) of key `key` of the row
__device__ __forceinline__ float attn_drop_mult(const AttnDrop& s, uint32_t rowkey, int key) {
  if (!s.on) return 1.f;
  const uint32_t b = attn_drop_bits(rowkey, (uint32_t)key >> 1);
  const uint32_t lane16 = (key & 1) ? (b >> 16) : (b & 0xffffu);
  return lane16 < s.thresh16 ? 0.f : s.scale;
}

// ---------------------------------------------------------------------------------------------
// shared-memory addressing, mbarrier, TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// non-blocking probe (try_wait may suspend the thread for a hardware-defined time; a thread that polls several barriers uses this)
__device__ __forceinline__ uint32_t mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug traps after ~2 s instead of hanging the GPU.  The timeout path is a separate function so that the
// dozens of wait sites of a kernel do not each carry a timer loop and a printf call (instruction-cache footprint).
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  unsigned long long t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ffu) == 0 && globaltimer_ns() - t0 > 2000000000ull) {
      printf("hamt: mbarrier wait timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity);
}

__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// 2D tiled load: coordinates (c0 = innermost/contiguous element index, c1 = row)
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr) {
  static_assert(kCols == 32 || kCols == 64 || kCols == 128 || kCols == 256 || kCols == 512, "TMEM cols: pow2 >= 32");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate.  One thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// TMEM -> registers: lane i of the warp reads 32 consecutive fp32 columns of TMEM lane (base_lane + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) variants: two CTAs of a cluster (same TPC) execute ONE 256-row UMMA; each CTA stages its own 128
// rows of A and HALF of the B tile, so the L2 -> SM operand traffic per flop halves compared with two independent 128-row CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of the pair; the bytes are counted on the mbarrier at `bar_cluster_addr` (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t smem_dst, const void* tmap, int c0, int c1, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(tmap), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_result_addr) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
// commit of the pair's MMAs: one arrival on the mbarrier at this smem offset in EVERY CTA of `mask`
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}

// UMMA shared-memory matrix descriptor (sm_100 format): start address [0,14) (>>4), leading byte
// offset [16,30) (>>4), stride byte offset [32,46) (>>4), version=1 at [46,48), layout type at
// [61,64) (2 = 128-byte swizzle).  Tiles are 1024-byte aligned so base_offset = 0.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// UMMA instruction descriptor, kind::f16: D=f32 (bits[4,6)=1), A=B=bf16 (bits[7,10)=1,[10,13)=1),
// a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

}  // namespace hamt
