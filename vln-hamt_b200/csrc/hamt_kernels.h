// Internal C++ launcher API shared by the .cu translation units; the public boundary is the
// C ABI in include/hamt_b200.h (implemented in hamt_abi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hamt {

struct GemmArgs {
  const void* A; int a_mn; long long lda;   // a_mn=0: A is [M,K] row-major; 1: A is [K,M] row-major
  const void* B; int b_mn; long long ldb;   // b_mn=0: B is [N,K] row-major; 1: B is [K,N] row-major
  void* out; long long ldo; int out_f32;
  int out_mode;                             // 0 store, 1 accumulate (RMW), 2 accumulate with split-K (atomics)
  int M, N, K;
  const float* bias; int act;               // act: 0 none, 1 gelu(erf), 2 relu
  int aux_mode; void* aux; long long ld_aux;  // 1 store pre-activation, 2 multiply by dgelu(aux), 3 multiply by (aux>0)
  float alpha;
  int tile_n;                               // 0 = auto, else 128 / 256
  int splits;                               // 0 = auto (only with out_mode 2)
  float* colsum;                            // fp32 [N] or null: += column sums of the stored bf16 output (fused bias gradient)
};
int gemm_bf16(const GemmArgs& a, cudaStream_t st);
void gemm_set_auto_pair(int on);
void gemm_set_sm_limit(int n);
void gemm_set_wide_epilogue(int on);

struct DropArgs { const unsigned long long* seed_ptr; unsigned int site; float p; };

// y = LN(drop(x) + res) * gamma + beta ; optionally stores z = drop(x)+res (bf16; may alias x) and row stats
// res32 (fp32, instead of res) / y32 (fp32 copy of y): the full-precision residual stream, either may be null
// z32 (fp32, may be null): the un-normalised sum as well -- the residual stream of PRE-LN blocks (ViT); x may be null (z = res32).
int ln_fwd(const void* x, const void* res, const float* res32, const float* gamma, const float* beta, void* y, float* y32, void* z_out,
           float* z32, float* mean, float* rstd, int M, int H, float eps, DropArgs drop, cudaStream_t st);
// backward of the above.  dx = grad wrt x (dropout applied), dres = grad wrt res (+ dres_in), column sums accumulated
// atomically into dgamma/dbeta/dbias (fp32, may be null).
// prenorm = 1 (ViT blocks): z is the residual stream itself, so dx = dropout-mask o (dz + dres_in) instead of dropout-mask o dz.
int ln_bwd(const void* dy, const void* z, const float* mean, const float* rstd, const float* gamma, const void* dres_in, void* dx,
           void* dres, float* dgamma, float* dbeta, float* dbias, int M, int H, DropArgs drop, int prenorm, cudaStream_t st);

struct AttnArgs {
  const void* q; const void* k; const void* v;       // bf16; element (b, s, h, d) at ptr + b*bstride + s*ld + h*64 + d
  long long q_bstride, kv_bstride; long long ldq, ldkv;
  const float* mask;                                  // additive fp32 [B, Sk] or null
  void* out; long long ldo; long long o_bstride;      // bf16 [B, Sq, heads*64]
  float* lse;                                         // fp32 [B, heads, Sq]
  int B, heads, Sq, Sk;
  float scale;
  DropArgs drop;
};
int attn_fwd(const AttnArgs& a, cudaStream_t st);
struct AttnBwdArgs {
  AttnArgs f;                 // forward description (q,k,v,mask,out(=saved ctx),lse)
  const void* dout; long long lddo; long long do_bstride;
  void* dq; void* dk; void* dv;                      // bf16, same layout/strides as q / k / v
  float* dbq; float* dbk; float* dbv;                // fp32 [heads*64] or null: += column sums of dq / dk / dv (fused bias gradients)
};
int attn_bwd(const AttnBwdArgs& a, cudaStream_t st);
// tcgen05 / TMA kernels (hamt_attn_tc.cu): return 1 when they took the problem (status in *rc), 0 -> run the legacy kernel
int attn_fwd_tc(const AttnArgs& a, cudaStream_t st, int* rc);
int attn_bwd_tc(const AttnBwdArgs& a, cudaStream_t st, int* rc);
void attn_set_impl(int v);      // 0 = auto, 1 = legacy mma.sync kernels only

// text embedding: out = drop(LN(word[ids] + pos[s] + type0))
int embed_text_fwd(const long long* ids, const float* word, const float* pos, const float* type0, const float* gamma, const float* beta,
                   void* out, int B, int L, int H, float eps, DropArgs drop, cudaStream_t st);
int embed_text_bwd(const void* dy, const long long* ids, const float* word, const float* pos, const float* type0, const float* gamma,
                   float* dword, float* dpos, float* dtype0, float* dgamma, float* dbeta, int B, int L, int H, float eps, DropArgs drop,
                   cudaStream_t st);

// feature-token embedding (ImageEmbeddings / HistoryEmbeddings / pano tokens):
//   s = LN_img(t) + LN_ang(ang @ Wang^T + bang) [+ add_vec] [+ nav_table[nav_ids]] [+ extra] [+ pos_table[pos_ids]]
//   out = final_ln ? drop(LN_f(s)) : drop(s)
struct EmbedFeatArgs {
  const void* t;                  // bf16 [M,H]  (image linear output incl. bias)
  const float* ang;               // fp32 [M,A]
  int A;                          // angle feature size (<= 8)
  const float* w_ang; const float* b_ang;            // [H,A], [H]
  const float* g_img; const float* b_img;            // LN_img
  const float* g_ang; const float* be_ang;           // LN_ang
  const float* add_vec;                              // [H] or null
  const float* nav_table; const long long* nav_ids;  // [*,H], [M] or null
  const float* extra;                                // fp32 [M,H] added to the sum (pano mean) or null
  const float* pos_table; const long long* pos_ids; int pos_mod;  // pos id of row r = pos_ids ? pos_ids[r] : (r % pos_mod); null table = unused
  const float* g_f; const float* b_f;                // final LN or null
  void* out;                                         // bf16 [M,H]
  int M, H; float eps; DropArgs drop;
};
int embed_feat_fwd(const EmbedFeatArgs& a, cudaStream_t st);
struct EmbedFeatBwdArgs {
  EmbedFeatArgs f;
  const void* dy;                 // bf16 [M,H]
  void* dt;                       // bf16 [M,H]
  float* dw_ang; float* db_ang; float* dg_img; float* db_img; float* dg_ang; float* dbe_ang; float* dadd_vec; float* dnav_table;
  float* dextra;                  // fp32 [M,H] (written, not accumulated) or null
  float* dpos_table; float* dg_f; float* db_f;
  float* db_lin;                  // bias grad of the image linear (column sum of dt) or null
};
int embed_feat_bwd(const EmbedFeatBwdArgs& a, cudaStream_t st);

// end-to-end ViT stage (hamt_vit.cu)
int patchify_bf16(const float* img, void* out, int N, int C, int Hh, int Ww, int ps, cudaStream_t st);   // fp32 NCHW -> bf16 [N*gh*gw, C*ps*ps]
int vit_embed_fwd(const void* t0, const float* cls, const float* pos, float* x32, void* x16, int N, int S, int H, DropArgs drop, cudaStream_t st);
int vit_embed_bwd(const void* dx, void* dfull, void* dt0, int N, int S, int H, DropArgs drop, cudaStream_t st);

// misc bandwidth kernels
int cast_f32_to_bf16(const float* in, void* out, long long n, cudaStream_t st);
int colsum_bf16(const void* x, long long ld, float* out, int M, int N, cudaStream_t st);          // out[n] += sum_m x[m,n]
int mean_pool_fwd(const void* x, float* out, int N, int P, int H, cudaStream_t st);                // bf16 [N,P,H] -> fp32 [N,H]
int mean_pool_bwd(const float* dy, void* dx, int N, int P, int H, cudaStream_t st);                // fp32 [N,H] -> bf16 [N,P,H]
int add_bf16(const void* a, const void* b, void* out, long long n, cudaStream_t st);
int mul_rows_bf16(const void* a, const void* b, void* out, int B, int S, int H, cudaStream_t st);  // out[b,s,:] = a[b,s,:]*b[b,:]

// fused AdamW (+ global-norm clip, bf16 shadow refresh, gradient zeroing) over the flat arena; hamt_optim.cu
struct AdamWArgs {
  float* param; float* grad; float* exp_avg; float* exp_avg_sq; void* shadow;   // flat [total]; shadow (bf16) may be null
  long long total;                         // elements, multiple of 64
  const int* chunk_seg;                    // [total / 64] segment id of every 64-element chunk, -1 = padding
  const long long* seg_end;                // [nseg] one past the last element of the parameter (its alignment tail is not part of it)
  int nseg;
  const unsigned char* seg_active;         // [nseg] 1 = the parameter has a gradient this step
  const float* seg_wd;                     // [nseg] weight decay of the parameter's group
  int* seg_step;                           // [nseg] per-parameter step counters (state["step"]), updated
  float* seg_step_size;                    // [nseg] scratch
  const float* lr;                         // device scalar: this step's learning rate
  double beta1, beta2, eps;
  int correct_bias;
  float max_grad_norm;                     // <= 0: no clipping
  int want_norm;                           // compute the global gradient norm even without clipping
  int zero_grad;                           // zero the gradients of the active segments after use
  float* workspace;                        // [adamw_workspace_floats()]: [0] = grad norm, [1] = clip coefficient, rest scratch
};
int adamw_workspace_floats(void);
int adamw_step(const AdamWArgs& a, cudaStream_t st);

}  // namespace hamt
