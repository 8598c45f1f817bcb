// C ABI of libhamt_b200.so (declared in include/hamt_b200.h): thin extern "C" shims over the
// launchers, plus the error / launch-count plumbing.
#include <atomic>
#include <string.h>
#include <stdio.h>
#include "hamt_common.cuh"
#include "hamt_kernels.h"
#include "../../include/hamt_b200.h"

namespace hamt {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_last_error(const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    char buf[400];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    set_last_error(buf);
    return -10;
  }
  return 0;
}

}  // namespace hamt

using namespace hamt;

extern "C" {

int hamt_abi_version(void) { return HAMT_ABI_VERSION; }
const char* hamt_last_error(void) { return g_err; }
long long hamt_launch_count(void) { return g_launches.load(); }

int hamt_gemm_bf16(const void* A, int a_mn, long long lda, const void* B, int b_mn, long long ldb, void* out, long long ldo, int out_f32,
                   int out_mode, int M, int N, int K, const float* bias, int act, int aux_mode, void* aux, long long ld_aux, float alpha,
                   int tile_n, int splits, float* colsum, void* stream) {
  GemmArgs a{A, a_mn, lda, B, b_mn, ldb, out, ldo, out_f32, out_mode, M, N, K, bias, act, aux_mode, aux, ld_aux, alpha, tile_n, splits, colsum};
  return gemm_bf16(a, (cudaStream_t)stream);
}

int hamt_gemm_set_auto_pair(int on) { gemm_set_auto_pair(on); return 0; }

int hamt_ln_fwd(const void* x, const void* res, const float* res32, const float* gamma, const float* beta, void* y, float* y32, void* z_out,
                float* mean, float* rstd, int M, int H, float eps, const unsigned long long* seed_ptr, unsigned int site, float p, void* stream) {
  return ln_fwd(x, res, res32, gamma, beta, y, y32, z_out, nullptr, mean, rstd, M, H, eps, DropArgs{seed_ptr, site, p}, (cudaStream_t)stream);
}
int hamt_ln_fwd_prenorm(const void* x, const float* res32, const float* gamma, const float* beta, void* y, float* y32, void* z_out, float* z32,
                        float* mean, float* rstd, int M, int H, float eps, const unsigned long long* seed_ptr, unsigned int site, float p,
                        void* stream) {
  return ln_fwd(x, nullptr, res32, gamma, beta, y, y32, z_out, z32, mean, rstd, M, H, eps, DropArgs{seed_ptr, site, p}, (cudaStream_t)stream);
}
int hamt_patchify_bf16(const float* images, void* out, int N, int C, int H, int W, int patch, void* stream) {
  return patchify_bf16(images, out, N, C, H, W, patch, (cudaStream_t)stream);
}
int hamt_vit_embed_fwd(const void* t0, const float* cls, const float* pos, float* x32, void* x16, int N, int S, int H,
                       const unsigned long long* seed_ptr, unsigned int site, float p, void* stream) {
  return vit_embed_fwd(t0, cls, pos, x32, x16, N, S, H, DropArgs{seed_ptr, site, p}, (cudaStream_t)stream);
}
int hamt_vit_embed_bwd(const void* dx, void* dfull, void* dt0, int N, int S, int H, const unsigned long long* seed_ptr, unsigned int site,
                       float p, void* stream) {
  return vit_embed_bwd(dx, dfull, dt0, N, S, H, DropArgs{seed_ptr, site, p}, (cudaStream_t)stream);
}
int hamt_ln_bwd(const void* dy, const void* z, const float* mean, const float* rstd, const float* gamma, const void* dres_in, void* dx, void* dres,
                float* dgamma, float* dbeta, float* dbias, int M, int H, const unsigned long long* seed_ptr, unsigned int site, float p,
                void* stream) {
  return ln_bwd(dy, z, mean, rstd, gamma, dres_in, dx, dres, dgamma, dbeta, dbias, M, H, DropArgs{seed_ptr, site, p}, 0, (cudaStream_t)stream);
}
int hamt_ln_bwd_prenorm(const void* dy, const void* z, const float* mean, const float* rstd, const float* gamma, const void* dres_in, void* dx,
                        void* dres, float* dgamma, float* dbeta, float* dbias, int M, int H, const unsigned long long* seed_ptr, unsigned int site,
                        float p, void* stream) {
  return ln_bwd(dy, z, mean, rstd, gamma, dres_in, dx, dres, dgamma, dbeta, dbias, M, H, DropArgs{seed_ptr, site, p}, 1, (cudaStream_t)stream);
}

int hamt_attn_fwd(const void* q, const void* k, const void* v, long long q_bstride, long long kv_bstride, long long ldq, long long ldkv,
                  const float* mask, void* out, long long ldo, long long o_bstride, float* lse, int B, int heads, int Sq, int Sk, float scale,
                  const unsigned long long* seed_ptr, unsigned int site, float p, void* stream) {
  AttnArgs a{q, k, v, q_bstride, kv_bstride, ldq, ldkv, mask, out, ldo, o_bstride, lse, B, heads, Sq, Sk, scale, DropArgs{seed_ptr, site, p}};
  return attn_fwd(a, (cudaStream_t)stream);
}
int hamt_attn_bwd(const void* q, const void* k, const void* v, long long q_bstride, long long kv_bstride, long long ldq, long long ldkv,
                  const float* mask, const void* out, long long ldo, long long o_bstride, const float* lse, const void* dout, long long lddo,
                  long long do_bstride, void* dq, void* dk, void* dv, int B, int heads, int Sq, int Sk, float scale,
                  const unsigned long long* seed_ptr, unsigned int site, float p, float* dbias_q, float* dbias_k, float* dbias_v, void* stream) {
  AttnBwdArgs a{};
  a.f = AttnArgs{q, k, v, q_bstride, kv_bstride, ldq, ldkv, mask, const_cast<void*>(out), ldo, o_bstride, const_cast<float*>(lse),
                 B, heads, Sq, Sk, scale, DropArgs{seed_ptr, site, p}};
  a.dout = dout; a.lddo = lddo; a.do_bstride = do_bstride; a.dq = dq; a.dk = dk; a.dv = dv;
  a.dbq = dbias_q; a.dbk = dbias_k; a.dbv = dbias_v;
  return attn_bwd(a, (cudaStream_t)stream);
}

int hamt_embed_text_fwd(const long long* ids, const float* word, const float* pos, const float* type0, const float* gamma, const float* beta,
                        void* out, int B, int L, int H, float eps, const unsigned long long* seed_ptr, unsigned int site, float p, void* stream) {
  return embed_text_fwd(ids, word, pos, type0, gamma, beta, out, B, L, H, eps, DropArgs{seed_ptr, site, p}, (cudaStream_t)stream);
}
int hamt_embed_text_bwd(const void* dy, const long long* ids, const float* word, const float* pos, const float* type0, const float* gamma,
                        float* dword, float* dpos, float* dtype0, float* dgamma, float* dbeta, int B, int L, int H, float eps,
                        const unsigned long long* seed_ptr, unsigned int site, float p, void* stream) {
  return embed_text_bwd(dy, ids, word, pos, type0, gamma, dword, dpos, dtype0, dgamma, dbeta, B, L, H, eps, DropArgs{seed_ptr, site, p},
                        (cudaStream_t)stream);
}

static EmbedFeatArgs to_args(const hamt_embed_feat_desc* d) {
  EmbedFeatArgs a{};
  a.t = d->t; a.ang = d->ang; a.A = d->A; a.w_ang = d->w_ang; a.b_ang = d->b_ang; a.g_img = d->g_img; a.b_img = d->b_img; a.g_ang = d->g_ang;
  a.be_ang = d->be_ang; a.add_vec = d->add_vec; a.nav_table = d->nav_table; a.nav_ids = d->nav_ids; a.extra = d->extra;
  a.pos_table = d->pos_table; a.pos_ids = d->pos_ids; a.pos_mod = d->pos_mod; a.g_f = d->g_f; a.b_f = d->b_f; a.out = d->out; a.M = d->M;
  a.H = d->H; a.eps = d->eps; a.drop = DropArgs{d->seed_ptr, d->site, d->p};
  return a;
}
int hamt_embed_feat_fwd(const hamt_embed_feat_desc* d, void* stream) {
  if (!d) { set_last_error("embed_feat_fwd: null descriptor"); return -1; }
  return embed_feat_fwd(to_args(d), (cudaStream_t)stream);
}
int hamt_embed_feat_bwd(const hamt_embed_feat_desc* d, const hamt_embed_feat_grads* g, void* stream) {
  if (!d || !g) { set_last_error("embed_feat_bwd: null descriptor"); return -1; }
  EmbedFeatBwdArgs a{};
  a.f = to_args(d);
  a.dy = g->dy; a.dt = g->dt; a.dw_ang = g->dw_ang; a.db_ang = g->db_ang; a.dg_img = g->dg_img; a.db_img = g->db_img; a.dg_ang = g->dg_ang;
  a.dbe_ang = g->dbe_ang; a.dadd_vec = g->dadd_vec; a.dnav_table = g->dnav_table; a.dextra = g->dextra; a.dpos_table = g->dpos_table;
  a.dg_f = g->dg_f; a.db_f = g->db_f; a.db_lin = g->db_lin;
  return embed_feat_bwd(a, (cudaStream_t)stream);
}

int hamt_cast_f32_to_bf16(const float* in, void* out, long long n, void* stream) { return cast_f32_to_bf16(in, out, n, (cudaStream_t)stream); }
int hamt_colsum_bf16(const void* x, long long ld, float* out, int M, int N, void* stream) { return colsum_bf16(x, ld, out, M, N, (cudaStream_t)stream); }
int hamt_mean_pool_fwd(const void* x, float* out, int N, int P, int H, void* stream) { return mean_pool_fwd(x, out, N, P, H, (cudaStream_t)stream); }
int hamt_mean_pool_bwd(const float* dy, void* dx, int N, int P, int H, void* stream) { return mean_pool_bwd(dy, dx, N, P, H, (cudaStream_t)stream); }
int hamt_add_bf16(const void* a, const void* b, void* out, long long n, void* stream) { return add_bf16(a, b, out, n, (cudaStream_t)stream); }
int hamt_mul_rows_bf16(const void* a, const void* v, void* out, int B, int S, int H, void* stream) {
  return mul_rows_bf16(a, v, out, B, S, H, (cudaStream_t)stream);
}

int hamt_gemm_set_sm_limit(int n) { gemm_set_sm_limit(n); return 0; }
int hamt_gemm_set_wide_epilogue(int on) { gemm_set_wide_epilogue(on); return 0; }
int hamt_attn_set_impl(int v) { attn_set_impl(v); return 0; }
int hamt_adamw_workspace_floats(void) { return adamw_workspace_floats(); }
int hamt_adamw_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16, long long total, const int* chunk_seg, const long long* seg_end, int nseg,
                    const unsigned char* seg_active, const float* seg_wd, int* seg_step, float* seg_step_size, const float* lr, double beta1,
                    double beta2, double eps, int correct_bias, float max_grad_norm, int want_norm, int zero_grad, float* workspace, void* stream) {
  AdamWArgs a{param, grad, exp_avg, exp_avg_sq, shadow_bf16, total, chunk_seg, seg_end, nseg, seg_active, seg_wd, seg_step, seg_step_size, lr,
              beta1, beta2, eps, correct_bias, max_grad_norm, want_norm, zero_grad, workspace};
  return adamw_step(a, (cudaStream_t)stream);
}

}  // extern "C"
