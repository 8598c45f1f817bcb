/* hamt_b200.h -- C ABI of libhamt_b200.so: the sm_100a kernels behind the HAMT hot path.
 *
 * The reference (cshizhe/VLN-HAMT) has no native code and no FFI; its hot path is the
 * torch-eager op sequence of pretrain_src/model/vilmodel.py / pretrain_cmt.py and
 * finetune_src/models/vilmodel_cmt.py.  Each entry point below replaces one group of those eager ops
 * (cited per function).  The drop-in boundary a reference maintainer binds to is this header; the
 * ctypes binding used by this repo is vln-hamt_b200/_lib.py and the module-level mirror of the
 * reference API is vln-hamt_b200/{vilmodel,pretrain_cmt,vilmodel_cmt,model_HAMT}.py.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (PyTorch allocator); kernels never allocate;
 *  - activations are bf16 row-major, parameters / statistics / gradients of parameters are fp32;
 *  - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised;
 *  - return 0 on success, negative on invalid arguments or launch failure; hamt_last_error() gives the
 *    text (thread-local);  no C++ exceptions cross the boundary;
 *  - dropout: `seed_ptr` -> one uint64 in device memory (so captured CUDA graphs can be replayed with a
 *    fresh seed), `site` = unique id of the call site within a step, `p` = drop probability (0 = off).
 *    Forward and backward regenerate the identical mask from (seed, site, element index).
 */
#ifndef HAMT_B200_H
#define HAMT_B200_H
#ifdef __cplusplus
extern "C" {
#endif

#define HAMT_ABI_VERSION 2

int hamt_abi_version(void);
const char* hamt_last_error(void);
/* number of kernels launched by this library in the calling process (bench.py gpu_launches) */
long long hamt_launch_count(void);

/* D[M,N] = act(alpha * sum_k A(m,k) B(n,k) + bias[n])   tcgen05 GEMM, bf16 in, fp32 accumulate.
 * a_mn = 0: A stored [M,K] (pitch lda); 1: A stored [K,M].  b_mn likewise for B ([N,K] / [K,N]).
 *   forward  y = x W^T           : A = x (a_mn 0), B = W [N,K] (b_mn 0)          vilmodel.py:96-98,140,169,182
 *   dgrad    dx = dy W           : A = dy (a_mn 0), B = W [N,K] read as [K=N,N=K] (b_mn 1)
 *   wgrad    dW = dy^T x         : A = dy (a_mn 1), B = x (b_mn 1), fp32 out, out_mode 2
 * out_f32: 0 bf16 / 1 fp32 output.  out_mode: 0 store, 1 out += , 2 out += with split-K atomics.
 * act: 0 none, 1 exact-erf GELU (vilmodel.py:23-29), 2 ReLU (pretrain_cmt.py:16).
 * aux_mode: 0 none, 1 also store the pre-activation (bf16) to aux, 2 multiply by dGELU(aux), 3 by (aux > 0),
 *           4 (with act = gelu) store gelu'(pre-activation) to aux instead -- BertIntermediate forward, vilmodel.py:168-171 -- so that
 *           5 multiply by aux is the whole BertOutput-dgrad epilogue (the erf-GELU derivative is evaluated once, in the forward).
 * tile_n: 0 auto / 128 / 256 (one CTA per 128 x tile_n tile) / 512 (CTA pair, 256 x 256 tile, tcgen05 cta_group::2).  splits: 0 auto.
 * colsum: fp32 [N] or null; the epilogue ACCUMULATES the column sums of the stored bf16 output into it -- the bias gradient of
 *   the Linear that produced the activation whose gradient this GEMM writes (replaces a separate pass over [M,N]); bf16 store only. */
int hamt_gemm_bf16(const void* A, int a_mn, long long lda, const void* B, int b_mn, long long ldb, void* out, long long ldo, int out_f32,
                   int out_mode, int M, int N, int K, const float* bias, int act, int aux_mode, void* aux, long long ld_aux, float alpha,
                   int tile_n, int splits, float* colsum, void* stream);

/* tile_n == 0: the tile cost model may choose the CTA-pair kernel (default on; 0 restricts it to single-CTA tiles) */
int hamt_gemm_set_auto_pair(int on);
/* The persistent GEMM grids (one CTA per SM, ~200 KB of shared memory each: nothing else can co-reside) use at most n SMs; 0 = all.
 * Data-parallel runs set 148 - R during the backward pass so that the R CTAs of the NCCL all-reduce kernel that overlaps it (dp.py;
 * the reference overlaps its DDP buckets the same way, utils/misc.py:52-65) find free SMs instead of waiting for a GEMM to retire. */
int hamt_gemm_set_sm_limit(int n);
/* 16-warp epilogue for the ALU-bound dGELU dgrad (BertOutput backward: acc * gelu'(pre) + fused bias column sums) on fully aligned
 * 256-wide tiles (hamt_gemm.cu: epilogue_wide): 270 -> 183 us at M = 34 560 (profiles/r02_kbench_variants.txt).  On by default; 0 selects
 * the 8-warp epilogue (bit-identical results, kept for A/B measurements).  The same wide epilogue LOST on the store / GELU / accumulate
 * epilogues (157 -> 167 us for GELU) and is not instantiated for them. */
int hamt_gemm_set_wide_epilogue(int on);

/* y = LayerNorm(dropout(x) + res) ; BertSelfOutput / BertOutput tail (vilmodel.py:139-143,181-185).
 * z_out (may alias x, may be null) receives dropout(x)+res in bf16; mean/rstd fp32 [M] (may be null). */
int hamt_ln_fwd(const void* x, const void* res, const float* res32, const float* gamma, const float* beta, void* y, float* y32, void* z_out,
                float* mean, float* rstd, int M, int H, float eps, const unsigned long long* seed_ptr, unsigned int site, float p, void* stream);
/* Pre-LN variant for the end-to-end ViT stage (Block.forward, pretrain_src/model/vision_transformer.py:195-198:
 * x = x + sublayer(norm(x))): z = dropout(x) + res32 is the NEW residual stream (written in fp32 to z32 and in bf16 to z_out for the
 * backward), y = LayerNorm_next(z) the input of the next sublayer.  x may be null (z = res32: the first norm of the backbone). */
int hamt_ln_fwd_prenorm(const void* x, const float* res32, const float* gamma, const float* beta, void* y, float* y32, void* z_out, float* z32,
                        float* mean, float* rstd, int M, int H, float eps, const unsigned long long* seed_ptr, unsigned int site, float p,
                        void* stream);
/* backward of hamt_ln_fwd_prenorm: like hamt_ln_bwd, but z is the residual STREAM (it is consumed again by every later sublayer), so the
 * gradient of the sublayer output is the total gradient of z: dx = dropout-mask o (dz_LayerNorm + dres_in), dres = dz_LayerNorm + dres_in
 * (dres is required; dres_in null = no later consumer: the last sublayer). */
int hamt_ln_bwd_prenorm(const void* dy, const void* z, const float* mean, const float* rstd, const float* gamma, const void* dres_in, void* dx,
                        void* dres, float* dgamma, float* dbeta, float* dbias, int M, int H, const unsigned long long* seed_ptr, unsigned int site,
                        float p, void* stream);
/* PatchEmbed (vision_transformer.py:201-223): Conv2d(C, E, kernel = stride = patch) == GEMM over non-overlapping patches.  images fp32
 * [N, C, H, W] -> out bf16 [N * (H/patch) * (W/patch), C * patch * patch], columns ordered (channel, row, column) like the flattened
 * conv weight [E, C, patch, patch]. */
int hamt_patchify_bf16(const float* images, void* out, int N, int C, int H, int W, int patch, void* stream);
/* VisionTransformer.forward_features head (vision_transformer.py:337-342): x = pos_drop(cat(cls_token, patch tokens) + pos_embed).
 * t0 bf16 [N * (S - 1), Hd] (patch projection incl. bias), cls fp32 [Hd], pos fp32 [S, Hd] -> x32 fp32 / x16 bf16 [N * S, Hd]. */
int hamt_vit_embed_fwd(const void* t0, const float* cls, const float* pos, float* x32, void* x16, int N, int S, int Hd,
                       const unsigned long long* seed_ptr, unsigned int site, float p, void* stream);
/* backward of the above: dfull = dx o dropout mask (bf16 [N * S, Hd]; its column sums over N are the pos_embed / cls_token
 * gradients), dt0 = the patch-token rows of dfull (bf16 [N * (S - 1), Hd], operand of the patch-projection wgrad). */
int hamt_vit_embed_bwd(const void* dx, void* dfull, void* dt0, int N, int S, int Hd, const unsigned long long* seed_ptr, unsigned int site,
                       float p, void* stream);

/* backward: dx (grad of x, dropout applied; null to skip), dres = dz + dres_in (null to skip); dgamma/dbeta/dbias
 * (column sums, fp32) are ACCUMULATED into; any may be null. */
int hamt_ln_bwd(const void* dy, const void* z, const float* mean, const float* rstd, const float* gamma, const void* dres_in, void* dx, void* dres,
                float* dgamma, float* dbeta, float* dbias, int M, int H, const unsigned long long* seed_ptr, unsigned int site, float p,
                void* stream);

/* Attention implementation switch: 0 (default) = the TMA + tcgen05 packed-tile kernels (hamt_attn_tc.cu) wherever the shape fits their
 * envelope (key length <= 128 after padding to 8) and they win (more than 32 query rows per problem), legacy mma.sync kernels
 * (hamt_attn.cu) otherwise; 1 = legacy kernels only; 2 = tcgen05 kernels wherever the shape fits (A/B measurements, parity tests).  Both produce the same dropout masks (same counter hash over (sequence, head, query, key)). */
int hamt_attn_set_impl(int v);

/* fused attention, head_dim 64: out = dropout(softmax(q k^T * scale + mask)) v ; vilmodel.py:96-129 (self), :322-349 (cross).
 * element (b,s,h,d) of q at q + b*q_bstride + s*ldq + h*64 + d (same for k/v with kv strides, out with o strides).
 * mask: additive fp32 [B,Sk] (the reference's (1-m)*-10000 row) or null.  lse: fp32 [B,heads,Sq] (needed for backward). */
int hamt_attn_fwd(const void* q, const void* k, const void* v, long long q_bstride, long long kv_bstride, long long ldq, long long ldkv,
                  const float* mask, void* out, long long ldo, long long o_bstride, float* lse, int B, int heads, int Sq, int Sk, float scale,
                  const unsigned long long* seed_ptr, unsigned int site, float p, void* stream);
/* backward; dq/dk/dv use the q/k/v strides.  Sq, Sk <= 128: dS through shared memory; longer (RxR, L = 300; Sq + Sk <= ~790): the dQ pass
 * recomputes S and dP.  dbias_q/k/v (fp32 [heads*64], all three or none): += column sums of dq/dk/dv = the bias gradients of the
 * query/key/value Linears (vilmodel.py:80-82), replacing a separate pass over the [tokens, 3H] gradient. */
int hamt_attn_bwd(const void* q, const void* k, const void* v, long long q_bstride, long long kv_bstride, long long ldq, long long ldkv,
                  const float* mask, const void* out, long long ldo, long long o_bstride, const float* lse, const void* dout, long long lddo,
                  long long do_bstride, void* dq, void* dk, void* dv, int B, int heads, int Sq, int Sk, float scale,
                  const unsigned long long* seed_ptr, unsigned int site, float p, float* dbias_q, float* dbias_k, float* dbias_v, void* stream);

/* BertEmbeddings (vilmodel.py:54-69): out = dropout(LN(word[ids] + pos[s] + type0)), ids int64 [B,L]. */
int hamt_embed_text_fwd(const long long* ids, const float* word, const float* pos, const float* type0, const float* gamma, const float* beta,
                        void* out, int B, int L, int H, float eps, const unsigned long long* seed_ptr, unsigned int site, float p, void* stream);
int hamt_embed_text_bwd(const void* dy, const long long* ids, const float* word, const float* pos, const float* type0, const float* gamma,
                        float* dword, float* dpos, float* dtype0, float* dgamma, float* dbeta, int B, int L, int H, float eps,
                        const unsigned long long* seed_ptr, unsigned int site, float p, void* stream);

/* ImageEmbeddings / HistoryEmbeddings / pano-token embedding (vilmodel.py:496-505, :549-571):
 *   s = LN_img(t) + LN_ang(ang W_ang^T + b_ang) [+ add_vec] [+ nav_table[nav_ids]] [+ extra] [+ pos_table[pos]]
 *   out = dropout(g_f ? LN_f(s) : s)
 * t = img_linear(x) (bf16 [M,H], from hamt_gemm_bf16).  pos of row r = pos_ids ? pos_ids[r] : r % pos_mod. */
typedef struct {
  const void* t; const float* ang; int A;
  const float* w_ang; const float* b_ang; const float* g_img; const float* b_img; const float* g_ang; const float* be_ang;
  const float* add_vec; const float* nav_table; const long long* nav_ids; const float* extra;
  const float* pos_table; const long long* pos_ids; int pos_mod;
  const float* g_f; const float* b_f;
  void* out; int M; int H; float eps;
  const unsigned long long* seed_ptr; unsigned int site; float p;
} hamt_embed_feat_desc;
typedef struct {
  const void* dy; void* dt;
  float* dw_ang; float* db_ang; float* dg_img; float* db_img; float* dg_ang; float* dbe_ang; float* dadd_vec; float* dnav_table;
  float* dextra; float* dpos_table; float* dg_f; float* db_f; float* db_lin;
} hamt_embed_feat_grads;
int hamt_embed_feat_fwd(const hamt_embed_feat_desc* d, void* stream);
int hamt_embed_feat_bwd(const hamt_embed_feat_desc* d, const hamt_embed_feat_grads* g, void* stream);


/* streaming helpers */
int hamt_cast_f32_to_bf16(const float* in, void* out, long long n, void* stream);
int hamt_colsum_bf16(const void* x, long long ld, float* out, int M, int N, void* stream);      /* out[n] += sum_m x[m,n] (bias grads) */
int hamt_mean_pool_fwd(const void* x, float* out, int N, int P, int H, void* stream);            /* torch.mean(dim=2), vilmodel.py:563-564 */
int hamt_mean_pool_bwd(const float* dy, void* dx, int N, int P, int H, void* stream);
int hamt_add_bf16(const void* a, const void* b, void* out, long long n, void* stream);
int hamt_mul_rows_bf16(const void* a, const void* v, void* out, int B, int S, int H, void* stream); /* ob * txt[:, :1], pretrain_cmt.py:176 */

/* proxy-task head pieces (pretrain_src/model/pretrain_cmt.py)
 * rowdot: final Linear(H -> N<=4) of the SAP/SAR/SPREL/ITM heads (:13-47,:62-71): y fp32 [M,N] = x(bf16 [M,H]) w^T(fp32 [N,H]) + b;
 * backward: dx bf16 (null to skip), dw/db accumulated.  ce: F.cross_entropy(reduction='none') over fp32 logits with -inf entries
 * allowed (:150-153,:177-180,:254-259); ce_bwd writes gloss[m]*(softmax - onehot) as fp32 [M,N] (pitch ld_d) or as bf16 padded to
 * pitch ld_d (zeros beyond N) ready to be a TMA GEMM operand.  gather/scatter: hidden[mask] rows (:161-165). */
int hamt_rowdot_fwd(const void* x, const float* w, const float* b, float* y, int M, int N, int H, void* stream);
int hamt_rowdot_bwd(const float* dy, const void* x, const float* w, void* dx, float* dw, float* db, int M, int N, int H, void* stream);
int hamt_ce_fwd(const float* logits, long long ld, const long long* labels, float* loss, float* lse, int M, int N, void* stream);
int hamt_ce_bwd(const float* logits, long long ld, const long long* labels, const float* lse, const float* gloss, float* dl_f32, void* dl_bf16,
                long long ld_d, int M, int N, void* stream);
int hamt_gather_rows_bf16(const void* x, const long long* idx, void* out, int n, int H, void* stream);   /* out[i] = x[idx[i]] */
int hamt_scatter_rows_bf16(const void* x, const long long* idx, void* out, int n, int H, void* stream);  /* out[idx[i]] = x[i] */
/* device-side batch assembly from a resident bf16 feature table [x_rows, H] (replaces np.stack + pad_tensors of
 * pretrain_src/data/r2r_data.py:264-329, data/common.py:5-20 and the 106 MB host->device copy per batch): out[i] = idx[i] >= 0 ?
 * x[idx[i]] : 0.  Indices are NOT range-checked on the device; the host wrapper checks them. */
int hamt_gather_rows_pad_bf16(const void* x, long long x_rows, const long long* idx, void* out, long long n, int H, void* stream);

/* fused optimizer step over the flat parameter arena: HF-style AdamW exactly as pretrain_src/optim/adamw.py:53-110
 * (bias correction, eps outside the sqrt, decoupled decay AFTER the update with lr * wd; parameters without a gradient this step are
 * skipped and keep their step counter, :64-66) + torch.nn.utils.clip_grad_norm_ (main_r2r.py:271-274) + bf16 shadow refresh + gradient
 * zeroing (optimizer.zero_grad(), main_r2r.py:281).  Segment s = one parameter; chunk_seg maps every 64-element chunk of the flat
 * buffers to its segment (-1 = alignment padding); seg_end[s] = one past the parameter's last element.  workspace[0] = global gradient norm (before clipping), [1] = clip coefficient. */
int hamt_adamw_workspace_floats(void);
int hamt_adamw_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16, long long total, const int* chunk_seg, const long long* seg_end, int nseg,
                    const unsigned char* seg_active, const float* seg_wd, int* seg_step, float* seg_step_size, const float* lr, double beta1,
                    double beta2, double eps, int correct_bias, float max_grad_norm, int want_norm, int zero_grad, float* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HAMT_B200_H */
