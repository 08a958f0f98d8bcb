/* mmfb200.h - C ABI of the B200-native multimodal-fusion block (libmmfb200.so).
 *
 * The reference (facebookresearch/mmf) has no FFI for this path: the boundary is the Python-level
 * module contract of
 *   BertSelfAttention/BertAttention/BertLayer/BertEncoder.forward   mmf/modules/hf_layers.py:161-355
 *   BertBiAttention/BertBiOutput/BertConnectionLayer.forward         mmf/models/vilbert.py:388-556
 *   BertVisioLinguisticEmbeddings.forward                            mmf/modules/embeddings.py:423-459
 *   BertImageFeatureEmbeddings.forward                               mmf/models/vilbert.py:904-913
 *   ModalEmbeddings.forward                                          mmf/models/mmbt.py:84-129
 *   HuggingfaceEmbeddings.forward                                    mmf/models/transformers/backends/huggingface.py:131-159
 * and their autograd backward (mmf/trainers/core/training_loop.py:211-213).  This header is what a
 * binding for that path calls (mmf_b200/lib.py is the ctypes binding; INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless stated; bf16 = 16-bit bfloat, row-major, explicit
 *     leading dimensions in ELEMENTS; the caller owns every buffer (outputs, saved-for-backward,
 *     workspace); functions never allocate, never synchronise, enqueue on `stream` only
 *   - return value: 0 = ok, otherwise an mmfb_status; mmfb_last_error() gives the message of the
 *     last failure on the calling thread; no C++ exception crosses the ABI
 *   - there is no CPU fallback: without an sm_100 device every compute entry point fails
 */
#ifndef MMFB200_H_
#define MMFB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* mmfb_stream; /* cudaStream_t */

typedef enum {
  MMFB_OK = 0,
  MMFB_ERR_ARG = 1,     /* bad shape / alignment / null pointer  -> ValueError in the binding */
  MMFB_ERR_CUDA = 2,    /* CUDA runtime / driver failure         -> RuntimeError */
  MMFB_ERR_DEVICE = 3   /* no sm_100 device                      -> RuntimeError */
} mmfb_status;

/* epilogues of mmfb_gemm */
enum {
  MMFB_EPI_BIAS = 0,            /* C = acc + bias                      (bias may be NULL)            */
  MMFB_EPI_BIAS_GELU = 1,       /* C = acc + bias ; C2 = gelu_erf(C)   BertIntermediate              */
  MMFB_EPI_BIAS_DROP_RESID = 2, /* C = dropout(acc + bias) + aux       BertSelfOutput/BertOutput pre-LN */
  MMFB_EPI_GELU_BWD = 3,        /* C = acc * gelu_erf'(aux)            dgrad through BertIntermediate */
  MMFB_EPI_ADD_AUX = 4,         /* C = acc + aux                       dgrad + residual gradient     */
  MMFB_EPI_ATOMIC_F32 = 5,      /* C(fp32) += acc                      split-K weight gradient       */
  MMFB_EPI_BIAS_RELU = 6        /* C = relu(acc + bias)                FinetuneFasterRcnnFpnFc7, encoders.py:176-179 */
};

/* C[M,N] = epi( A[M,K] * B[N,K]^T ).  a_mn/b_mn = 0: operand stored [rows,K] (K contiguous);
 * = 1: operand stored [K,rows] (rows contiguous), i.e. the transposed view of a row-major tensor. */
typedef struct {
  const void* A; int64_t lda; int a_mn;
  const void* B; int64_t ldb; int b_mn;
  void* C; int64_t ldc;          /* bf16 [M,N]; fp32 [M,N] for MMFB_EPI_ATOMIC_F32 */
  void* C2;                      /* second bf16 output (MMFB_EPI_BIAS_GELU), same ldc */
  const void* bias;              /* bf16 [N] or NULL */
  const void* aux; int64_t ldaux;/* bf16 [M,N] residual / pre-activation or NULL */
  const uint32_t* drop_mask; int64_t ldmask; float drop_scale; /* keep-bits [M, ldmask] words or NULL */
  int M, N, K;
  int epi;
  int splits;                    /* split-K factor, MMFB_EPI_ATOMIC_F32 only (0/1 = none) */
  int block_n;                   /* 0 = auto, else 128 or 256 */
  int cluster;                   /* 0 = library default, 1 = single CTA, 2 = 2-CTA cluster sharing the B tile by TMA multicast */
} mmfb_gemm_args;

int mmfb_gemm(const mmfb_gemm_args* args, mmfb_stream stream);

/* Fused multi-head attention core (everything between the Q/K/V projections and the output projection):
 *   ctx = dropout(softmax(Q K^T / sqrt(d) + mask)) V        BertSelfAttentionJit.forward hf_layers.py:182-210,
 *   BertBiAttention.forward vilbert.py:421-461 (cross: q from one stream, k/v from the other).
 * q [B*Sq, ldq], k/v [B*Skv, ld*]: bf16, head h in columns [h*head_dim, (h+1)*head_dim); usually three
 * column slices of one fused projection buffer.  mask: additive fp32 [B, Skv] (the reference's
 * [B,1,1,Skv] tensor) or NULL.  lse2 [B, heads, Sq] fp32 receives the log2-domain row log-sum-exp
 * (saved for backward).  drop_mask: keep bits [B, heads, Sq, ceil(Skv/32)] (bit kv%32) or NULL.
 * Limits: head_dim 64 or 128; Skv <= 384 (d=64) / 256 (d=128) - MMF's path has S <= 324. */
typedef struct {
  const void* q; int64_t ldq;
  const void* k; int64_t ldk;
  const void* v; int64_t ldv;
  const float* mask;
  void* ctx; int64_t ldo;            /* bf16 [B*Sq, ldo]: output of fwd, input of bwd */
  float* lse2;
  const uint32_t* drop_mask; float drop_scale;
  void* ctx_lo;                      /* optional bf16 [B*Sq, heads*head_dim], contiguous: bf16(O - float(ctx)), the part of the
                                        fp32 output the bf16 ctx drops; written by fwd, read by bwd (delta = rowsum(dO*(ctx+ctx_lo))) */
  /* backward only */
  const void* dctx; int64_t ld_dctx; /* bf16 [B*Sq, ld_dctx] */
  float* delta;                      /* fp32 scratch [B, heads, Sq] */
  void* dq; int64_t ld_dq;           /* bf16 outputs, same head layout as q/k/v */
  void* dk; int64_t ld_dk;
  void* dv; int64_t ld_dv;
  int B, heads, Sq, Skv, head_dim;
} mmfb_attn_args;

int mmfb_attention_fwd(const mmfb_attn_args* args, mmfb_stream stream);
int mmfb_attention_bwd(const mmfb_attn_args* args, mmfb_stream stream);

/* LayerNorm over the last dimension (eps 1e-12 on this path), one warp per row.
 * fwd: x = LN(y)*gamma + beta, optional dropout on x (embeddings), saves mean / rstd (fp32 [M]).
 * bwd: dy = dLN(dx (+dx2)); dz = dropout-backward(dy) for the dense branch that fed y = dropout(dense)+resid;
 *      dgamma/dbeta/dbias (fp32 [H]) are ACCUMULATED (+=).  dz may equal dy (no dropout) or be NULL. */
typedef struct {
  const void* y; int64_t ldy;        /* bf16 [M,H] pre-LN input */
  const void* gamma; const void* beta; /* bf16 [H] */
  void* x; int64_t ldx;              /* fwd output bf16 [M,H] */
  float* mean; float* rstd;          /* fp32 [M]: written by fwd, read by bwd */
  const uint32_t* drop_mask; int64_t ldmask; float drop_scale;
  float eps;
  /* backward */
  const void* dx; int64_t lddx;      /* bf16 [M,H] gradient wrt LN output */
  const void* dx2; int64_t lddx2;    /* optional second gradient term added to dx, or NULL */
  void* dy; int64_t lddy;            /* bf16 [M,H] gradient wrt y (residual branch) */
  void* dz; int64_t lddz;            /* bf16 [M,H] gradient wrt the dropped dense output */
  float* dgamma; float* dbeta; float* dbias;
  int M, H;
} mmfb_ln_args;

int mmfb_layernorm_fwd(const mmfb_ln_args* args, mmfb_stream stream);
int mmfb_layernorm_bwd(const mmfb_ln_args* args, mmfb_stream stream);

/* out[n] += sum_m X[m,n]   (bias gradients), X bf16 [M,N], out fp32 [N] */
int mmfb_colsum(const void* X, int64_t ldx, float* out, int M, int N, mmfb_stream stream);

/* Philox keep-bits for nn.Dropout(p): word w bit j <-> element 32*w+j, P(bit=1) = 1-p (16-bit resolution) */
int mmfb_dropout_bits(uint32_t* out, int64_t nwords, uint64_t seed, uint64_t offset, float p, mmfb_stream stream);
/* the same with a step counter in DEVICE memory mixed into the Philox counter (counter += *epoch << 40): a CUDA-graph replay of a
 * captured training step draws fresh masks as long as the graph increments *epoch (mmf_b200/graphs.py) */
int mmfb_dropout_bits_epoch(uint32_t* out, int64_t nwords, uint64_t seed, uint64_t offset, const uint64_t* epoch, float p,
                            mmfb_stream stream);

/* Embedding row composer (K1): y[r] = src0[src_row0[r]] + src1[src_row1[r]] + tab0[idx0[r]] + tab1[idx1[r]] +
 * tab2[idx2[r]]; a NULL pointer or a negative index drops the term.  src*: bf16 [*, ldsrc] dense rows (e.g. the
 * region-feature projection), tab*: bf16 [*, H] embedding tables.  Index arrays are int32 [M] on the device.
 * Covers BertVisioLinguisticEmbeddings (embeddings.py:329-370), ModalEmbeddings (mmbt.py:92-129),
 * HuggingfaceEmbeddings (huggingface.py:131-159) and BertImageFeatureEmbeddings (vilbert.py:904-913). */
typedef struct {
  const void* src[2]; int64_t ldsrc[2]; const int32_t* src_row[2];
  const void* tab[3]; const int32_t* tab_idx[3];
  void* y; int64_t ldy;
  int M, H;
} mmfb_compose_args;
int mmfb_embed_compose(const mmfb_compose_args* args, mmfb_stream stream);

/* backward of the composer: dsrc_k[src_row_k[r]] = dy[r] (bf16 store), dtab_k[idx_k[r]] += dy[r] (fp32 atomics) */
typedef struct {
  void* dsrc[2]; int64_t ldsrc[2]; const int32_t* src_row[2];
  float* dtab[3]; const int32_t* tab_idx[3];
  const void* dy; int64_t lddy;
  int M, H;
} mmfb_scatter_args;
int mmfb_embed_scatter(const mmfb_scatter_args* args, mmfb_stream stream);

/* table gradient with run aggregation: `sorted_idx` = the table indices in ascending order (negative = skip),
 * `order[k]` = the row of dy that sorted position k came from.  dtab[sorted_idx[k]] += dy[order[k]].
 * Equal indices are summed in registers first (one atomic per run of <= 32 rows instead of one per row). */
int mmfb_embed_scatter_sorted(const void* dy, int64_t lddy, const int32_t* order, const int32_t* sorted_idx, float* dtab,
                              int M, int H, mmfb_stream stream);

/* dz = dy where y > 0 else 0 (backward of the ReLU epilogue); contiguous bf16 buffers of n elements */
int mmfb_relu_bwd(const void* dy, const void* y, void* dz, int64_t n, mmfb_stream stream);

/* du = dh * GELU'(u) (erf GELU, the arithmetic of the MMFB_EPI_GELU_BWD epilogue) for a GELU that is NOT followed by a GEMM:
 * BertPredictionHeadTransform (dense -> gelu -> LayerNorm), HF modeling_bert via mmf/models/visual_bert.py:205-214.
 * Contiguous bf16 buffers of n elements. */
int mmfb_gelu_bwd(const void* dh, const void* u, void* du, int64_t n, mmfb_stream stream);

/* Cross-entropy over vocabulary-sized rows, forward and backward in ONE kernel: for every row m of the bf16 logits [M, V]
 * (row stride ldl, V % 8 == 0) with label[m] != ignore_index,  loss_m = logsumexp(z_m) - z_m[label[m]]  is added to
 * *loss_sum (and stored in row_loss[m] when given) and the row is OVERWRITTEN by
 * d(logits) = (softmax(z_m) - onehot(label[m])) * grad_scale;  ignored rows are overwritten by zeros.
 * Replaces CrossEntropyLoss(ignore_index=-1) over prediction_scores.view(-1, vocab_size) of the masked-LM heads
 * (mmf/models/visual_bert.py:215,269-277; mmf/models/transformers/heads/mlm.py:46-97) inside the chunked
 * linear + cross-entropy head: the full [tokens, 30522] logits tensor never exists in HBM. */
int mmfb_ce_rows(void* logits, int64_t ldl, const int64_t* labels, int64_t ignore_index, int M, int V, float grad_scale,
                 float* loss_sum, float* row_loss, mmfb_stream stream);

/* out = a + b; contiguous bf16 buffers of n elements.  The residual-gradient join of a pre-LN layer (ViT: the residual
 * branches off BEFORE the LayerNorm, mmf/modules/vit.py:96-108), which a post-LN layer gets for free from the GEMM epilogue. */
int mmfb_add_bf16(const void* a, const void* b, void* out, int64_t n, mmfb_stream stream);

/* out[m, h] = keep-bit(m, h) ? x[m, h] * scale : 0  (bf16 [M, H] with row strides, bits = int32 words [M, ldm] as produced by
 * mmfb_dropout_bits).  Backward of a dropout that no LayerNorm follows: HF ViTSelfOutput / ViTOutput via vit.py:66-68. */
int mmfb_dropout_apply(const void* x, int64_t ldx, const uint32_t* bits, int64_t ldm, float scale, void* out, int64_t ldo,
                       int M, int H, mmfb_stream stream);

/* fp32 -> bf16 cast of a flat (parameter) buffer */
int mmfb_cast_f32_bf16(const float* in, void* out, int64_t n, mmfb_stream stream);

/* Fused AdamW step over a FLAT fp32 parameter / gradient buffer (the engine's ParamPack), optionally refreshing the
 * bf16 compute copy in the same pass.  SURVEY.md 8(f) item 2 - staged: parity tests exist, not yet run on the GPU.
 * Replaces the per-tensor loop of the reference's optimizer "adam_w" (mmf/modules/optimizers.py:8-17):
 *   mode 0: transformers AdamW arithmetic as restated in the reference (optimizers.py:60-84):
 *           m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= step_size * m / (sqrt(v) + eps);  p -= lr * wd * p
 *           with step_size = lr * sqrt(1-b2^t)/(1-b1^t) (or lr when correct_bias is off), computed by the host
 *   mode 1: torch.optim.AdamW arithmetic (the reference's fallback when transformers has no AdamW):
 *           p *= 1 - lr wd;  m += (g-m)(1-b1);  v = b2 v + (1-b2) g^2;  p -= step_size * m / (sqrt(v)/bc2_sqrt + eps)
 *           with step_size = lr/(1-b1^t), bc2_sqrt = sqrt(1-b2^t)
 * `group` (nullable) gives the hyper-parameter group of every 8-element block (parameters are padded to multiples of
 * 8 in the flat buffer); grad_scale multiplies the gradient first (1/loss_scale, clipping coefficient). */
#define MMFB_ADAMW_MAX_GROUPS 8
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  void* param_bf16;          /* nullable */
  const uint8_t* group;      /* nullable: all blocks use group 0 */
  int64_t n;                 /* elements, multiple of 8 */
  int n_groups;
  float lr[MMFB_ADAMW_MAX_GROUPS];
  float weight_decay[MMFB_ADAMW_MAX_GROUPS];
  float step_size[MMFB_ADAMW_MAX_GROUPS];
  float bc2_sqrt[MMFB_ADAMW_MAX_GROUPS];
  float beta1, beta2, eps, grad_scale;
  int mode;
} mmfb_adamw_args;
int mmfb_adamw(const mmfb_adamw_args* args, mmfb_stream stream);

/* library / diagnostics */
const char* mmfb_last_error(void);
int mmfb_version(void);
int mmfb_device_ok(void);          /* 1 if the current device is sm_100 */
int64_t mmfb_launch_count(void);   /* number of kernels this library has launched in this process */

#ifdef __cplusplus
}
#endif
#endif /* MMFB200_H_ */
