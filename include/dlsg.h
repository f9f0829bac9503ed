/* libdlsg - C-ABI of the B200 (sm_100a) kernels behind the D-LSG model API.
 *
 * The reference (baiyang4/D-LSG-Video-Caption) has no FFI: its boundary is the Python module API
 * models/{model,layer,sublayer,allennlp_beamsearch}.py.  Our drop-in mirror of those modules
 * (d-lsg-video-caption_b200/models/) binds these symbols with ctypes (dlsg/_lib.py); each entry
 * cites the reference call site whose arithmetic it replaces (file:line in /root/reference).
 *
 * Conventions: every pointer is a caller-owned DEVICE pointer unless stated; sizes in elements;
 * `stream` is a cudaStream_t; return 0 on success, nonzero on error (dlsg_last_error() explains).
 * No allocation, no host sync, no global state besides the lazily resolved driver entry point for
 * cuTensorMapEncodeTiled: every entry is CUDA-graph capturable and re-entrant per stream.
 */
#ifndef DLSG_H_
#define DLSG_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { DLSG_F32 = 0, DLSG_BF16 = 1 };

/* ---- library ------------------------------------------------------------------------------ */
int dlsg_version(void);
int dlsg_sm_arch(void);                /* 100 = built for sm_100a                               */
const char* dlsg_last_error(void);
/* Debug only: when dev_buf != NULL every following DLSG_GEMM_TC launch writes per-CTA phase timestamps
 * (148 CTAs x 8 phases x {globaltimer ns, clock64}) into it; NULL switches tracing off (default).     */
void dlsg_debug_gemm_trace(void* dev_buf);

/* ---- dense projections (nn.Linear / LSTM gate GEMMs / vocab projection / bmm) -------------
 * D[b](M,N) = epi( A[b](M,K) . B[b](N,K)^T + bias ).  Replaces torch addmm/bmm at
 * layer.py:51,179,184 (embeddings), layer.py:52,571,593 + model.py:152 (LSTM gates),
 * sublayer.py:29-31,66-68,80 (attention projections), layer.py:600 (word_restore),
 * model.py:148 (conv1d k=1), layer.py:187,191 / sublayer.py:69,76,191,196 (bmm).
 *  impl DLSG_GEMM_TC  : TMA -> smem -> tcgen05.mma (TMEM accum), bf16 operands, fp32 accumulate.
 *                        Each operand is K-major (sak == 1 / sbk == 1) or MN-major (a transposed view: sam == 1 /
 *                        sbn == 1, read in place through 64x64 TMA boxes and the UMMA MN-major descriptor - no
 *                        transposed copy is ever needed); the non-unit pitch is a multiple of 8 elements,
 *                        16-byte aligned bases.  splitk>1 writes partial sums to D + s*stride_split.
 *  impl DLSG_GEMM_SIMT: fp32 FFMA tiles, arbitrary element strides (sam,sak,sbn,sbk) and dtypes.
 */
enum { DLSG_GEMM_SIMT = 0, DLSG_GEMM_TC = 1 };
enum {
  DLSG_EPI_BIAS_N = 1,   /* bias[n] added                                                      */
  DLSG_EPI_BIAS_M = 2,   /* bias[m] added                                                      */
  DLSG_EPI_TANH = 4,     /* tanh after bias                                                    */
  DLSG_EPI_ACCUM = 8,    /* D += result                                                        */
  DLSG_EPI_STORE_T = 16, /* store D^T: element (m,n) goes to D[n*ldd + m]                      */
  DLSG_EPI_ATOMIC = 32,  /* D += result through fp32 atomic adds (fp32 D, no tanh): DLSG_GEMM_TC may then split K  */
                         /* over the idle SMs with no workspace and no reduce launch (D holds the addend, e.g.     */
                         /* zeros).  The order in which the splits land is not fixed: last-bit run-to-run noise.   */
                         /* DLSG_GEMM_SIMT treats it as DLSG_EPI_ACCUM.                                             */
  DLSG_GEMM_A_STATIC = 64,  /* promise: operand A (resp. B) is not written by the kernels that precede this launch in the  */
  DLSG_GEMM_B_STATIC = 128  /* stream (a weight inside a recurrent loop).  DLSG_GEMM_TC then requests its first ring-full  */
                            /* of tiles BEFORE the programmatic-dependent-launch wait, overlapping the weight stream with  */
                            /* the tail of the preceding kernels.  Results are identical; ignored by DLSG_GEMM_SIMT.       */
};
typedef struct {
  const void* A; const void* B; void* D; const float* bias;
  int32_t M, N, K, batch;
  int64_t sam, sak, sbn, sbk, ldd;          /* element strides; TC requires sak == sbk == 1     */
  int64_t stride_a, stride_b, stride_d;     /* batch strides (elements)                         */
  int32_t a_dtype, b_dtype, d_dtype, impl, flags, _pad0;
  int32_t splitk; int32_t _pad; int64_t stride_split;
  float alpha;                              /* result scale applied before bias                 */
  int32_t _pad1;
  void* workspace; int64_t workspace_bytes; /* optional scratch: lets DLSG_GEMM_TC split K automatically for skinny   */
                                            /* problems (partials -> workspace, then a reduce+epilogue kernel)        */
} dlsg_gemm_t;
int dlsg_gemm(const dlsg_gemm_t* p, void* stream);

/* ---- layout / dtype conversion ------------------------------------------------------------
 * dst[r,c] = src[r,c] (cast); optional dstT[c,r] = src[r,c].  Feeds bf16 GEMM operands.        */
int dlsg_convert2d(const void* src, int src_dtype, int64_t ld_src, void* dst, int dst_dtype, int64_t ld_dst,
                   void* dstT, int64_t ld_dstT, int64_t rows, int64_t cols, void* stream);
int dlsg_convert2d_batched(const void* src, int src_dtype, int64_t ld_src, void* dst, int dst_dtype, int64_t ld_dst,
                           void* dstT, int64_t ld_dstT, int64_t rows, int64_t cols, int64_t batch,
                           int64_t bs_src, int64_t bs_dst, int64_t bs_dstT, void* stream);
/* One launch for MANY conversions (all GEMM-operand copies of the parameters after an optimizer step):
 * a DEVICE-resident table of segments dst[r,c] = cast(src[r,c] (+ src2[r,c])), fp32 sources (src2 optional, same
 * pitch) or bf16 sources (no src2: a reduced gradient bucket read back), bf16 / fp32 destination with its own pitch, and a DEVICE table of int32 triples (segment, first row, rows),
 * one CTA per triple.                                                                              */
typedef struct {
  const void* src; const void* src2; void* dst;
  int64_t rows, cols, ld_src, ld_dst;
  int32_t src_dtype, dst_dtype;               /* a DLSG_BF16 source takes no src2 */
} dlsg_seg_t;
int dlsg_multi_convert(const dlsg_seg_t* segs_dev, const int32_t* chunks_dev, int32_t nchunks, void* stream);
/* Same conversions from a HOST table that travels inside the kernel parameters (up to 384 segments per launch, one CTA per
 * ~chunk_elems elements): no table upload, so it may be issued with fresh pointers during a CUDA-graph capture (the
 * per-block gradient packs of the data-parallel step).                                              */
int dlsg_multi_convert_host(const dlsg_seg_t* segs_host, int32_t nsegs, int32_t chunk_elems, void* stream);
/* Multi-tensor Adam (torch.optim.Adam semantics, no weight decay / amsgrad / maximize; run_gun.py:91,100) over a table of
 * 2-D segments (p, m, v fp32 with one common pitch; g fp32 or bf16 with its own pitch - the all-reduced gradient bucket of
 * a data-parallel step is read in place; dst16 = optional bf16 GEMM-operand copy of the updated values with its own pitch).  The table is a HOST array: it travels to the device inside the kernel's parameters (up to 256 segments
 * per launch, one CTA per ~chunk_elems elements), so a CUDA graph keeps it in the kernel node and no table upload exists.
 * `step_dev` holds the (already incremented) step count t as a float; the learning rate is *lr_dev when lr_dev != NULL (so
 * a captured graph follows a scheduler), else lr.  SURVEY 8f-3: the optimizer pass emits the bf16 weights.               */
typedef struct {
  float* p; const void* g; float* m; float* v; void* dst16;
  int64_t rows, cols, ld, ld_dst;
  int64_t ld_g;                               /* pitch of g (its own: a slice of a flat gradient bucket)            */
  int32_t g_dtype, _pad;                      /* DLSG_F32 | DLSG_BF16: data-parallel buckets are reduced in bf16    */
  const float* step;                          /* optional DEVICE scalar: this parameter's own (already incremented)  */
                                              /* step count; NULL = the launch-wide step_dev                         */
} dlsg_adam_seg_t;
int dlsg_adam_multi(const dlsg_adam_seg_t* segs_host, int32_t nsegs, int32_t chunk_elems, const float* step_dev,
                    const float* lr_dev, float lr, float beta1, float beta2, float eps, void* stream);
/* out[c] += sum_r x[r,c] : bias gradients (fp32 accumulate into out)                            */
int dlsg_colsum(const void* x, int dtype, int64_t ld, int64_t rows, int64_t cols, float* out, void* stream);

/* ---- Tanh->LayerNorm family ---------------------------------------------------------------
 * y = post( LN( pre(x + res) ) ) * dropmask ; pre/post in {identity,tanh}.  Replaces
 * layer.py:145-163,53,57,574,599, sublayer.py:21-26,183-187, model.py:128-131,153.
 * stats (rows,2) = mean,rstd saved for backward.  y2 = optional second (bf16) copy of y.      */
enum {
  DLSG_NORM_PRE_TANH = 1, DLSG_NORM_POST_TANH = 2,
  DLSG_NORM_IN_IS_TANH = 4   /* x already holds tanh(.) (fused in the GEMM epilogue); backward  */
                             /* still applies (1-x^2)                                           */
};
typedef struct {
  const void* x; const void* res; void* y; void* y2; const float* gamma; const float* beta; float* stats;
  int64_t rows; int32_t D; int32_t flags;
  int64_t ldx, ldres, ldy, ldy2;
  int32_t x_dtype, res_dtype, y_dtype, y2_dtype;
  float drop_p; uint32_t _pad; uint64_t seed, offset;     /* dropout: keep iff philox(seed,offset+idx) >= p */
} dlsg_norm_fwd_t;
int dlsg_norm_fwd(const dlsg_norm_fwd_t* p, void* stream);
typedef struct {
  const void* dy; const void* x; const void* res; const float* gamma; const float* beta; const float* stats;
  void* dx; float* dgamma; float* dbeta;                   /* dgamma/dbeta are ACCUMULATED (+=)    */
  int64_t rows; int32_t D; int32_t flags;
  int64_t lddy, ldx, ldres, lddx;
  int32_t dy_dtype, x_dtype, res_dtype, dx_dtype;
  float drop_p; int32_t dx_accum; uint64_t seed, offset;
  float* dxsum;   /* optional (D): column sums of dx, ACCUMULATED (+=) - the bias gradient of the Linear that produced x; */
                  /* only the streaming bf16 form computes it (dlsg_norm_bwd_streaming(p) == 1), else it must be NULL   */
} dlsg_norm_bwd_t;
int dlsg_norm_bwd(const dlsg_norm_bwd_t* p, void* stream);
/* Backward of the (plain) LayerNorm backward with respect to (dy, x, gamma), for a cotangent u of dx: the second-order
 * term of the WGAN-GP gradient penalty through the critic's LayerNorms (run_gun.py:362-375 over model.py:128-131,153,
 * layer.py:665-681).  All tensors contiguous fp32 (rows, D), D <= 1024; stats = (mean, rstd) of x saved by the forward;
 * g_gamma (D) is ACCUMULATED (+=); g_dy / g_x / g_gamma may be NULL.                                */
typedef struct {
  const float* x; const float* dy; const float* u; const float* gamma; const float* stats;
  float* g_dy; float* g_x; float* g_gamma;
  int64_t rows; int32_t D; int32_t _pad;
} dlsg_norm_bwd2_t;
int dlsg_norm_bwd2(const dlsg_norm_bwd2_t* p, void* stream);
int dlsg_norm_bwd_streaming(const dlsg_norm_bwd_t* p);   /* 1 if this call runs the streaming bf16 kernel (large bf16 x/dy/dx) */

/* ---- LSTM cell pointwise (nn.LSTMCell / nn.LSTM step, gate order i,f,g,o) ------------------
 * layer.py:52,571,593, model.py:152.  gates: nsplit partial (B,4H) fp32 buffers (split-K GEMM
 * output) + optional per-row bias (B,4H) + optional bias (4H).  Writes activated gates back to
 * gates[0] (saved for backward), c_out, h_out (fp32) and up to two extra copies of h (any dtype,
 * own leading dimension) straight into the next GEMM's operand buffers.                        */
typedef struct {
  float* gates; int32_t nsplit; int32_t _pad0; int64_t stride_split;
  const float* row_bias; int64_t ld_row_bias; const float* bias;
  const float* c_prev; float* c_out; float* h_out;
  void* h2; int64_t ldh2; int32_t h2_dtype; int32_t _pad1;
  void* h3; int64_t ldh3; int32_t h3_dtype; int32_t _pad2;
  int32_t B, H;
  float drop_p; int32_t _pad3; uint64_t seed, offset;     /* dropout on h (layer.py:594)           */
} dlsg_lstm_cell_fwd_t;
int dlsg_lstm_cell_fwd(const dlsg_lstm_cell_fwd_t* p, void* stream);
typedef struct {
  const float* acts; const float* c_prev; const float* c_new;
  const float* dh; int64_t lddh; const float* dh2; int64_t lddh2;  /* dh_total = dh + dh2 (dh2 may be NULL) */
  const float* dc_next;                                    /* dc_next may be NULL (=0)              */
  float* dgates; void* dgates2; int64_t ld_dgates2; int32_t dgates2_dtype; int32_t _pad0;
  void* dgatesT; int64_t ld_dgatesT; int32_t dgatesT_dtype; int32_t _pad1;  /* (4H, .) transposed copy */
  float* dc_prev;
  int32_t B, H;
  float drop_p; int32_t _pad2; uint64_t seed, offset;
  int32_t dh2_nsplit; int32_t _pad3; int64_t dh2_stride_split;   /* dh2 = sum of dh2_nsplit (<=1: one) split-K partial buffers */
  /* injections of the critic's second-order reverse pass (dlsg.generic._LstmBptt2): contiguous fp32, may be NULL */
  const float* dc_next2;      /* (B,H)  added to dc_next                                              */
  const float* dgates_add;    /* (B,4H) added to the gate gradients before they are written anywhere  */
  float* dh_total;            /* (B,H)  out: dh + dh2 (before dropout), saved for dlsg_lstm_cell_bwd2 */
  int64_t ld_dgates;          /* row pitch of dgates (0: contiguous, 4H) - a (B,4H) slice of a batch-major (B,T,4H) tensor */
} dlsg_lstm_cell_bwd_t;
int dlsg_lstm_cell_bwd(const dlsg_lstm_cell_bwd_t* p, void* stream);
/* Backward of the cell backward (second-order term of the WGAN-GP gradient penalty through DiscV2.lstm,
 * run_gun.py:362-375 over model.py:152).  The cell backward maps (dh, dc_next; pre-activations, c_prev) ->
 * (dgates (B,4H), dc_prev); given cotangents u (B,4H) of dgates and w (B,H) of dc_prev (either may be NULL = 0)
 * this returns the cotangents of dh, dc_next, the gate pre-activations (B,4H) and c_prev (outputs may be NULL).
 * All tensors contiguous fp32.                                                                    */
typedef struct {
  const float* acts; const float* c_prev; const float* c_new; const float* dh; const float* dc_next;
  const float* u; const float* w;
  float* g_dh; float* g_dc; float* g_pre; float* g_cprev;
  int32_t B, H;
  /* u_total = u + sum of u2_nsplit (<=1: one) split-K partial buffers (B,4H) of the recurrent product g_dh(t-1) W^T  */
  const float* u2; int64_t u2_stride_split; int32_t u2_nsplit; int32_t g_dh2_dtype;
  void* g_dh2; int64_t ld_g_dh2;     /* optional second copy of g_dh (any dtype, own pitch): the next step's GEMM operand */
  int64_t ld_u, ld_g_dh;             /* row pitches of u / g_dh (0: contiguous): slices of batch-major (B,T,.) tensors    */
} dlsg_lstm_cell_bwd2_t;
int dlsg_lstm_cell_bwd2(const dlsg_lstm_cell_bwd2_t* p, void* stream);

/* ---- per-step fusions (one CTA per batch row, H <= 2048): cell + LayerNorm forward, LayerNorm + cell backward.
 * fwd: split-K partials/bias -> gates -> c,h (dropout on h) -> y = [tanh](LN(h)) (dropout on y); replaces
 *      reduce + lstm_cell_fwd + norm_fwd (layer.py:571-574, 593-599).
 * bwd: dh = LNbwd(dy; x=h) + cell.dh + cell.dh2, then the cell backward; dgates_sum (optional) += dgates. */
typedef struct {
  dlsg_lstm_cell_fwd_t cell;
  const float* gamma; const float* beta; float* stats;
  void* y; int64_t ldy; void* y2; int64_t ldy2;
  int32_t y_dtype, y2_dtype, post_tanh, _pad;
  float ydrop_p; int32_t _pad2; uint64_t yseed, yoffset;
} dlsg_lstm_cell_norm_fwd_t;
int dlsg_lstm_cell_norm_fwd(const dlsg_lstm_cell_norm_fwd_t* p, void* stream);
typedef struct {
  dlsg_lstm_cell_bwd_t cell;            /* cell.dh / cell.dh2: recurrent gradients wrt the dropped h (may be NULL) */
  const float* dy; int64_t lddy; const float* x; int64_t ldx;
  const float* gamma; const float* beta; const float* stats;
  float* dgamma; float* dbeta; int64_t ld_dparam;   /* per-ROW contributions (B,H), written (not accumulated): the caller */
                                                    /* sums them over rows / time steps with one dlsg_colsum after BPTT   */
  float* dgates_sum;
  int32_t post_tanh, _pad;
  float ydrop_p; int32_t _pad2; uint64_t yseed, yoffset;
} dlsg_norm_lstm_cell_bwd_t;
int dlsg_norm_lstm_cell_bwd(const dlsg_norm_lstm_cell_bwd_t* p, void* stream);
int dlsg_fused_step_supported(int32_t H);          /* H % 4 == 0 && H <= 2048 */

/* ---- softmax over an arbitrary axis (layer.py:188 dim=1, sublayer.py:34,74,192, layer.py:706)
 * x viewed as (outer, n, inner) with element strides; optional scale, optional mask (>0 keeps,
 * else fill -9e15 BEFORE softmax: sublayer.py:70-72) or post-mask (zero AFTER softmax: layer.py:707). */
typedef struct {
  const float* x; float* y; const float* mask;
  int64_t outer, n, inner; int64_t so, sn, si;   /* strides of x / y / mask (same layout)        */
  float scale; int32_t mask_mode;                /* 0 none, 1 pre (-9e15), 2 post (zero)          */
} dlsg_softmax_t;
int dlsg_softmax_fwd(const dlsg_softmax_t* p, void* stream);
/* dx = scale * y_unmasked*(dy' - sum(dy'*y_unmasked)), dy' = dy*postmask; y = forward output   */
int dlsg_softmax_bwd(const dlsg_softmax_t* p, const float* dy, float* dx, void* stream);
/* Backward OF dlsg_softmax_bwd with respect to (dy, x) for a cotangent u of dx (same layout as x): the second-order term of
 * the WGAN-GP penalty through the critic's three softmaxes (run_gun.py:362-375 over sublayer.py:34,74,192, layer.py:706):
 *   s = softmax(scale x [pre-masked]), g = dy [o postmask], u' = u [o premask], A = <g,s>, B = <u',s>
 *   g_dy = scale s (u' - B) [o postmask];  q = scale (u' (g - A) - g B);  g_x = scale s (q - <q,s>).
 * g_dy / g_x may be NULL. */
int dlsg_softmax_bwd2(const dlsg_softmax_t* p, const float* dy, const float* u, float* g_dy, float* g_x, void* stream);

/* ---- small fused element-wise forms (contiguous fp32, n elements): the pieces of the critic's first and second backward
 * that are neither a GEMM, a LayerNorm, a softmax nor an LSTM cell (model.py:145-168, sublayer.py:167-170, run_gun.py:355-358).
 * in[] / out[] are used as listed per op; `cols` is the row length for ops with a per-row operand e (index i / cols).   */
enum {
  DLSG_EW_TANH_BWD = 0,       /* in: dy, y            out0 = dy (1 - y^2)                                             */
  DLSG_EW_TANH_BWD2 = 1,      /* in: dy, y, u         out0 = u (1 - y^2) [cot. of dy];  out1 = -2 y dy u [cot. of y]  */
  DLSG_EW_MUL_BWD = 2,        /* in: dy, a, b         out0 = dy b [da];  out1 = dy a [db]                             */
  DLSG_EW_MUL_BWD2 = 3,       /* in: dy, a, b, u0, u1 out0 = u0 b + u1 a [cot. of dy]; out1 = u1 dy [of a]; out2 = u0 dy [of b] */
  DLSG_EW_LERP_ROWS = 4,      /* in: a, b, e(rows)    out0 = a e + b (1 - e)                                          */
  DLSG_EW_LERP_ROWS_BWD = 5   /* in: dy, e(rows)      out0 = dy e;  out1 = dy (1 - e)       (any output may be NULL)  */
};
typedef struct {
  const float* in[5]; float* out[3];
  int64_t n; int64_t cols;
  int32_t op; int32_t _pad;
} dlsg_ew_t;
int dlsg_ew(const dlsg_ew_t* p, void* stream);

/* ---- AttentionShare core, one decode step (sublayer.py:32-39) ------------------------------
 * logits[r,p] = Kp[node(r),p,:].qp[r,:]/sqrt(H); alpha = softmax_p; ctx[r,:] = sum_p alpha*Vp.
 * node(r) = r / rows_per_node (beam rows share their clip's nodes).  heads: qp (rows, nh*H),
 * Kp/Vp (nodes, nh, P, H) fp32, alpha (rows, nh*P), ctx (rows, nh*H).                          */
typedef struct {
  const float* Kp; const float* Vp; const float* qp; float* alpha; void* ctx;
  int32_t rows, nh, P, H, rows_per_node, ctx_dtype; int64_t ldctx, ldalpha;
  int32_t nodes; int32_t _pad;
} dlsg_node_attn_fwd_t;
int dlsg_node_attn_fwd(const dlsg_node_attn_fwd_t* p, void* stream);
typedef struct {
  const float* Kp; const float* Vp; const float* qp; const float* alpha; const float* dctx; const float* dalpha_ext;
  void* dqp; float* dKp; float* dVp;                      /* dKp,dVp ACCUMULATED over steps        */
  int32_t rows, nh, P, H; int64_t lddctx, ldalpha;
  int32_t dqp_dtype; int32_t _pad;
} dlsg_node_attn_bwd_t;
int dlsg_node_attn_bwd(const dlsg_node_attn_bwd_t* p, void* stream);

/* ---- hoisted AttentionShare: query / output projections folded into the node tensors once per sequence
 * (KW = K Wq (nh,nodes,P,Hk), VW = V Wo^T (nh,nodes,P,Hv)); a step is ONE kernel: logits_p = KW_p.q*scale, softmax over
 * nodes, co = sum_p alpha_p VW_p (pre-LayerNorm context, sublayer.py:29-41).  Requires P<=8, Hk,Hv<=1024 (multiples of 4).
 * bwd: dq (rows,Hk) += , dKW / dVW ACCUMULATED over the steps; heads are summed in a fixed order (deterministic).       */
typedef struct {
  const float* KW; const float* VW; const float* q; float* alpha; float* co;
  int64_t ldq, ldalpha, ldco;
  int32_t rows, nh, P, Hk, Hv, rows_per_node, nodes; float scale;
  /* optional fused tail = the context output layer (sublayer.py:41 Tanh -> LayerNorm -> Dropout), active when y != NULL:
   * y[r, hd*Hv + c] = dropout(LN_hd(tanh(co[r, hd*Hv + c]))), stats[hd*stats_head_stride + 2r] = {mean, rstd};
   * dropout keeps element iff philox(seed, offset + hd*offset_head_stride + r*Hv + c) >= drop_p.                    */
  void* y; int64_t ldy; int32_t y_dtype; int32_t _pad0;
  const float* gamma[2]; const float* beta[2]; float* stats; int64_t stats_head_stride;
  float drop_p; int32_t _pad1; uint64_t seed, offset, offset_head_stride;
} dlsg_attn2_fwd_t;
typedef struct {
  const float* KW; const float* VW; const float* q; const float* alpha; const float* dco; const float* dalpha_ext;
  float* dq; float* dKW; float* dVW;
  int64_t ldq, ldalpha, lddco, lddq;
  int32_t rows, nh, P, Hk, Hv; float scale;
  /* optional fused head = backward of the context output layer, active when dy != NULL (dco is then ignored):
   * dco = d/dco [ dropout(LN_hd(tanh(co))) ] . dy ; the LayerNorm parameter gradients are written as per-ROW
   * contributions dgamma_rows / dbeta_rows (rows, nh*Hv) [ld_dparam] for one dlsg_colsum after BPTT (no atomics). */
  const float* dy; int64_t lddy; const float* co; int64_t ldco;
  const float* gamma[2]; const float* stats; int64_t stats_head_stride;
  float* dgamma_rows; float* dbeta_rows; int64_t ld_dparam;
  float drop_p; int32_t _pad1; uint64_t seed, offset, offset_head_stride;
  /* optional: DEFER the node gradients.  When dl_save != NULL this step records d(logits) (rows, nh*P) [ld_dl_save] and
   * d(co) (rows, nh*Hv) [ld_dco_save] and does NOT touch dKW / dVW; dlsg_attn2_bwd_nodes accumulates them over all steps
   * in one launch after the time loop (no per-step read-modify-write of the (nh, rows, P, H) node-gradient tensors). */
  float* dl_save; int64_t ld_dl_save; float* dco_save; int64_t ld_dco_save;
} dlsg_attn2_bwd_t;
int dlsg_attn2_supported(int32_t nh, int32_t P, int32_t Hk, int32_t Hv);
/* The query LSTM's cell + LayerNorm (dlsg_lstm_cell_norm_fwd) and the hoisted attention step + context output layer
 * (dlsg_attn2_fwd) of one decode step in ONE launch (layer.py:571-591): `at.q` is ignored - the attention reads
 * q = dropout(LN(query_h)) from the cell part (which still writes y / y2 as before).  One CTA per batch row, 256 threads
 * per head.  Supported: cell.H == at.Hk <= 1024, at.Hv <= 1024, nh <= 2, P <= 8, nsplit <= 4, cell.B == at.rows.       */
typedef struct { dlsg_lstm_cell_norm_fwd_t cn; dlsg_attn2_fwd_t at; } dlsg_cell_norm_attn2_fwd_t;
int dlsg_cell_norm_attn2_supported(const dlsg_cell_norm_attn2_fwd_t* f);
int dlsg_cell_norm_attn2_fwd(const dlsg_cell_norm_attn2_fwd_t* f, void* stream);
int dlsg_attn2_fwd(const dlsg_attn2_fwd_t* p, void* stream);
int dlsg_attn2_bwd(const dlsg_attn2_bwd_t* p, void* stream);
/* dKW[hd][r][j][:] (+)= sum_t dl[t][r][hd*P+j] q[t][r][:] ; dVW[hd][r][j][:] (+)= sum_t alpha[t][r][hd*P+j] dco[t][r][hd*Hv+:]
 * over T recorded steps (step t of tensor X at X + t * X_step_stride elements).  accumulate != 0: += onto dKW / dVW.       */
int dlsg_attn2_bwd_nodes(const float* q_all, int64_t ldq, int64_t q_step_stride, const float* dl_all, int64_t lddl, int64_t dl_step_stride,
                         const float* alpha_all, int64_t ldalpha, int64_t alpha_step_stride, const float* dco_all, int64_t lddco,
                         int64_t dco_step_stride, float* dKW, float* dVW, int32_t T, int32_t rows, int32_t nh, int32_t P, int32_t Hk,
                         int32_t Hv, int32_t accumulate, void* stream);

/* ---- LatentPSL pooling (sublayer.py:191-196), one fused kernel per direction, one CTA per clip (P<=8, T<=32, H%4==0)
 * fwd: Gs (B,T,P) = softmax over T of X theta^T ; N (B,P,H) = Gs^T X.   bwd: dX (B,T,H) written, dtheta (P,H) ACCUMULATED. */
int dlsg_latent_psl_fwd(const float* X, const float* theta, float* Gs, float* N, int32_t B, int32_t T, int32_t P, int32_t H, void* stream);
int dlsg_latent_psl_bwd(const float* X, const float* theta, const float* Gs, const float* dN, float* dX, float* dtheta,
                        int32_t B, int32_t T, int32_t P, int32_t H, void* stream);
/* the same for E <= 2 independent poolings in ONE launch (host arrays of E pointers: the object and the motion encoder)   */
int dlsg_latent_psl_fwd_multi(const float* const* X, const float* const* theta, float* const* Gs, float* const* N, int32_t E,
                              int32_t B, int32_t T, int32_t P, int32_t H, void* stream);
int dlsg_latent_psl_bwd_multi(const float* const* X, const float* const* theta, const float* const* Gs, const float* const* dN,
                              float* const* dX, float* const* dtheta, int32_t E, int32_t B, int32_t T, int32_t P, int32_t H, void* stream);

/* ---- one LSTM time step in one launch (nn.LSTM of EncoderVisual, layer.py:52; up to DLSG_LSTM_STEP_MAXG independent groups) -----
 * gates = h_in W^T + gin ; i,f,o = sigmoid, g = tanh ; c_out = f c_in + i g ; h = o tanh(c_out)   (torch gate order i,f,g,o)
 * Each CTA owns 16 hidden units (64 rows of W), streams its weight slab and h_in through shared memory (cp.async, bf16),
 * mma.sync with fp32 accumulation, cell in the epilogue: replaces a split-K GEMM launch + a cell launch per direction.
 * Supported: B <= 64 rows per group, H a multiple of 128, bf16 W / h_in with 16-byte aligned rows.  A group is a direction of
 * the BiLSTM (own weights) or a 64-row slice of a larger batch sharing one weight (the critic's stacked LSTM, model.py:152). */
#define DLSG_LSTM_STEP_MAXG 4
typedef struct {
  const void* W[DLSG_LSTM_STEP_MAXG];                       /* bf16 (4H, H) row-major recurrent weights                                   */
  const void* h_in[DLSG_LSTM_STEP_MAXG]; int64_t ldh_in;    /* bf16 (B, H) previous hidden state; NULL (all directions) = first step       */
  const float* gin[DLSG_LSTM_STEP_MAXG]; int64_t ldgin;     /* fp32 (B, 4H) input projection + biases of this step                         */
  const float* c_in[DLSG_LSTM_STEP_MAXG]; float* c_out[DLSG_LSTM_STEP_MAXG];  /* fp32 (B, H) contiguous; c_in NULL = zeros                                   */
  float* acts[DLSG_LSTM_STEP_MAXG];                         /* out fp32 (B, 4H) contiguous: activated gates [i|f|g|o] (saved for BPTT)     */
  float* h_out[DLSG_LSTM_STEP_MAXG]; int64_t ldh_out;       /* out fp32 (B, H) view, optional                                              */
  void* h_op[DLSG_LSTM_STEP_MAXG]; int64_t ldh_op;          /* out bf16 (B, H) view, optional: the next step's h_in                        */
  int32_t B, H, ndir, _pad;
} dlsg_lstm_step_t;
int dlsg_lstm_step_supported(int32_t B, int32_t H);
int dlsg_lstm_step_fwd(const dlsg_lstm_step_t* p, void* stream);

/* ---- region -> frame aggregation of EncoderVisualGraphTUN (models/layer.py:184-192), fused -------------------------------
 * Replaces, per encoder e (E <= 2, both in one launch): obj_norm LayerNorm over the tanh'ed region projection Y, the
 * frame x region score product, the softmax over ALL T*R regions of a clip (layer.py:188, dim=1) and the weighted sum.
 *   O_r = LN(Y_r) ; S_tr = F_t . O_r ; A = softmax_r(scale * S) ; U_t = sum_r A_tr xhat_r ; agg_t = gamma o U_t + beta
 * fwd : one CTA per (clip, encoder) streams the (T*R, H) bf16 tile of Y ONCE (bulk copies into a 2-stage shared-memory
 *       ring), LayerNorm folded into the two mma.sync products, online softmax over the 32-row tiles.
 *       scores_only != 0: only St (= F . O, any fp32 F, e.g. the aggregate's gradient) and tconst are produced.
 * bwd : (after a scores_only pass with F := dA that yields dSm = dA . O and tcA) a small kernel turns the (T, T*R) matrices
 *       into per-tile mma operands and row scalars (softmax backward, closed-form LayerNorm-backward statistics), then one
 *       CTA per (256-column slice, clip, encoder) re-reads its slice of Y once and writes d(pre-activation of the region projection) as bf16, with the
 *       softmax backward, the LayerNorm backward (its row reductions in closed form from the T x T*R matrices) and the
 *       tanh derivative fused; dF = dA + gamma o V (the residual path of layer.py:192 included), obj_norm parameter
 *       gradients and the projection's bias gradient are ACCUMULATED with fp32 atomics.
 * Supported: H == 1024, T <= 26, bf16 Y with 16-byte aligned rows (dlsg_region_aggregate_supported).              */
typedef struct {
  const void* Y[2]; int64_t ldy;            /* bf16 (B*TR, H) per encoder, row pitch in elements                              */
  const float* F[2]; int64_t ldf;           /* fp32 (B*T, H)                                                                  */
  const float* gamma[2]; const float* beta[2];
  float* agg[2]; int64_t ldagg;             /* out fp32 (B*T, H) (unused when scores_only)                                    */
  float* U[2]; int64_t ldu;                 /* out, optional: aggregate before the affine part (the backward's dgamma needs it) */
  float* stats[2];                          /* out, optional (B*TR, 2): mean, rstd of every region row                        */
  float* St[2];                             /* out, optional (B, T, TR): raw scores F_t . O_r                                 */
  float* tconst[2];                         /* out, optional (B*T, 4): sum_h bf16(F*gamma), F.beta, and (full pass only) the
                                               softmax normalisers m_t, 1/l_t: A_tr = exp(scale S_tr - m_t) / l_t            */
  int32_t B, E, T, TR, H, scores_only; float scale; int32_t _pad;
} dlsg_region_agg_fwd_t;
typedef struct {
  const void* Y[2]; int64_t ldy;
  const float* stats[2];
  const float* St[2]; const float* dSm[2];                         /* (B, T, TR) fp32: raw scores, dA . O                    */
  const float* F[2]; int64_t ldf; const float* dA[2]; int64_t ldda; const float* U[2]; int64_t ldu;
  const float* tcF[2]; const float* tcA[2];                        /* tconst of the forward (F) and of the dA scores pass    */
  const float* gamma[2]; const float* beta[2];
  void* dpre[2]; int64_t ldd;                                      /* out bf16 (B*TR, H)                                     */
  float* dF[2]; int64_t lddf;                                      /* out fp32 (B*T, H) = dA + gamma o V                     */
  float* dgamma[2]; float* dbeta[2]; float* dbias[2];              /* (H) fp32, ACCUMULATED                                  */
  void* work[2];                                                   /* scratch, dlsg_region_aggregate_bwd_workspace() bytes each */
  int32_t B, E, T, TR, H, _pad; float scale; int32_t _pad2;
} dlsg_region_agg_bwd_t;
int dlsg_region_aggregate_supported(int32_t T, int32_t TR, int32_t H);
int dlsg_region_aggregate_fwd(const dlsg_region_agg_fwd_t* p, void* stream);
int64_t dlsg_region_aggregate_bwd_workspace(int32_t B, int32_t T, int32_t TR);   /* bytes per encoder, 16-byte aligned base */
int dlsg_region_aggregate_bwd(const dlsg_region_agg_bwd_t* p, void* stream);       /* two launches: prep + streaming pass */

/* ---- embedding (layer.py:421,438,535) and small reductions ---------------------------------- */
int dlsg_embedding_gather(const float* table, const int64_t* ids, int64_t ld_ids, int32_t rows, int32_t W,
                          void* out, int out_dtype, int64_t ldo, void* out2, int out2_dtype, int64_t ldo2,
                          float drop_p, uint64_t seed, uint64_t offset, void* stream);
int dlsg_embedding_scatter_add(float* dtable, const int64_t* ids, int64_t ld_ids, int32_t rows, int32_t W,
                               const float* dout, int64_t lddo, float drop_p, uint64_t seed, uint64_t offset,
                               void* stream);
/* y[b,:] = mean_p x[b,p,:] (layer.py:407-409); bwd: dx[b,p,:] += dy[b,:]/P                     */
int dlsg_mean_nodes_fwd(const float* x, int32_t B, int32_t P, int32_t H, float* y, int64_t ldy, void* stream);
int dlsg_mean_nodes_bwd(const float* dy, int64_t lddy, int32_t B, int32_t P, int32_t H, float* dx, void* stream);
/* generic elementwise: y = a*x + b*y (n elements) ; y = x + pe (broadcast over batch)           */
int dlsg_axpby(const float* x, float a, float* y, float b, int64_t n, void* stream);
int dlsg_add_rowbcast(const float* x, const float* pe, float* y, int64_t batch, int64_t inner, float drop_p, uint64_t seed,
                      uint64_t offset, void* stream);
/* inverted dropout y = x*mask/(1-p), mask from philox(seed, offset+i); backward = same call on dy (nn.Dropout sites:
 * layer.py:53,422,439,574,594, sublayer.py:25,60,104,186)                                        */
int dlsg_dropout(const float* x, float* y, int64_t n, float drop_p, uint64_t seed, uint64_t offset, void* stream);
/* ResBlock pieces (sublayer.py:117-119): r = relu(x) in place; dx = (x>0)*dr                    */
int dlsg_relu(float* x, int64_t n, void* stream);
int dlsg_relu_bwd(const float* r, const float* dr, float* dx, int64_t n, void* stream);
int dlsg_mul(const float* a, const float* b, float* y, int64_t n, void* stream);

/* ---- vocabulary rows: log-softmax / argmax / CE (layer.py:437,540; run_gun.py:189-197) ------ */
int dlsg_row_argmax(const float* logits, int64_t ld, int32_t rows, int32_t V, int64_t* ids, int64_t ld_ids, void* stream);
int dlsg_log_softmax(const float* logits, int64_t ld, int32_t rows, int32_t V, float* out, int64_t ldo, void* stream);
/* masked CE over (B,L,V): rows with t < len[b] count.  loss_sum/count are ACCUMULATED (zero first).
 * dlogits (may be NULL) = (softmax - onehot)*gscale/count_total for counted rows, 0 otherwise.   */
int dlsg_ce_masked(const float* logits, const int64_t* targets, const int32_t* lens, int32_t B, int32_t L, int32_t V,
                   float* loss_sum, float* dlogits, float inv_count, const float* inv_count_dev /* overrides if non-NULL */,
                   float* row_loss /* optional B*L scratch: per-row losses, summed in a fixed order (deterministic
                                      loss_sum); NULL = atomicAdd accumulation */,
                   void* stream);

/* ---- beam search (allennlp_beamsearch.py:127,186-260,272-292) ------------------------------- */
/* per row: log-softmax, force <end> if last==end (:186-190), top-k (value desc, lowest index on ties) */
int dlsg_beam_topk(const float* logits, int64_t ld, int32_t rows, int32_t V, const int64_t* last, int32_t end_index,
                   int32_t k, float* top_lp, int64_t* top_id, int32_t normalize /* 0: input already log-probs */, void* stream);
/* (B, beam*k) candidates + parent log-probs -> top beam: new log-probs, classes, back-pointers   */
int dlsg_beam_merge(const float* top_lp, const int64_t* top_id, const float* last_lp, int32_t B, int32_t beam, int32_t k,
                    float* new_lp, int64_t* new_cls, int64_t* backptr, int32_t* all_end, int32_t end_index, void* stream);
/* dst[b,j,:] = src[b, backptr[b,j], :] for contiguous state rows of row_bytes bytes (any dtype)    */
int dlsg_beam_gather(const void* src, void* dst, const int64_t* backptr, int32_t B, int32_t beam, int32_t row_bytes, void* stream);
/* the same for n <= 4 buffers in one launch (host arrays of n pointers / row sizes): the h / c rows of both LSTMs          */
int dlsg_beam_gather_multi(const void* const* src, void* const* dst, const int32_t* row_bytes, int32_t n, const int64_t* backptr,
                           int32_t B, int32_t beam, void* stream);
/* back-track: preds (S,B,beam), backs (S-1,B,beam) -> out (B,beam,S)                             */
int dlsg_beam_backtrack(const int64_t* preds, const int64_t* backs, int32_t S, int32_t B, int32_t beam, int64_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
