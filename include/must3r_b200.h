/*
 * must3r_b200 — C ABI of the B200-native MUSt3R inference hot path (libm3r_b200.so).
 *
 * Plain pointers and sizes only (no torch types).  All pointers are DEVICE pointers unless noted;
 * buffers are borrowed for the duration of the call; work is enqueued on `stream` (the caller's current
 * CUDA stream, as a cudaStream_t cast to void*).  Every entry point returns 0 on success or a non-zero
 * code; m3r_last_error() returns a human-readable message for the calling thread's last failure.
 *
 * Each function names the reference interface it replaces (paths relative to the reference repo).
 * 16-bit operands are fp16 (is_bf16 = 0) or bf16 (is_bf16 = 1); accumulation, LayerNorm, softmax and the
 * residual stream are fp32.
 */
#ifndef MUST3R_B200_H
#define MUST3R_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M3R_ABI_VERSION 4
#define M3R_MAX_PEERS 8
#define M3R_MAX_GROUPS 16

/* m3r_decoder_call.mem_mode (MEMORY_MODES, must3r/model/blocks/layers.py:9) */
#define M3R_MEM_KV 0
#define M3R_MEM_NORM_Y 1
#define M3R_MEM_RAW 2
/* out_dtype */
#define M3R_OUT_F32 0
#define M3R_OUT_16 1
/* activation */
#define M3R_ACT_NONE 0
#define M3R_ACT_GELU 1

const char* m3r_last_error(void);
int m3r_abi_version(void);
/* Number of kernels this library has launched in this process (bench.py reports it as gpu_launches). */
long long m3r_launch_count(void);
/* Optional per-kernel device timing for bench.py: CUDA events are recorded on the launch stream around every
 * GEMM / attention / LayerNorm kernel while enabled.  m3r_prof_read synchronises and fills
 * out[cat*4 + {0: ms, 1: launches, 2: algorithmic flops, 3: algorithmic bytes}] for the 7 categories
 * gemm_kernel<256>, gemm_kernel<128>, gemm_kernel<64>, attn_kernel<QT=2>, attn_kernel<QT=1> (key splits merged in-kernel), layernorm,
 * other (28 doubles). */
void m3r_prof_enable(int on);
int m3r_prof_read(double* out);
/* Debug hook (tools/trace_attn.py, tools/trace_gemm.py): device buffer that subsequent attention / GEMM launches fill
 * with %globaltimer stamps (64 / 16 uint64 per CTA: entry, barrier waits, epilogue, exit); NULL = off.  Only builds made
 * with M3R_TRACE=1 carry the stamps; otherwise the call fails with an error. */
int m3r_debug_trace(void* buf);

/* ---------------------------------------------------------------------------------------------------
 * Linear layer y = act(x W^T + b) (+ residual) on tcgen05 tensor cores.
 * Replaces nn.Linear / croco Mlp call sites: must3r/model/blocks/attention.py:88-89,108-111,
 * dust3r/croco/models/blocks.py:67-70, must3r/model/decoder.py:50, must3r/model/blocks/head.py:67,
 * and, with rope_tab != NULL, the fused qkv projection + curope.rope_2d
 * (dust3r/croco/models/curope/curope.cpp:49) of Attention.forward (attention.py:92-96).
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  const void* A;          /* [M,K] 16-bit, row-major, leading dim lda (elements) */
  int64_t lda;
  const void* W;          /* [N,K] 16-bit, row-major (nn.Linear.weight layout), leading dim ldw */
  int64_t ldw;
  int32_t M, N, K;        /* N % 64 == 0, K % 64 == 0 */
  int32_t is_bf16;
  const float* bias;      /* [N] fp32 or NULL */
  int32_t act;            /* M3R_ACT_* (applied after bias/rope, before residual) */
  const float* residual;  /* [M,N] fp32 (leading dim ldr) added after the activation, or NULL */
  int64_t ldr;
  const float* rowbias;   /* [N] fp32 added to rows with (row % rb_period) >= rb_first, or NULL
                             (image2_embed, must3r/model/decoder.py:282,287) */
  int32_t rb_period, rb_first;
  const float* rope_tab;  /* [rope_period,64] fp32 = per token (cosY[16], sinY[16], cosX[16], sinX[16]);
                             rotates columns < rope_cols in 64-wide heads; or NULL */
  int32_t rope_cols, rope_period;
  void* out;              /* [M,N] fp32 or 16-bit, leading dim ldc (elements) */
  int64_t ldc;
  int32_t out_dtype;      /* M3R_OUT_* */
  int32_t rows_per_batch; /* output row remap for appending into [B,cap,N] buffers:              */
  int64_t batch_stride_rows; /* out_row = (row / rows_per_batch) * batch_stride_rows + row % rows_per_batch;
                                rows_per_batch <= 0 disables the remap */
  /* Fused "GEMM -> all-gather": when n_peer_out > 0 the 16-bit output tile is ALSO stored, with the same row mapping
   * and ldc, to these device pointers - the same buffer in the other GPUs of the NVSwitch domain, mapped into this
   * process with m3r_ipc_open (peer stores travel over NVLink while the kernel is still computing other tiles). */
  int32_t n_peer_out;
  void* peer_out[M3R_MAX_PEERS];
  /* 1 = W is not written by the launch that precedes this one in the stream (model weights): the kernel then requests
   * its first weight tiles before the programmatic-dependency wait, hiding their HBM latency behind the predecessor's
   * tail.  0 = no assumption. */
  int32_t w_static;
  /* LayerNorm fused into the PRODUCING GEMM (must3r/model/blocks/layers.py:42,46,65,70,71,76 call sites): when norm_out
   * != NULL the output must be the fp32 residual stream x [M,N] with N = the model width (N % 64 == 0, N <= 768) and the
   * epilogue also writes normalised rows (x - mean) * rsqrt(var + norm_eps) as 16-bit to norm_out (leading dim ldn): the
   * A operand of the next GEMM, whose weight / bias carry the LayerNorm affine (W * diag(gamma), b + W beta).  The CTAs
   * holding the column tiles of one 128-row block exchange per-row (mean, M2) partials through device memory and meet
   * at a device-scope counter, so the whole problem must fit in one wave (ceil(M/128) * N/64 <= SM count) and such
   * launches must not run concurrently on two streams of one device. */
  void* norm_out;
  int64_t ldn;
  float norm_eps;
} m3r_gemm_args;

int m3r_gemm(const m3r_gemm_args* args, void* stream);

/* Grouped form: `groups` problems of identical shape [args->M, args->N, args->K] in one launch.  A = [groups*M, K]
 * (group g = rows [g*M, (g+1)*M)), W rows [g*w_group_rows, +N), bias + g*bias_group; args->out / peer_out are ignored,
 * group g writes out[g] (and peer_out[g*M3R_MAX_PEERS + r], r < args->n_peer_out) with args' ldc / row remap.
 * Replaces the per-level K|V projections of the memory append (must3r/model/decoder.py:323-330). */
typedef struct {
  int32_t groups;
  int64_t w_group_rows;
  int64_t bias_group;
  void* out[M3R_MAX_GROUPS];
  void* peer_out[M3R_MAX_GROUPS * M3R_MAX_PEERS];
  int32_t max_ctas;           /* > 0: run on at most this many SMs (persistent tile loop).  The memory append runs on a side
                                 stream next to the last decoder block; its CTAs (slow when they store over NVLink) must not
                                 take the SMs the LayerNorm-emitting GEMMs of the main chain need to be co-resident */
} m3r_gemm_group;

int m3r_gemm_grouped(const m3r_gemm_args* args, const m3r_gemm_group* grp, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * LayerNorm over the last dim of fp32 rows (optionally of x + add), output 16-bit (GEMM operand) or fp32.
 * Replaces nn.LayerNorm call sites: must3r/model/blocks/layers.py:42,46,65,70,71,76, encoder.py:37,
 * decoder.py:76, feedback_mechanism.py:14 (eps 1e-5), and the `new_mem[l] + offset` add of
 * feedback_mechanism.py:49-51 when add != NULL.
 * ------------------------------------------------------------------------------------------------- */
int m3r_layernorm(const float* x, int64_t ldx, const float* add, int64_t ldadd, const float* gamma,
                  const float* beta, float eps, int32_t M, int32_t D, void* out, int64_t ldo, int32_t out_dtype,
                  int32_t is_bf16, void* stream);

/* Affine-free LayerNorm: out16 = ((x [+ add]) - mean) * rsqrt(var + eps) as 16-bit rows = the A operand of a GEMM whose
 * weights carry the LayerNorm affine (W * diag(gamma), b + W beta).  `add` (may be NULL) is applied to rows < add_rows and
 * repeats with period add_period rows: one launch normalises new_mem[l] + offset of every decoder level
 * (must3r/model/feedback_mechanism.py:49-51, decoder.py:323-330).  Used where the producing GEMM cannot emit the
 * normalised rows itself (m3r_gemm_args.norm_out: multi-wave problems). */
int m3r_normalize16(const float* x, int64_t ldx, const float* add, int64_t ldadd, int32_t add_rows, int32_t add_period,
                    float eps, int32_t M, int32_t D, void* out16, int64_t ldo, int32_t is_bf16, void* stream);

/* fp32 -> 16-bit cast of [M,D] rows (encoder features entering the decoder projector, decoder.py:274). */
int m3r_cast16(const float* x, int64_t ldx, int32_t M, int32_t D, void* out, int64_t ldo, int32_t is_bf16,
               void* stream);

/* LayerNorm of 16-bit rows -> 16-bit rows (memory_mode 'raw': norm_y applied to the stored tokens at use,
 * must3r/model/blocks/layers.py:92), and fp32 (x + add) -> 16-bit (what 'raw' stores, layers.py:82-83 after the feedback
 * offset of decoder.py:323-330). */
int m3r_layernorm16(const void* x16, int64_t ldx, const float* gamma, const float* beta, float eps, int32_t M, int32_t D,
                    void* out16, int64_t ldo, int32_t is_bf16, void* stream);
int m3r_add_cast16(const float* x, int64_t ldx, const float* add, int64_t ldadd, int32_t M, int32_t D, void* out,
                   int64_t ldo, int32_t is_bf16, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * RoPE trig table: tab[t] = (cos, sin)(pos[t,axis] * F0 / base^(d/16)), d<16, for axis y then x.
 * Same arithmetic as dust3r/croco/models/curope/kernels.cu:45-55 (fp32 powf/cosf/sinf).
 * pos: [T,2] int64 (y,x).
 * ------------------------------------------------------------------------------------------------- */
int m3r_rope_table(const int64_t* pos, int32_t T, float base, float f0, float* tab, void* stream);

/* Stand-alone in-place 2-D RoPE with the curope operator contract (curope.cpp:49-69):
 * tokens [B,N,H,D=64] 16-bit or fp32 with strides (sB,sN,sH,1) in elements, pos [B,N,2] int64. */
int m3r_rope_2d(void* tokens, int32_t dtype /*0 f32,1 f16,2 bf16*/, int32_t B, int32_t N, int32_t H, int32_t D,
                int64_t sB, int64_t sN, int64_t sH, const int64_t* pos, float base, float fwd, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * FlashAttention-style softmax(Q K^T / 8) V, head_dim 64, on tcgen05 (S and P·V in TMEM, online softmax).
 * Replaces CoreAttention.attention (must3r/model/blocks/attention.py:37-79) for both self-attention and
 * the memory cross-attention, including the "drop each image's own new tokens" mask of
 * MUSt3R.make_mem_mask (must3r/model/decoder.py:119-139) and the torch.cat / expand / boolean gather of
 * decoder.py:306-317, which become two key/value segments plus a skip range.
 *
 * Query rows of (batch b, head h): Q + (b*Nq + i)*ldq + h*64.
 * Keys come from up to two segments; segment s of batch b: rows K_s + ((b / kv_group) * kv_bstride_s + j)*ldk_s
 * (+ h*64), j < Nk_s.  kv_group query batches share one K/V batch (the views of one scene).
 * Keys with global index in [skip_lo + (b % kv_group)*skip_step, +skip_len) are not attended (skip_len = 0:
 * no mask); global index = j for segment 0 and Nk_0 + j for segment 1.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  const void* Q; int64_t ldq;
  const void* K0; const void* V0; int64_t ldk0; int64_t kv_bstride0; int32_t Nk0;
  const void* K1; const void* V1; int64_t ldk1; int64_t kv_bstride1; int32_t Nk1;
  void* O; int64_t ldo;     /* [B*Nq, H*64] 16-bit */
  int32_t B, H, Nq;
  int32_t kv_group;         /* >= 1 */
  int32_t skip_lo, skip_step, skip_len;
  int32_t is_bf16;
  float scale;              /* 1/sqrt(64) */
  /* Export mode (both non-NULL; O may then be NULL): write the UNNORMALISED attention state of this key set instead of
   * the normalised output - export_o [B*Nq, H*64] fp32 = sum_j 2^(s_ij*c - m_i) v_j, export_ml [B*Nq, H, 2] fp32 =
   * (m_i, l_i) with c = scale*log2(e), m_i the running maximum in log2 units (-inf if the row saw no key) - so that key
   * shards held by different GPUs can be merged exactly (m3r_attn_merge): context-parallel memory cross-attention. */
  float* export_o;
  float* export_ml;
} m3r_attn_args;

int m3r_attention(const m3r_attn_args* args, void* stream);

/* Merge `n` exported attention states of the same queries over disjoint key sets:
 * out[r, c] = sum_k o_k[r, c] 2^(m_k[r,h] - m) / sum_k l_k[r,h] 2^(m_k[r,h] - m), m = max_k m_k, h = c / 64.
 * parts_o[k] / parts_ml[k]: device pointers as exported above (rows x H*64 fp32, rows x H x 2 fp32); out: [rows, H*64]
 * 16-bit with leading dim ldo.  m3r_attn_state_fill writes the state of an EMPTY key set (o = 0, m = -inf, l = 0). */
int m3r_attn_merge(const float* const* parts_o, const float* const* parts_ml, int32_t n, int64_t rows, int32_t H,
                   void* out, int64_t ldo, int32_t is_bf16, void* stream);
int m3r_attn_state_fill(float* export_o, float* export_ml, int64_t rows, int32_t H, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Patch embedding front end: im2col of [V,3,H,W] fp32 images into [V*N, 768] 16-bit rows
 * (col = c*256 + i*16 + j, the flattened Conv2d(3,D,16,16) weight order, dust3r/croco/models/blocks.py:222),
 * to be followed by m3r_gemm with the [D,768] weight.  Tokens in raster order (patch_embed.py:20-29).
 * ------------------------------------------------------------------------------------------------- */
int m3r_im2col16(const float* img, int32_t V, int32_t H, int32_t W, void* out, int32_t is_bf16, void* stream);

/* Head unpatchify (must3r/tools/image.py:9-14 + head.py:71): proj [V*N,1792] fp32 ->
 * out[v,16y+i,16x+j,c] = proj[v, y*w+x, c*256+16i+j], out [V,H,W,7] fp32. */
int m3r_unpatchify(const float* proj, int32_t V, int32_t H, int32_t W, int32_t C, float* out, void* stream);

/* postprocess(compute_cam=False) (must3r/engine/inference.py:16-27, must3r/tools/geometry.py:14-18):
 * pm [P,7] -> pts3d [P,3], pts3d_local [P,3], conf [P]. */
int m3r_postprocess(const float* pm, int64_t P, float* pts3d, float* pts3d_local, float* conf, void* stream);

/* Nearest-neighbour distance of Q query points [Q,3] fp32 to a database [P,3] fp32: out[i] = min_j |q_i - p_j| (inf for an
 * empty database).  Replaces the per-frame scipy KD-tree query of the SLAM keyframe test
 * (must3r/slam/nns.py:57-62, called from must3r/slam/model.py:82). */
int m3r_nn_min_dist(const float* queries, int32_t Q, const float* db, int64_t P, float* out, void* stream);

/* ===================================================================================================
 * Whole-model entry points: one call enqueues every kernel of a forward pass from C++ (no per-layer
 * Python/ctypes overhead).  Weight structs hold DEVICE pointers (16-bit matrices in nn.Linear layout,
 * fp32 biases / LayerNorm affines); the `blocks` arrays themselves live in HOST memory.
 * ================================================================================================= */

/* Block (must3r/model/blocks/layers.py:36-54): norm1, attn.{qkv,proj}, norm2, mlp.{fc1,fc2} */
typedef struct {
  const float* norm1_w; const float* norm1_b;
  const void* qkv_w; const float* qkv_b;       /* [3D,D] */
  const void* proj_w; const float* proj_b;     /* [D,D]  */
  const float* norm2_w; const float* norm2_b;
  const void* fc1_w; const float* fc1_b;       /* [4D,D] */
  const void* fc2_w; const float* fc2_b;       /* [D,4D] */
} m3r_enc_block;

/* Dust3rEncoder (must3r/model/encoder.py:13-52) */
typedef struct {
  int32_t embed_dim, depth, num_heads, mlp_hidden;
  float ln_eps, rope_base, rope_f0;
  int32_t is_bf16;
  const void* patch_w; const float* patch_b;   /* [D,768] = patch_embed.proj.weight.view(D,-1) */
  const m3r_enc_block* blocks;                 /* host array [depth] */
  const float* norm_w; const float* norm_b;    /* norm_enc */
} m3r_encoder_weights;

int64_t m3r_encoder_workspace_bytes(const m3r_encoder_weights* w, int32_t V, int32_t H, int32_t W);

/* Dust3rEncoder.forward (encoder.py:46-52): img [V,3,H,W] fp32 -> x [V*N, D] fp32.
 * pos: [N,2] int64 (y,x) positions of ONE view's tokens (all views of a call share the grid). */
int m3r_encoder_forward(const m3r_encoder_weights* w, const float* img, int32_t V, int32_t H, int32_t W,
                        const int64_t* pos, float* out_x, void* workspace, int64_t workspace_bytes, void* stream);

/* CachedDecoderBlock (must3r/model/blocks/layers.py:57-99).  Every LayerNorm of the block is applied WITHOUT its affine
 * (by the epilogue of the GEMM that produced the rows, m3r_gemm_args.norm_out, or by m3r_normalize16); the affine lives in the
 * consuming Linear: W' = W * diag(gamma), b' = b + W beta (exact in real arithmetic; rounded to 16 bits once, like W).
 * Because norm1 and norm_y normalise the same block input, the self-attention qkv projection and the K|V projection of
 * the new memory tokens (layers.py:81-88,91) become ONE GEMM with the stacked weight a_w. */
typedef struct {
  const void* a_w; const float* a_b;           /* [5D,D]: rows [0,3D) attn.qkv (norm1 folded), rows [3D,5D) cross_attn.projk ;
                                                  projv (norm_y folded).  blocks[l].a_w == blocks[0].a_w + l*5D*D and
                                                  blocks[l].a_b == blocks[0].a_b + l*5D let the memory append run as one
                                                  grouped GEMM; otherwise it falls back to one GEMM per level */
  const void* proj_w; const float* proj_b;     /* attn.proj [D,D] */
  const void* q_w; const float* q_b;           /* cross_attn.projq [D,D], norm2 folded */
  const void* cproj_w; const float* cproj_b;   /* cross_attn.proj [D,D] */
  const void* fc1_w; const float* fc1_b;       /* mlp.fc1 [4D,D], norm3 folded */
  const void* fc2_w; const float* fc2_b;       /* mlp.fc2 [D,4D] */
  const float* normy_w; const float* normy_b;  /* norm_y affine: memory_mode norm_y / raw only (rows stored / normalised at use) */
  const void* kv_w; const float* kv_b;         /* UNFOLDED [projk ; projv] [2D,D]: memory_mode norm_y / raw (K|V of stored rows
                                                  projected at use, layers.py:92-96); may be NULL for memory_mode kv */
} m3r_dec_block;

/* MUSt3R (must3r/model/decoder.py:14-156) */
typedef struct {
  int32_t enc_dim, embed_dim, depth, num_heads, mlp_hidden, out_dim;
  float ln_eps, fb_ln_eps, rope_base, rope_f0;
  int32_t is_bf16;
  int32_t feedback;                            /* 0 none, 1 single_mlp, 2 single_linear */
  const void* embed_w; const float* embed_b;   /* feat_embed_enc_to_dec [D,enc_dim] */
  const float* image2_embed;                   /* [D] */
  const m3r_dec_block* blocks;                 /* host array [depth] */
  const void* fb1_w; const float* fb1_b;       /* feedback_layer.fc1 [4D,D] (or the single linear [D,D]), feedback_norm folded */
  const void* fb2_w; const float* fb2_b;       /* feedback_layer.fc2 [D,4D] */
  const void* head_w; const float* head_b;     /* head_dec.proj [out_dim,D], norm_dec folded */
} m3r_decoder_weights;

/* One aspect-ratio group of a decoder call (MUSt3R.forward_list, decoder.py:158): B scenes x n_views views */
typedef struct {
  int32_t n_views, N, H, W;
  const float* x_enc;      /* [B*n_views*N, enc_dim] fp32 encoder features */
  const int64_t* pos;      /* [B*n_views*N, 2] int64 */
  float* pointmaps;        /* out: [B*n_views, H, W, out_dim/256] fp32 */
} m3r_dec_group;

typedef struct {
  int32_t B, G;
  const m3r_dec_group* groups;     /* host array [G] */
  int32_t Nm;                      /* memory tokens per scene before this call */
  const void* const* mem;          /* host array [depth] of device ptrs: [B, >=Nm, 2D] 16-bit (NULL if Nm==0) */
  int64_t mem_bstride_rows;        /* rows between consecutive scenes in mem[l] */
  int32_t render;                  /* 1: read-only pass (decoder.py:307); 0: memory update */
  int32_t is_init;                 /* current_mem is None: view 0 of group 0 gets no image2_embed (decoder.py:176,280) */
  void* const* mem_out;            /* update only: host array [depth] of device ptrs [B, >=Nm+Nt, 2D]; rows [0,Nm) are
                                      copied from mem[l] unless mem_out[l]==mem[l]; the new rows follow */
  int64_t mem_out_bstride_rows;
  int32_t new_only;                /* update only: 1 = mem_out[l] is [B, Nt, 2D] and receives ONLY the new post-feedback
                                      K|V rows (no copy of the old memory) - used by the sharded schedule, where the new
                                      tokens of all ranks are all-gathered into a pre-allocated memory buffer */
  int32_t n_peers;                 /* > 0 (with new_only): the post-feedback K|V GEMM epilogue stores the new rows straight
                                      into every rank's memory buffer (fused GEMM -> all-gather over NVLink peer memory) */
  void* const* peer_mem;           /* host array [n_peers * depth]: peer_mem[r * depth + l] = rank r's memory buffer of
                                      level l ([1, cap, 2D] 16-bit), already offset to the row where THIS rank's tokens go */
  int32_t mem_mode;                /* M3R_MEM_KV: memory rows are K|V (2D wide); M3R_MEM_NORM_Y: rows are norm_y(x) (D wide) and
                                      K|V are projected at use; M3R_MEM_RAW: rows are x (D wide), norm_y + projection at use
                                      (must3r/model/blocks/layers.py:81-96).  All widths "2D" above become D for the last two.
                                      peer output (n_peers > 0) needs M3R_MEM_KV. */
  /* ---- context-parallel memory cross-attention (no reference counterpart): cp_world > 1 means the
   * memory tokens of the ONE scene (B == 1, mem_mode kv) are SHARDED over cp_world GPUs, each running this same call on the
   * same views.  mem[] / Nm describe only THIS rank's rows (Nm may be 0); every rank computes the attention state of the
   * queries over its shard (m3r_attn_args.export_o), copies it into slot cp_rank of every rank's staging buffer
   * (m3r_peer_bcast over NVLink), the ranks meet at a device-side flag barrier (m3r_peer_signal / m3r_peer_wait, epoch
   * cp_epoch0 + layer + 1) and merge the cp_world states (m3r_attn_merge): exactly the softmax over the union of the
   * shards.  New tokens attend to each other (calls with several views) on rank 0.  Only the rank with cp_owner = 1
   * appends the call's new rows (mem_out); the others skip the feedback MLP and the append.  The caller guarantees that
   * the scene's memory is non-empty globally (the very first call of a scene runs without cp). */
  int32_t cp_world, cp_rank, cp_owner;
  void* const* cp_stage;           /* host array [cp_world]: rank q's staging buffer as mapped here: 2 * cp_world slots */
  int64_t cp_slot_bytes;           /* bytes per slot, >= m3r_decoder_cp_slot_bytes() */
  void* const* cp_flag_slots;      /* host array [cp_world]: this rank's uint32 slot in rank q's flag array */
  const void* cp_flags_local;      /* this rank's own flag array */
  uint32_t cp_epoch0;              /* the call consumes epochs cp_epoch0 + 1 .. cp_epoch0 + depth */
} m3r_decoder_call;

/* bytes of one staging slot for a call with M token rows: M * H * (64 * 4 + 8), rounded up to 256 */
int64_t m3r_decoder_cp_slot_bytes(const m3r_decoder_weights* w, int64_t M);

int64_t m3r_decoder_workspace_bytes(const m3r_decoder_weights* w, const m3r_decoder_call* call);

/* ---------------------------------------------------------------------------------------------------
 * Peer memory for the multi-GPU schedule (one process per GPU): buffers allocated with m3r_peer_alloc can be exported
 * as a 64-byte CUDA IPC handle, exchanged through torch.distributed, and mapped by the other ranks of the node.
 * ------------------------------------------------------------------------------------------------- */
int m3r_peer_alloc(int64_t bytes, void** ptr);
int m3r_peer_free(void* ptr);
int m3r_ipc_export(void* ptr, void* handle64);
int m3r_ipc_open(const void* handle64, void** ptr);
int m3r_ipc_close(void* ptr);
/* Device-side barrier of the multi-GPU schedule (no reference counterpart; replaces a host-blocking collective per update
 * round).  Each rank owns slot `rank` of a uint32[world] flag array that lives in every rank's peer-visible memory.
 * m3r_peer_signal(flag_slots = address of MY slot in each of the n ranks' arrays, value): enqueued after the kernels whose
 * peer stores must be visible; writes `value` (a monotonically increasing epoch) with system-scope release semantics.
 * m3r_peer_wait(local_flags = this rank's own array, rank_mask, value): a kernel that spins on the GPU until every slot r
 * with bit r of rank_mask set has reached `value`; work enqueued after it sees the rows those ranks stored. */
/* Copy `bytes` (multiple of 16) from src to the same-sized block dsts[k], k < n: this GPU's staging buffer and its peers'
 * (NVLink stores): the exchange step of the context-parallel cross-attention. */
int m3r_peer_bcast(const void* src, void* const* dsts, int32_t n, int64_t bytes, void* stream);
int m3r_peer_signal(void* const* flag_slots, int32_t n, uint32_t value, void* stream);
int m3r_peer_wait(const void* local_flags, uint32_t rank_mask, uint32_t value, void* stream);

/* MUSt3R.forward / forward_list (decoder.py:158-350) for memory_mode 'kv'. */
int m3r_decoder_forward(const m3r_decoder_weights* w, const m3r_decoder_call* call, void* workspace,
                        int64_t workspace_bytes, void* stream);


#ifdef __cplusplus
}
#endif
#endif /* MUST3R_B200_H */
