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

#define M3R_ABI_VERSION 1

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
} m3r_gemm_args;

int m3r_gemm(const m3r_gemm_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * LayerNorm over the last dim of fp32 rows (optionally of x + add), output 16-bit (GEMM operand) or fp32.
 * Replaces nn.LayerNorm call sites: must3r/model/blocks/layers.py:42,46,65,70,71,76, encoder.py:37,
 * decoder.py:76, feedback_mechanism.py:14 (eps 1e-5), and the `new_mem[l] + offset` add of
 * feedback_mechanism.py:49-51 when add != NULL.
 * ------------------------------------------------------------------------------------------------- */
int m3r_layernorm(const float* x, int64_t ldx, const float* add, int64_t ldadd, const float* gamma,
                  const float* beta, float eps, int32_t M, int32_t D, void* out, int64_t ldo, int32_t out_dtype,
                  int32_t is_bf16, void* stream);

/* fp32 -> 16-bit cast of [M,D] rows (encoder features entering the decoder projector, decoder.py:274). */
int m3r_cast16(const float* x, int64_t ldx, int32_t M, int32_t D, void* out, int64_t ldo, int32_t is_bf16,
               void* stream);

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
} m3r_attn_args;

int m3r_attention(const m3r_attn_args* args, void* stream);

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

/* y[M,D] (fp32) += row-broadcast add... : x_out = x + off (fp32), used by tests only through layernorm. */

#ifdef __cplusplus
}
#endif
#endif /* MUST3R_B200_H */
