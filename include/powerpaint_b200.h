/*
 * powerpaint_b200.h — C ABI of the B200-native PowerPaint denoising hot path.
 *
 * The reference (open-mmlab/PowerPaint) is pure Python over diffusers; it has no FFI of
 * its own. The boundary a maintainer binds is therefore the set of *operators* its
 * per-step hot path executes, each entry point citing the reference call site it
 * replaces (paths relative to the reference checkout):
 *
 *   pp_gemm_conv      nn.Conv2d 3x3 / 1x1 and nn.Linear inside ResnetBlock2D,
 *                     Transformer2DModel, Downsample2D, Upsample2D, conv_in/conv_out,
 *                     BrushNet zero-convs  (powerpaint/models/unet_2d_blocks.py:789,807,
 *                     1274,1289,1319,1428,2499,2514,2542,2672,2689;
 *                     unet_2d_condition.py:256,477; BrushNet_CA.py:223-228,330-376,446-454)
 *                     with bias / time-embedding broadcast / skip add / BrushNet add /
 *                     output scale / GEGLU fused into the epilogue
 *                     (unet_2d_condition.py:1223,1300; unet_2d_blocks.py:1388-1398,2627-2638)
 *   pp_attention      Attention + AttnProcessor2_0 -> F.scaled_dot_product_attention
 *                     (processors referenced at unet_2d_condition.py:24-31)
 *   pp_group_norm     nn.GroupNorm(32)+SiLU in ResnetBlock2D / Transformer2DModel.norm /
 *                     conv_norm_out (unet_2d_condition.py:466,1351-1353), also performs the
 *                     up-path torch.cat (unet_2d_blocks.py:2589,2732) while normalising
 *   pp_layer_norm     nn.LayerNorm x3 per BasicTransformerBlock
 *   pp_upsample2x     Upsample2D's F.interpolate(scale_factor=2, mode="nearest")
 *   pp_time_embed     Timesteps(320, flip_sin_to_cos=True, shift=0)
 *                     (unet_2d_condition.py:914-938)
 *   pp_cfg_ddim_step  CFG combine + DDIMScheduler.step + next-step input build
 *                     (pipelines/pipeline_PowerPaint.py:990-996,1018-1023;
 *                      pipeline_PowerPaint_Brushnet_CA.py:1390,1444-1449)
 *   pp_unipc_step     CFG combine + UniPCMultistepScheduler.step (the v2 app's scheduler, app.py:197)
 *   pp_upsample_nearest  Upsample2D with an explicit output size (unet_2d_condition.py:1120-1126,1311-1312)
 *                     Also carries: LayerNorm of BasicTransformerBlock folded into the producing / consuming
 *                     GEMMs (row_stats / ln_stats; diffusers LayerNorm norm1/2/3), and
 *                     to_q | to_k | to_v^T of a self-attention as one launch (PP_EPI_ROWS_THEN_TRANSPOSED).
 *                     Long-K launches run as 2-CTA clusters (tcgen05 cta_group::2) on their own.
 *   pp_gemm_stats_geometry  host-only: layout of the GroupNorm partial sums a GEMM / conv emits from its
 *                     epilogue (pp_gemm_desc.chan_stats -> pp_gn_desc.part0 / part1)
 *   pp_gemm_row_stats_records, pp_gemm_splitk_bytes  host-only: scratch sizes of the LayerNorm-statistics
 *                     producer and of the optional split-K workspace
 *   pp_softmax_rows, pp_image_preprocess_u8, pp_image_postprocess   VAE attention softmax and the uint8
 *                     pre/post-processing either side of vae.encode / vae.decode
 *                     (pipeline_PowerPaint.py:39-153,657-669,1051,1062)
 *   pp_embed_gather, pp_causal_attention_small   CLIP text encoder: embeddings with the task-prompt splice
 *                     (utils/utils.py:256-483) and the causal 77-token attention (pipeline_PowerPaint.py:317-518)
 *   pp_program_*      a recorded list of the above, replayed per denoising step (pp_program_run_range: a slice
 *                     of it as plain launches, for bisecting a step op by op)
 *                     (the `for i, t in enumerate(timesteps)` loops,
 *                      pipeline_PowerPaint.py:988-1041, Brushnet_CA.py:1384-1466,
 *                      ControlNet.py:1663-1741)
 *
 * Conventions: all pointers are DEVICE pointers unless named host_*; activations are
 * bf16, channels-last (NHWC == [batch, tokens, channels]); every call enqueues on the
 * given cudaStream_t and never synchronises the host; no allocation happens inside a
 * call; a non-zero return code means nothing was enqueued and pp_last_error() holds the
 * reason. Handles are thread-compatible (one thread per handle at a time).
 */
#ifndef POWERPAINT_B200_H_
#define POWERPAINT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int pp_status; /* 0 = ok, 1 = invalid argument, 2 = CUDA error, 3 = unsupported */
typedef void* pp_stream; /* cudaStream_t */

const char* pp_last_error(void);
/* ABI version of this library (bumped on incompatible struct changes). */
int pp_abi_version(void);
/* 1 if the current device is sm_100 (B200) and the kernels can run on it. */
int pp_device_supported(void);

/* ------------------------------------------------------------------ GEMM / conv */
enum {
    PP_A_MATRIX = 0,    /* A is [M, K] row-major (a Linear or 1x1 conv over tokens) */
    PP_A_CONV3X3 = 1,   /* implicit GEMM, 3x3 stride 1 pad 1 over NHWC [nb, h, w, c] */
    PP_A_CONV3X3_S2 = 2,  /* implicit GEMM, 3x3 stride 2 pad 1 (Downsample2D); output ceil(h/2) x ceil(w/2) */
    PP_A_CONV3X3_S2P0 = 3 /* 3x3 stride 2 over F.pad(x, (0,1,0,1)) (the VAE encoder's Downsample2D(padding=0));
                             output floor(h/2) x floor(w/2) */
};
enum {
    PP_EPI_PLAIN = 0,
    PP_EPI_GEGLU = 1,      /* out[:, j] = (acc_a + bias_a) * gelu(acc_g + bias_g); weights tile-interleaved */
    PP_EPI_TRANSPOSED = 2, /* out[(m / t_rows) * N + n][m % t_rows], row pitch t_ld (V^T for attention) */
    PP_EPI_ROWS_THEN_TRANSPOSED = 3 /* columns [0, trans_from_col) row-major into `out` (pitch ldc), columns
                               [trans_from_col, N) transposed into `out_t` like PP_EPI_TRANSPOSED with
                               N - trans_from_col channels: to_q | to_k | to_v^T of one attention in ONE launch */
};
enum { PP_ACT_NONE = 0, PP_ACT_SILU = 1, PP_ACT_QUICK_GELU = 2 /* x * sigmoid(1.702 x): CLIP text encoder MLP */ };

/* Geometry of the per-channel GroupNorm partial sums a GEMM / conv can emit from its epilogue
   (pp_gemm_desc.chan_stats) and pp_group_norm consumes (pp_gn_desc.part0 / part1): one record
   {sum of (x - shift), sum of (x - shift)^2, shift, 0} per (m-tile, sample segment of the tile, output channel),
   fp32, laid out [m_tiles][segs][N][4]; shift = the segment's first row, so the variance does not cancel for a
   large mean over a small spread. Filled by pp_gemm_stats_geometry(). */
typedef struct pp_stats_geom {
    int32_t supported;   /* 0: this GEMM cannot emit statistics (the consumer runs its own statistics pass) */
    int32_t channels;    /* N */
    int32_t segs;        /* samples per m-tile */
    int32_t seg_rows;    /* rows of one sample inside an m-tile */
    int32_t tiles_per_group; /* m-tiles that cover one group of `segs` samples */
    int32_t tiles_x, tiles_y, bw, bh, wo, ho; /* pixel-tile geometry (matrix mode: a 1-D strip) */
    int64_t bytes;       /* size of the partials buffer */
} pp_stats_geom;

typedef struct pp_gemm_desc {
    int32_t a_mode;
    int32_t epilogue;
    /* A operand: up to two sources concatenated along K / channels (skip-concat without torch.cat) */
    const void* a0;
    const void* a1;      /* may be NULL */
    int32_t c0, c1;      /* channels (K per tap) of each source; c1 = 0 if a1 is NULL */
    int64_t lda0, lda1;  /* PP_A_MATRIX: row pitch in elements; conv modes: ignored (tight NHWC) */
    int32_t nb, h, w;    /* conv modes: INPUT batch / height / width */
    int32_t M;           /* PP_A_MATRIX: rows. conv modes: ignored (= nb*ho*wo) */
    int32_t N;           /* GEMM N (for GEGLU: 2x the output width) */
    /* B operand: weights [N, Kw] bf16 row-major, Kw = taps * (pad64(c0) + pad64(c1)) when
       taps == 9 or a1 != NULL, else row pitch ldb >= c0 */
    const void* b;
    int64_t ldb;
    /* epilogue: v = acc + bias[n] + rowvec[m / rows_per_group][n] + res1[m][n];
                 v = v * alpha + res2[m][n]; v = act(v) */
    const float* bias;
    const float* rowvec;
    int32_t rows_per_group;
    int64_t rowvec_ld; /* row pitch of rowvec in elements; 0 = N */
    const void* res1; /* bf16 */
    int64_t ldr1;
    const void* res2; /* bf16 */
    int64_t ldr2;
    float alpha;
    int32_t act;
    void* out;
    int64_t ldc;
    int32_t out_fp32; /* 0: bf16 output, 1: fp32 output */
    int32_t t_rows;   /* PP_EPI_TRANSPOSED */
    int64_t t_ld;
    int32_t block_n;  /* 0 = auto, else one of 64/128/160/256 */
    int32_t t_fp16;   /* PP_EPI_TRANSPOSED: store fp16 instead of bf16 (V^T for pp_attention) */
    /* alpha is multiplied by alpha_dev[(alpha_step ? *alpha_step : 0) * alpha_stride] when alpha_dev != NULL:
       the BrushNet / ControlNet conditioning scale x per-step keep flag lives in a device table, so one
       recorded program serves every scale (Brushnet_CA.py:1369-1376,1403-1409; ControlNet.py:1652-1658) */
    const float* alpha_dev;
    const int32_t* alpha_step;
    int32_t alpha_stride;
    /* optional GroupNorm partial sums of the stored output (bf16 row-major outputs only), see pp_stats_geom;
       rows_per_group must hold the rows per sample in PP_A_MATRIX mode */
    float* chan_stats;
    /* LayerNorm folded into the GEMMs either side of it (BasicTransformerBlock norm1/2/3):
       PRODUCER (PP_A_MATRIX, bf16 row-major output): row_stats != NULL makes the epilogue emit, per output row and
       per half n-tile, one record {sum of (x - shift), sum of (x - shift)^2, shift, count} (fp32 x 4) of the values
       it computes, laid out [records][row_stats_ld] with records = pp_gemm_row_stats_records(desc) — scratch of this
       launch. The CTA that finishes the LAST n-tile of a 128-row block (row_ticket[m-tile], zero-initialised int32,
       left at zero again) folds the records of its rows (Chan's combination) into
         row_final[m] = {rstd, -rstd * mean}   (fp32 x 2, eps = ln_eps).
       CONSUMER: ln_stats != NULL (a producer's row_final) applies LayerNorm(A) algebraically in the epilogue. With
       W' = W * gamma (folded into b on the host), ln_u[n] = sum_k W'[n, k] and bias[n] = sum_k W[n, k] beta[k]
       (+ the layer's own bias):  acc' = rstd[m] * acc - rstd[m] mean[m] * ln_u[n];  everything after (bias, GEGLU
       gate, transposed store ...) is unchanged. */
    float* row_stats;
    int64_t row_stats_ld;
    float* row_final;
    int32_t* row_ticket;
    const float* ln_stats;
    const float* ln_u;
    float ln_eps;
    /* PP_EPI_ROWS_THEN_TRANSPOSED: transposed destination and the first transposed column (a multiple of the tile
       width: pass block_n explicitly). t_rows must be a multiple of 128 and divide M. */
    void* out_t;
    int32_t trans_from_col;
    /* optional split-K x2 for long-K launches whose tiles fill at most half of the GPU (8x8-resolution layers): fp32
       workspace of pp_gemm_splitk_bytes(desc) bytes and the same number of tiles of zero-initialised int32 flags (the
       kernel leaves them at zero). NULL: never split. Results are deterministic (owner + donor, fixed order). */
    float* splitk_ws;
    int32_t* splitk_flags;
} pp_gemm_desc;

pp_status pp_gemm_conv(const pp_gemm_desc* d, pp_stream stream);
/* host-only query: can this GEMM emit GroupNorm partial sums, and with which layout */
pp_status pp_gemm_stats_geometry(const pp_gemm_desc* d, pp_stats_geom* out);
/* host-only query: bytes of split-K workspace this launch could use (0: it would not split); flags: one int32 per
   (m-tile, n-tile), i.e. bytes / (128 * block_n * 4) of them — pass `*tiles_out` to size the flag array */
int64_t pp_gemm_splitk_bytes(const pp_gemm_desc* d, int32_t* tiles_out);
/* host-only query: number of per-row LayerNorm records this GEMM emits through row_stats (0: it cannot emit them) */
int32_t pp_gemm_row_stats_records(const pp_gemm_desc* d);

/* ------------------------------------------------------------------ attention */
typedef struct pp_attn_desc {
    /* q: [batch, nq, heads, d] with token pitch q_ld (elements); k likewise with nk, k_ld.
       vt: V transposed, [batch, heads*d, vt_ld] (keys contiguous, vt_ld >= nk, vt_ld % 8 == 0), bf16 or fp16.
       out: [batch, nq, heads*d] row pitch o_ld. softmax(q k^T * scale) v, no mask. */
    const void* q;
    const void* k;
    const void* vt;
    void* out;
    int32_t batch, heads, d, nq, nk;
    int64_t q_ld, k_ld, vt_ld, o_ld;
    int64_t q_batch_stride, k_batch_stride; /* elements between batches */
    float scale;
    int32_t vt_fp16; /* != 0: vt holds fp16 (P is then computed in fp16 too); 0: bf16 */
} pp_attn_desc;

pp_status pp_attention(const pp_attn_desc* d, pp_stream stream);

/* ------------------------------------------------------------------ norms */
typedef struct pp_gn_desc {
    /* GroupNorm over NHWC x = concat(x0[..., :c0], x1[..., :c1]) -> y [.., c0+c1] bf16 */
    const void* x0;
    const void* x1; /* may be NULL */
    int32_t c0, c1;
    int32_t batch, hw, groups;
    const float* gamma;
    const float* beta;
    float eps;
    int32_t silu;
    float* stats; /* scratch of pp_group_norm_scratch_bytes() bytes, 16-byte aligned: the per-(sample, group)
                     sum / sum of squares [batch, groups, 2] fp32, then the ticket counters and per-block
                     partials of the deterministic reduction */
    void* y;
    int32_t stats_prezeroed; /* 1: the scratch was zeroed once when it was allocated (every call leaves its
                                ticket counters zero again) or by one memset per step over all scratch;
                                0: the call clears its ticket counters with a memset of its own */
    /* from_partials != 0: the statistics pass is replaced by a small finalize kernel that folds the
       per-tile channel sums the producing GEMMs emitted (part0 for x0, part1 for x1) into
       (mean, rstd) per (sample, group) — Chan's parallel combination in fp64, fixed order. `stats` then
       only needs batch * groups * 2 floats. */
    int32_t from_partials;
    const float* part0;
    const float* part1;
    pp_stats_geom geom0, geom1;
} pp_gn_desc;
pp_status pp_group_norm(const pp_gn_desc* d, pp_stream stream);
/* bytes of `stats` scratch a GroupNorm over [batch, hw, channels] with `groups` groups needs (0: invalid) */
int64_t pp_group_norm_scratch_bytes(int32_t batch, int32_t hw, int32_t channels, int32_t groups);

pp_status pp_layer_norm(const void* x, void* y, const float* gamma, const float* beta,
                        int32_t rows, int32_t c, float eps, pp_stream stream);

/* ------------------------------------------------------------------ small ops */
/* nearest 2x upsample of NHWC bf16 [nb,h,w,c] -> [nb,2h,2w,c] */
pp_status pp_upsample2x(const void* x, void* y, int32_t nb, int32_t h, int32_t w, int32_t c,
                        pp_stream stream);
/* F.interpolate(x, size=(ho, wo), mode="nearest") on NHWC bf16 — Upsample2D with an explicit output size
   (unet_2d_condition.py:1120-1126,1311-1312: latent sizes that are not multiples of 8) */
pp_status pp_upsample_nearest(const void* x, void* y, int32_t nb, int32_t h, int32_t w, int32_t c,
                              int32_t ho, int32_t wo, pp_stream stream);
/* y = a + b (bf16, n elements, n % 8 == 0) — ControlNet skip residuals */
pp_status pp_add(const void* a, const void* b, void* y, int64_t n, pp_stream stream);
/* sinusoidal timestep embedding [batch, dim] bf16: cat(cos, sin) (flip_sin_to_cos=True, shift 0).
   t comes from timesteps[*step_idx] when step_idx != NULL, else timesteps[0..batch). */
pp_status pp_time_embed(const float* timesteps, const int32_t* step_idx, void* out, int32_t batch,
                        int32_t dim, pp_stream stream);
/* NCHW fp32 <-> NHWC bf16 (c padded to c_pad with zeros on the way in) */
pp_status pp_nchw_to_nhwc(const float* x, void* y, int32_t nb, int32_t c, int32_t hw,
                          int32_t c_pad, pp_stream stream);
pp_status pp_nhwc_to_nchw(const void* x, int32_t x_is_fp32, float* y, int32_t nb, int32_t c,
                          int32_t hw, int32_t c_ld, pp_stream stream);

/* ------------------------------------------------------------------ VAE attention / image I/O */
/* p[r, :] = softmax(s[r, :]): fp32 scores in, bf16 probabilities out, cols <= 16384. The VAE mid-block
   attention (one head of 512 channels; pipeline_PowerPaint.py:657-669,:1051 through vae.encode / decode) runs as
   S = Q K^T (pp_gemm_conv, fp32 out, alpha = 1/sqrt(512)) -> pp_softmax_rows -> O = P V (pp_gemm_conv). */
pp_status pp_softmax_rows(const float* s, void* p, int64_t rows, int32_t cols, int64_t ld_s, int64_t ld_p,
                          pp_stream stream);
/* uint8 NCHW [nb,3,h,w] (+ mask [nb,1,h,w]: mask_mode 1 = uint8, 2 = fp32, 0 = none) -> bf16 NHWC
   [nb, hw, c_pad]: (px / divisor + shift) * (mask < 0.5), channels >= 3 zero — `prepare_mask_and_masked_image`
   (pipeline_PowerPaint.py:39-153) on the device */
pp_status pp_image_preprocess_u8(const uint8_t* image, const void* mask, int32_t mask_mode, void* out, int32_t nb,
                                 int32_t hw, int32_t c_pad, float divisor, float shift, pp_stream stream);
/* decoded image NHWC (channels 0..2 of c_ld) -> clamp(x/2 + 0.5, 0, 1) as uint8 NHWC [nb,hw,3] (x255, rounded)
   and / or fp32 NCHW [nb,3,hw] — `VaeImageProcessor.postprocess` (pipeline_PowerPaint.py:1062) */
pp_status pp_image_postprocess(const void* x, int32_t x_is_fp32, int32_t c_ld, uint8_t* out_u8, float* out_f32,
                               int32_t nb, int32_t hw, pp_stream stream);

/* ------------------------------------------------------------------ text encoder */
/* out[r] = (idx[r] < vocab ? base[idx[r]] : ext[idx[r] - vocab]) + pos[r % seq]  (bf16 [rows, dim]; fp32 tables).
   CLIP token + position embedding with the EmbeddingLayerWithFixes task-prompt splice (utils.py:256-483) resolved
   by the host into one gather index per position. */
pp_status pp_embed_gather(const int32_t* idx, const float* base, const float* ext, const float* pos, void* out,
                          int32_t rows, int32_t vocab, int32_t seq, int32_t dim, pp_stream stream);
/* causal softmax(q k^T * scale) v over seq <= 128 tokens; qkv [batch*seq, 3*heads*d] bf16 (q | k | v),
   out [batch*seq, heads*d] — the CLIP text transformer's self-attention (pipeline_PowerPaint.py:317-518) */
pp_status pp_causal_attention_small(const void* qkv, void* out, int32_t batch, int32_t seq, int32_t heads, int32_t d,
                                    float scale, pp_stream stream);

/* ------------------------------------------------------------------ CFG + DDIM */
typedef struct pp_cfg_ddim_desc {
    /* eps: model output for the 2*batch CFG-duplicated samples (unconditional half first),
       NHWC [2*batch, hw, eps_ld] (fp32 or bf16), channels 0..3 used.
       latents: fp32 NHWC [batch, hw, 4], updated in place (x_t -> x_{t-1}).
       coef: [n_steps, 8] fp32 rows {sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev-sigma^2),
             sigma, 0, 0, 0}; row *step_idx is used, then *step_idx is incremented when
             advance_step != 0 (so a captured CUDA graph can be replayed once per step).
       noise: optional fp32 NHWC [batch, hw, 4] variance noise (eta > 0).
       next_in: optional bf16 NHWC [n_copies*batch, hw, next_c]: channels 0..3 = x_{t-1},
             channels 4..4+extra_c = extra[b] (mask + masked-image latents, fp32 NHWC
             [batch, hw, extra_c]), rest zero; written for both CFG halves. */
    const void* eps;
    int32_t eps_fp32;
    int32_t eps_ld;
    float* latents;
    const float* coef;
    int32_t* step_idx;
    int32_t advance_step;
    const float* noise;
    float guidance_scale;
    int32_t do_cfg; /* 0: eps has `batch` samples and is used as is */
    int32_t batch, hw;
    void* next_in;
    int32_t next_c, n_copies;
    const float* extra;
    int32_t extra_c;
    int32_t guidance_from_coef; /* != 0: guidance scale = coef[row][5] (graph-replay friendly) */
    int32_t extra_per_copy;     /* != 0: extra is [n_copies*batch, hw, extra_c] (one set per CFG half) */
    /* 4-channel-UNet blend after the step (pipeline_PowerPaint.py:1025-1035):
         x = (1 - m) * (a * x0 + sqrt(1 - a^2) * noise_b) + m * x,  a = coef[row][7] (sqrt(alpha_bar) of the NEXT
       timestep, 1 for the last step); x0 / m are the FIRST image's latents [hw, 4] / latent mask [hw] (the
       reference indexes `[:1]` and broadcasts over the batch), noise_b [batch, hw, 4]. NULL = no blend. */
    const float* blend_x0;
    const float* blend_mask;
    const float* blend_noise;
} pp_cfg_ddim_desc;
pp_status pp_cfg_ddim_step(const pp_cfg_ddim_desc* d, pp_stream stream);

/* CFG combine + UniPCMultistepScheduler.step (solver_order 2, bh2, predict_x0, epsilon prediction; the v2 app's
   scheduler, app.py:197) + next-step input build. The multistep update is linear in the tensors it touches, so
   the host folds the schedule into per-step scalars (ucoef rows of 12 floats):
     m_t    = u[0] * x + u[1] * eps                                   (x0 prediction, `convert_model_output`)
     x_c    = u[2] != 0 ? u[3] * last + u[4] * m1 + u[5] * m2 + u[6] * m_t : x      (UniC corrector)
     x_next = u[7] * x_c + u[8] * m_t + u[9] * m1                                     (UniP predictor)
     last <- x_c, m2 <- m1, m1 <- m_t, latents <- x_next
   guidance scale = coef[row][5] of the shared 8-float table. State tensors are fp32 NHWC [batch, hw, 4]. */
typedef struct pp_unipc_desc {
    const void* eps;
    int32_t eps_fp32;
    int32_t eps_ld;
    float* latents;
    float* last_sample;
    float* m1;
    float* m2;
    const float* coef;   /* [n_steps, 8], column 5 = guidance scale */
    const float* ucoef;  /* [n_steps, 12] */
    int32_t* step_idx;
    int32_t advance_step;
    int32_t do_cfg;
    int32_t batch, hw;
    void* next_in;       /* bf16 NHWC [n_copies*batch, hw, next_c]: channels 0..3 refreshed */
    int32_t next_c, n_copies;
} pp_unipc_desc;
pp_status pp_unipc_step(const pp_unipc_desc* d, pp_stream stream);

/* ------------------------------------------------------------------ programs */
/* A program records op descriptors once (tensor maps are encoded at record time) and
   replays them with one call per denoising step; it can be instantiated as a CUDA graph. */
typedef struct pp_program pp_program;
pp_status pp_program_create(pp_program** out);
void pp_program_destroy(pp_program* p);
pp_status pp_program_add_gemm(pp_program* p, const pp_gemm_desc* d);
pp_status pp_program_add_attention(pp_program* p, const pp_attn_desc* d);
pp_status pp_program_add_group_norm(pp_program* p, const pp_gn_desc* d);
pp_status pp_program_add_layer_norm(pp_program* p, const void* x, void* y, const float* gamma,
                                    const float* beta, int32_t rows, int32_t c, float eps);
pp_status pp_program_add_upsample2x(pp_program* p, const void* x, void* y, int32_t nb, int32_t h,
                                    int32_t w, int32_t c);
pp_status pp_program_add_upsample_nearest(pp_program* p, const void* x, void* y, int32_t nb, int32_t h,
                                          int32_t w, int32_t c, int32_t ho, int32_t wo);
pp_status pp_program_add_add(pp_program* p, const void* a, const void* b, void* y, int64_t n);
pp_status pp_program_add_time_embed(pp_program* p, const float* timesteps,
                                    const int32_t* step_idx, void* out, int32_t batch, int32_t dim);
pp_status pp_program_add_cfg_ddim(pp_program* p, const pp_cfg_ddim_desc* d);
pp_status pp_program_add_unipc(pp_program* p, const pp_unipc_desc* d);
pp_status pp_program_add_memset(pp_program* p, void* ptr, int64_t bytes);
pp_status pp_program_add_embed_gather(pp_program* p, const int32_t* idx, const float* base, const float* ext,
                                      const float* pos, void* out, int32_t rows, int32_t vocab, int32_t seq, int32_t dim);
pp_status pp_program_add_causal_attention_small(pp_program* p, const void* qkv, void* out, int32_t batch, int32_t seq,
                                                int32_t heads, int32_t d, float scale);
pp_status pp_program_add_softmax_rows(pp_program* p, const float* s, void* out, int64_t rows, int32_t cols,
                                      int64_t ld_s, int64_t ld_p);
int32_t pp_program_num_ops(const pp_program* p);
/* number of kernel launches one run enqueues */
int32_t pp_program_num_launches(const pp_program* p);
/* enqueue every recorded op on `stream` (plain launches) */
pp_status pp_program_run(pp_program* p, pp_stream stream);
/* diagnostic: replay only ops [first, first + count) as plain launches */
pp_status pp_program_run_range(pp_program* p, int32_t first, int32_t count, pp_stream stream);
/* capture the program into a CUDA graph (once), then replay it with pp_program_graph_launch */
pp_status pp_program_graph_build(pp_program* p, pp_stream stream);
pp_status pp_program_graph_launch(pp_program* p, pp_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* POWERPAINT_B200_H_ */
