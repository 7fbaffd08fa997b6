/*
 * icd_b200.h — C ABI of the B200-native iCD hot path (libicd_b200.so).
 *
 * The reference (yandex-research/invertible-cd) has no native layer: every call below replaces a
 * torch / diffusers-0.25.1 library call that the reference reaches from
 *     utils/generation.py:241-244     model.unet(latents, t, timestep_cond=w_emb, encoder_hidden_states=ctx)
 *     utils/generation_sdxl.py:288-295, 445-453   pipe.unet(..., added_cond_kwargs=...)
 *     utils/p2p.py:321-342            explicit-probabilities attention (softmax(QK^T*scale) -> controller -> P.V)
 *     utils/generation.py:136-155     predicted_origin (consistency update)
 *     utils/generation.py:96-122      guidance_scale_embedding
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller (PyTorch); the library never allocates in a call
 *     (CUDA-graph safe) except for a host-side cache of TMA descriptors;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - activations are fp16, channels-last: an image tensor is [B][H][W][C] == a token matrix [B*H*W][C];
 *   - weights are fp16 [N][K] (K contiguous); 3x3 conv weights are [Cout][ky][kx][Cin];
 *   - every function returns 0 on success, non-zero on error; icd_last_error() returns the message
 *     (the Python host raises RuntimeError with it);
 *   - one in-flight call per stream; not thread-safe by contract (mirrors the reference's single-threaded use).
 */
#ifndef ICD_B200_H_
#define ICD_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

const char* icd_last_error(void);
/* library / device probe: returns 0 and fills sm_count, cc_major, cc_minor */
int icd_device_info(int* sm_count, int* cc_major, int* cc_minor);
int icd_abi_version(void);
/* Programmatic dependent launch (every kernel of the library is launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization and calls griddepcontrol.wait before its first global-memory
 * access, so consecutive kernels overlap prologue with tail — eagerly and inside captured CUDA graphs).
 * Default on (env ICD_PDL=0 disables); returns the previous setting. No reference counterpart: the reference's
 * eager PyTorch launches (utils/generation.py:241) are fully serialised. */
int icd_set_pdl(int enabled);

/* ------------------------------------------------------------------------------------------------
 * Dense contraction on tcgen05 tensor cores:  D = epilogue(alpha * A . B^T), fp32 accumulation.
 * Replaces F.linear / F.conv2d(3x3, 1x1) / torch.baddbmm / torch.bmm of the U-Net forward
 * (diffusers UNet2DConditionModel; call sites listed above).
 * ------------------------------------------------------------------------------------------------ */
typedef struct IcdGemm {
  /* A operand (fp16). a_mode 0: A[z2][z1][m][k] with strides; a_mode 1: NHWC image(s), implicit 3x3 conv, pad 1 */
  const void* a0;
  const void* a1;          /* optional second source, concatenated after a0 along K (channel concat); may be NULL */
  int a_mode;
  int K0, K1;              /* K (or channel) extent of a0 / a1 */
  long long a0_ld, a1_ld;  /* row (pixel) stride in elements */
  long long a_z1_stride, a_z2_stride; /* batch strides in elements (a_mode 0) */
  int ZA1;                 /* z = z2 * ZA1 + z1 */
  int B, H, W;             /* a_mode 1: image batch and spatial size; M = B*H*W */
  /* B operand (fp16): [N][K] K-contiguous (b_mn_major=0) or [K][N] N-contiguous (b_mn_major=1) */
  const void* b;
  long long b_ld;
  long long b_z1_stride, b_z2_stride;
  int ZB1;
  int b_mn_major;
  /* problem */
  int M, N, K;             /* K = total reduction length per filter tap source (K0+K1), conv multiplies by 9 */
  int Z;                   /* batch entries (1 for plain GEMM) */
  /* epilogue */
  float alpha;
  const float* bias;       /* [N] fp32 or NULL */
  const float* rowvec;     /* [M/rows_per_img][ldv] fp32 added per image, or NULL */
  int rows_per_img;
  int ldv;
  const void* residual;    /* fp16 [M][ldr] or NULL */
  long long ldr, res_zstride;
  void* out;               /* fp16 or fp32 */
  long long ldc, out_z1_stride, out_z2_stride, out_imgstride; /* batch offset = (z % ZA1)*z1 + (z / ZA1)*z2 */
  int out_fp32;
  int out_mode;            /* 0 row-major [M][ldc]; 1 transposed out[img][n][row_in_img] with column stride ldc */
  int geglu;               /* B rows packed per BN tile as [h | gate]; out[:, j] = h_j * gelu(g_j); N counts packed rows */
  int force_bn;            /* 0 = heuristic, else 64/128/160/256 */
  int force_bm;            /* 0 = heuristic, else 128/256 (rows per CTA tile) */
  int force_splits;        /* 0 = heuristic, else split-K factor (needs ws) */
  void* ws;                /* optional fp32 split-K workspace (device), may be NULL */
  long long ws_bytes;
  /* fused consistency update on the (transposed, fp32) output — utils/generation.py:136-155 */
  const float* upd_x;
  float* upd_out;
  float alpha_t, sigma_t, alpha_s, sigma_s;
  /* softmax-from-statistics epilogue (fp16 row-major output only): out = exp2(alpha*acc - s[0]) * s[1] with
   * s = exp_stats[(z*M + row)*2 ..] — the (scaled reference maximum, 1 / row sum) pairs icd_attention_ex wrote.
   * Materialises normalised attention probabilities (AttentionStore capture of self-attention maps,
   * utils/p2p.py:145-149) in ONE pass over the scores, without a separate softmax kernel. NULL = off. */
  const float* exp_stats;
} IcdGemm;

int icd_gemm(const IcdGemm* g, void* stream);
/* N-tile width the heuristic (or force_bn) picks — the weight packer needs it for the GEGLU row interleave */
int icd_gemm_pick_bn(int M, int N, int Z, int geglu, int b_mn_major, int force_bn);

/* ------------------------------------------------------------------------------------------------
 * Fused attention core (replaces F.scaled_dot_product_attention, and baddbmm+softmax+bmm of
 * utils/p2p.py:335-338 when the controller does not edit the probabilities).
 *   q: [B][Nq][H*D]  k,v: [B][Nk][H*D]  out: [B][Nq][H*D]   (row strides given in elements)
 *   probs_out (optional): [B*H][Nq][probs_ld] fp16 normalised probabilities (AttentionStore capture)
 * ------------------------------------------------------------------------------------------------ */
int icd_attention(const void* q, const void* k, const void* v, void* out, int B, int H, int Nq, int Nk, int D,
                  long long q_ld, long long k_ld, long long v_ld, long long out_ld, float scale, void* probs_out,
                  long long probs_ld, void* stream);
/* Same, plus stats_out (optional): fp32 [B*H][Nq][2] = (reference maximum * scale * log2(e), 1 / row sum) of the
 * online softmax, for any N_kv: icd_gemm's exp_stats epilogue turns Q.K^T into the normalised probabilities with them. */
int icd_attention_ex(const void* q, const void* k, const void* v, void* out, int B, int H, int Nq, int Nk, int D,
                     long long q_ld, long long k_ld, long long v_ld, long long out_ld, float scale, void* probs_out,
                     long long probs_ld, float* stats_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Memory-bound kernels (HBM roofline): normalisations, activations, layout, embeddings, update.
 * ------------------------------------------------------------------------------------------------ */
/* GroupNorm(32 groups) [+ SiLU] over NHWC fp16; optional second source concatenated along C.
 * stats_ws: fp32 workspace of B*4096 floats (per-chunk partial sums). Replaces F.group_norm + F.silu
 * (diffusers ResnetBlock2D.norm1/norm2, Transformer2DModel.norm, conv_norm_out).
 * One launch (single pass: the activation chunk of every CTA stays in shared memory between the statistics and the
 * normalisation) when the whole tensor fits on chip, otherwise two (statistics, apply).
 * The single-pass kernel is launched COOPERATIVELY (the driver co-schedules all of its CTAs or refuses the launch, in
 * which case the two-kernel path runs) and synchronises through cooperative-groups' grid barrier: no library-global
 * state, safe on concurrent streams and across CUDA-graph replays. 32 groups; channels per group 4 or >= 8. */
int icd_groupnorm(const void* x0, int C0, const void* x1, int C1, void* y, int B, int HW, int groups, float eps,
                  const float* gamma, const float* beta, int apply_silu, float* stats_ws, void* stream);
/* Number of kernel launches icd_groupnorm will use for this shape (1 = single pass, 2 = statistics + apply). */
int icd_groupnorm_launches(int B, int HW, int C);
/* LayerNorm over the last dim of [rows][C] fp16 -> fp16. Replaces F.layer_norm. */
int icd_layernorm(const void* x, void* y, int rows, int C, float eps, const float* gamma, const float* beta,
                  void* stream);
/* In-place row softmax over [rows][ld] fp16 (first `cols` entries of each row); fp32 statistics. */
int icd_softmax(void* x, long long rows, int cols, long long ld, void* stream);
/* Same with a causal mask: row r is query (r % causal_period) of its sequence and sees keys 0..(r % causal_period);
 * the masked tail is written as zeros (CLIP text transformer, transformers' CLIPAttention causal_attention_mask). */
int icd_softmax_causal(void* x, long long rows, int cols, long long ld, int causal_period, void* stream);
/* nearest-neighbour 2x upsample NHWC fp16 (Upsample2D interpolate). */
int icd_upsample2x(const void* x, void* y, int B, int H, int W, int C, void* stream);
/* gather for the stride-2 3x3 Downsample2D conv: y[B*Ho*Wo][9*C] from NHWC x (pad 1). */
int icd_im2col_s2(const void* x, void* y, int B, int H, int W, int C, void* stream);
/* same with the low-side padding explicit: pad = 1 is icd_im2col_s2; pad = 0 is the VAE encoder's Downsample2D
 * (diffusers pads (0,1,0,1) and convolves unpadded: zeros on the bottom / right edge only). */
int icd_im2col_s2_pad(const void* x, void* y, int B, int H, int W, int C, int pad, void* stream);
/* NCHW fp32 latent -> NHWC fp16 with channels zero-padded to Cpad (conv_in operand). */
int icd_latent_to_nhwc(const float* x, void* y, int B, int C, int HW, int Cpad, void* stream);
/* sinusoidal embeddings: diffusers Timesteps(dim, flip_sin_to_cos=True, freq_shift=0) -> [n][dim] fp16 = [cos | sin].
 * freqs[dim/2] (fp32) is the frequency table; the host computes it once with the reference's fp32 op order so the
 * sin/cos arguments are bit-identical to the reference's. */
int icd_timestep_embedding(const float* t, const float* freqs, void* y, int n, int dim, void* stream);
/* guidance_scale_embedding (utils/generation.py:96-122): [n] fp32 w -> [n][dim] fp16 = [sin | cos] of (1000 w) f_i */
int icd_guidance_embedding(const float* w, const float* freqs, void* y, int n, int dim, void* stream);
/* y = silu(x) elementwise fp16 (time-embedding MLP activations) */
/* y = act(x) over n fp16 elements (n even): kind 0 SiLU, 1 quick_gelu (x * sigmoid(1.702 x)), 2 exact-erf GELU —
 * the CLIP text encoders' MLP activations (utils/generation.py:286-303, utils/generation_sdxl.py:9-46 call them). */
int icd_act(const void* x, void* y, long long n, int kind, void* stream);
/* CLIPTextEmbeddings: out[r][:] = tok[ids[r]][:] + pos[r % T][:], fp16 [rows][C]; ids int64 on the device. */
int icd_embed_tokens(const long long* ids, const void* tok, const void* pos, void* out, long long rows, int T, int C,
                     int vocab, void* stream);
int icd_silu(const void* x, void* y, long long n, void* stream);
/* fp16 -> fp16 elementwise add:  y = a + b */
int icd_add(const void* a, const void* b, void* y, long long n, void* stream);
/* consistency update (predicted_origin, utils/generation.py:136-155), fp32 NCHW, per-sample t/s scalars */
int icd_consistency_update(const float* eps, const float* x, float* out, long long per_sample, int B,
                           const float* alpha_t, const float* sigma_t, const float* alpha_s, const float* sigma_s,
                           void* stream);

/* ------------------------------------------------------------------------------------------------
 * fp32 validation path (ABI 3). `load_models(dtype='fp32')` — the dtype the reference runs SD1.5 editing in
 * (running/sd1.5/launch_editing_iCD_sd1.5.sh:38, utils/loading.py:38-41) — executes the SAME U-Net executor on these
 * kernels: fp32 operands and activations, contractions on the FMA pipe with fp32 accumulation, same packed layouts
 * and epilogue semantics as the fp16 entry points above. A correctness mode (element-wise agreement with the fp32
 * oracle), not a fast path. Each call replaces the fp32 torch op the fp16 entry point of the same name replaces.
 * ------------------------------------------------------------------------------------------------ */
typedef struct IcdSgemm {
  /* out[z][m][n] = alpha * sum_k A[z][m][k] * Bm[z][n][k] (+ bias[n]) (+ rowvec[m / rows_per_img][n]) (+ residual) */
  const float* a0;       /* A = [a0 | a1] along K (per filter tap when conv = 1) */
  const float* a1;
  int C0, C1;            /* columns (channels) of a0 / a1 */
  long long a0_ld, a1_ld;
  int conv;              /* 0: plain rows; 1: implicit 3x3 / pad 1 / stride 1 over NHWC images, K = (tap, channel) */
  int B, H, W;           /* conv: image geometry, M = B*H*W */
  int w_tap_ld;          /* conv: weight columns per filter tap (the packed, padded C_in) */
  const float* b;        /* weights / second operand */
  long long b_ld;
  int b_kn;              /* 0: b is [N][K] (K contiguous); 1: b is [K][N] (P.V with V as it lies in memory) */
  int M, N, K;           /* K is implied (C0 + C1, or 9 * (C0 + C1)) and ignored on input */
  int Z, ZH;             /* batch index z -> (zb, zh) = (z / ZH, z % ZH) */
  long long a_zb, a_zh, b_zb, b_zh, c_zb, c_zh, r_zb, r_zh;   /* element offsets per batch index (a0, b, out, residual) */
  float alpha;
  const float* bias;     /* [N] or NULL */
  const float* rowvec;   /* [images][ldv] added to every row of image m / rows_per_img, or NULL */
  int rows_per_img;
  long long ldv;
  const float* residual; /* [M][ldr] or NULL */
  long long ldr;
  float* out;
  long long ldc;
  int vec;               /* internal (set by the library) */
} IcdSgemm;
int icd_sgemm_f32(const IcdSgemm* g, void* stream);
int icd_groupnorm_f32(const float* x0, int C0, const float* x1, int C1, float* y, int B, int HW, int groups, float eps,
                      const float* gamma, const float* beta, int silu, void* stream);
int icd_layernorm_f32(const float* x, float* y, int rows, int C, float eps, const float* gamma, const float* beta,
                      void* stream);
/* in place over the first `cols` entries of `rows` rows of stride ld; entries [cols, ld) become 0 */
int icd_softmax_f32(float* x, long long rows, int cols, long long ld, void* stream);
int icd_silu_f32(const float* x, float* y, long long n, void* stream);
/* y[m][j] = h * gelu(g) of a [M][2F] projection whose rows are interleaved per bn-wide tile (packing.pack_geglu) */
int icd_geglu_f32(const float* x, float* y, long long M, int F, int bn, void* stream);
int icd_upsample2x_f32(const float* x, float* y, int B, int H, int W, int C, void* stream);
int icd_im2col_s2_f32(const float* x, float* y, int B, int H, int W, int C, void* stream);
/* NCHW -> NHWC with channels zero-padded to Cpad; NHWC (row stride ld, leading C channels) -> NCHW */
int icd_nchw_to_nhwc_f32(const float* x, float* y, int B, int C, int HW, int Cpad, void* stream);
int icd_nhwc_to_nchw_f32(const float* x, long long ld, float* y, int B, int C, int HW, void* stream);
/* y[r] = sin_first ? [sin a | cos a] : [cos a | sin a], a = (v[r] * scale) * freqs[k]: Timesteps (scale 1, cos first)
 * and guidance_scale_embedding (scale 1000, sin first; utils/generation.py:96-122) with fp32 output */
int icd_sincos_embedding_f32(const float* v, const float* freqs, float* y, int n, int dim, float scale, int sin_first,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ICD_B200_H_ */
