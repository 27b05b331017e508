/* mmdgan_b200 -- C ABI of the B200-native MMD-GAN (SNGan + repulsive MMD) training hot path.
 *
 * The reference (richardwth/MMD-GAN) has no FFI layer of its own: every device kernel on this path is a TensorFlow-1.8
 * library call.  Each entry point below therefore names the reference CALL SITE it replaces (file:line under
 * /root/reference).  The Python host mirror in mmd-gan_b200/ (layer_func / math_func / my_sngan) binds these symbols
 * with ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions: extern "C"; every function returns int (0 = ok, < 0 = MMDGAN_E*); mmdgan_last_error() gives the
 * thread-local message; all tensor arguments are caller-owned DEVICE pointers to contiguous memory (16-byte aligned),
 * fp32 unless typed mmdgan_bf16; `stream` is a cudaStream_t passed as void*; no allocation, no synchronisation and no
 * host copies inside, so every call is CUDA-graph capturable; workspaces are sized by the *_workspace() queries and
 * provided by the caller.
 *
 * Internal GEMM-operand format ("planes"): NHWC, [plane][N*H*W][C] raw bf16 bits, C in {8, 16, 32} or a multiple of 64,
 * consecutive planes *_plane ELEMENTS apart (a multiple of 8).  An fp32 value x is carried as
 *   p0 = bf16(x), p1 = bf16(x - p0), p2 = bf16(x - p0 - p1):
 * three planes hold the fp32 value (error <= 2^-25 |x|), two planes 16 significand bits.  The tcgen05 kernels multiply
 * plane pairs with fp32 accumulation: npass = 6 -> {00,01,10,02,20,11} (fp32-grade products: forward passes, whose
 * errors the MMD loss amplifies), npass = 3 -> {00,01,10} (~2^-17: input / weight gradients, linear in the operands),
 * npass = 1 -> {00} (plain bf16 speed mode).  A launch reads the first 3 / 2 / 1 planes of its operands.
 *
 * fp16 plane formats for the operands of FORWARD launches (`*_fmt` fields): fp16 has 11 significand bits, so TWO planes of
 * scale * x carry 22 bits and npass = 3 is already fp32-grade (~2^-21) -- half the tensor work and two thirds of the
 * operand bytes of the six-product bf16 mode.  MMDGAN_FMT_F16A (activations, spectral-norm vectors) stores 16 * x,
 * MMDGAN_FMT_F16W (packed forward weights) 64 * x; the power-of-two factors keep the second plane of ordinary magnitudes in
 * fp16's normal range and leave head-room up to |x| < 4094 / 1023 (conversions saturate, they never produce inf, and the
 * producers report a saturation through the optional `sat_flag` so that the host can fail loudly); the
 * caller folds 1 / (16 * 64) into alpha_k.  The two operands of one MMA must have the SAME element type (measured: an fp16 x
 * bf16 descriptor is an illegal instruction), so gradient launches read bf16 planes only (mmdgan_convert_planes).
 */
#ifndef MMDGAN_B200_H
#define MMDGAN_B200_H

#include <stddef.h>

typedef unsigned short mmdgan_bf16; /* raw 16-bit plane element (bf16 bits, or fp16 bits in the MMDGAN_FMT_F16* formats) */
#define MMDGAN_FMT_BF16 0
#define MMDGAN_FMT_F16A 1 /* two fp16 planes of 16 * value */
#define MMDGAN_FMT_F16W 2 /* two fp16 planes of 64 * value */
#define MMDGAN_F16A_SCALE 16.0f
#define MMDGAN_F16W_SCALE 64.0f

#ifdef __cplusplus
extern "C" {
#endif

#define MMDGAN_OK 0
#define MMDGAN_EINVAL (-1) /* bad argument (null pointer, unsupported size, unknown loss type) */
#define MMDGAN_ESHAPE (-2) /* inconsistent shapes / alignment */
#define MMDGAN_EARCH (-3)  /* not an sm_100 device */
#define MMDGAN_ECUDA (-4)  /* CUDA runtime / driver error */
#define MMDGAN_ENCCL (-5)  /* collective error (reserved) */

const char* mmdgan_last_error(void);
int mmdgan_version(void);
/* 0 if the current device can run the kernels (compute capability 10.x), MMDGAN_EARCH otherwise */
int mmdgan_check_device(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Layout at the boundary.  Replaces the NCHW float32 batch contract of ReadTFRecords
 * (GeneralTools/input_func.py:837-868) and tf.concat of real + generated batches (DeepLearning/my_sngan.py:244-256):
 * the real batch is written straight into rows [0, B) of the discriminator's 2B input buffer. */
int mmdgan_nchw_to_nhwc(const float* src, mmdgan_bf16* dst, long long dst_plane, int npl, int fmt, int N, int C, int H, int W, int Cpad,
                        void* stream);
int mmdgan_nhwc_to_nchw(const mmdgan_bf16* src, long long src_plane, int npl, int fmt, float* dst, int N, int C, int H, int W, int Cpad,
                        void* stream);
/* fp32 [n] <-> bf16 planes [npl][n] in the same element order */
int mmdgan_to_planes(const float* x, mmdgan_bf16* dst, long long dst_plane, int npl, int fmt, long long n, void* stream);
int mmdgan_from_planes(const mmdgan_bf16* src, long long src_plane, int npl, int fmt, float* out, long long n, void* stream);
/* planes in one format -> planes in another (the weight-gradient GEMM needs bf16 planes of the fp16 forward activations:
 * an MMA cannot mix an fp16 operand with a bf16 one) */
int mmdgan_convert_planes(const mmdgan_bf16* src, long long src_plane, int src_npl, int src_fmt, mmdgan_bf16* dst, long long dst_plane,
                          int dst_npl, int dst_fmt, long long n, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Weight packing: canonical reference layouts (conv [k,k,Cin,Cout] layer_func.py:584, transposed conv
 * [k,k,Cout,Cin] layer_func.py:595, dense [in,out] layer_func.py:577) -> K-major GEMM operand [planes][classes *
 * rows_pad][kpad] of npl bf16 planes (3 for a forward operand, 2 for an input-gradient operand). */
#define MMDGAN_PACK_CONV_FWD 0
#define MMDGAN_PACK_CONV_DGRAD_S1 1
#define MMDGAN_PACK_CONV_DGRAD_S2 2
#define MMDGAN_PACK_TC_FWD 3
#define MMDGAN_PACK_TC_DGRAD 4
#define MMDGAN_PACK_DENSE_FWD 5
#define MMDGAN_PACK_DENSE_DGRAD 6
typedef struct mmdgan_pack_desc {
    const float* w;
    mmdgan_bf16* out;
    long long plane;
    int npl, fmt; /* fmt: MMDGAN_FMT_BF16 or MMDGAN_FMT_F16W */
    int mode, k, Cin, Cout, Cs, rows_pad, kpad, classes;
    int in_C, in_HW, out_C, out_HW; /* dense: NCHW-flatten <-> NHWC-flatten feature permutation (HW <= 1: identity) */
} mmdgan_pack_desc;
int mmdgan_pack_weights(const mmdgan_pack_desc* d, void* stream);
/* every parameter-derived buffer of a net in ONE launch after an optimiser update: `jobs_device` is an array of at most 256
 * jobs in DEVICE memory (kind 0 = pack weights, kind 1 = permute / pad a per-feature vector), built once at start-up; job j
 * owns the blocks [block_start_j, block_start_{j+1}) of the flat grid of `total_blocks` blocks (block_start ascending, first 0;
 * a pack job gets one block per 32 x 64 tile of its packed operand) */
typedef struct mmdgan_refresh_job {
    int kind, block_start;
    mmdgan_pack_desc pack;
    const float* src;
    float* dst;
    int n, C, HW, inverse;
} mmdgan_refresh_job;
int mmdgan_refresh(const mmdgan_refresh_job* jobs_device, int njobs, long long total_blocks, void* stream);
int mmdgan_permute_features(const float* src, float* dst, int n, int C, int HW, int inverse, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Gather-GEMM on tcgen05 tensor cores: tf.matmul / tf.nn.conv2d / tf.nn.conv2d_transpose forward
 * (GeneralTools/layer_func.py:909-928) and their input gradients (DeepLearning/my_sngan.py:301-304), with the
 * kernel * (act_k / sigma) scaling (layer_func.py:884-887), bias add (946-952), activation (104-167) or activation
 * derivative, bf16 plane split and per-tile column sums fused into the epilogue. */
typedef struct mmdgan_gemm_class {
    int oy, ox, ooy, oox, wrow, pad0, pad1, pad2;
} mmdgan_gemm_class;
typedef struct mmdgan_gemm_desc {
    const mmdgan_bf16* src;
    long long src_plane;
    int Nimg, Hs, Ws, Cs;
    int Hg, Wg, sy, sx, TH, TW;
    const mmdgan_bf16* w; /* packed weights */
    long long w_plane;
    long long w_rows; /* classes * rows_pad */
    int kpad, classes;
    int src_fmt, w_fmt; /* MMDGAN_FMT_* of src and w */
    void* dst;          /* out_mode 0: mmdgan_bf16 planes [dst_npl][rows][Cd]; out_mode 2: float [rows][Cd] */
    long long dst_plane;
    int dst_npl;        /* planes written in out_mode 0 (1..3) */
    int dst_fmt;        /* MMDGAN_FMT_* of the planes written in out_mode 0 */
    int Hd, Wd, Cd, osy, osx, Ncols;
    float alpha_k;
    const float* sigma; /* alpha = sigma ? alpha_k / *sigma : alpha_k */
    const float* bias;
    int act;            /* 0 linear, 1 lrelu(0.1), 2 relu, 3 tanh */
    const mmdgan_bf16* aux; /* planes of the layer OUTPUT a: the result is multiplied by act'(a) */
    long long aux_plane;
    int aux_npl, aux_fmt;
    int aux_mode;       /* 1 lrelu', 2 relu' (sign of plane 0), 3 tanh' = 1 - a^2 (all aux_npl planes) */
    long long aux_wrap_at, aux_wrap_len;
    float* colsum;      /* [tiles_m * classes][Ncols] or null */
    float* colsumsq;
    long long colsum_rows;
    int out_mode;       /* 0 bf16 planes, 2 raw fp32 */
    int bn;             /* N tile: 16, 32, 64, 128, 256 */
    int npass;          /* 6, 3 or 1 plane-pair products per k-block (see the header comment) */
    int* sat_flag;      /* optional device int: set to 1 when a value written as fp16 planes exceeds the format's range */
    int cta_pair;       /* 1: tcgen05 cta_group::2 -- a 2-CTA cluster shares one 256 x bn tile (bn 64, 128 or 256) */
    mmdgan_gemm_class cls[4];
} mmdgan_gemm_desc;
int mmdgan_gather_gemm(const mmdgan_gemm_desc* d, void* stream);
/* number of M tiles (rows of the colsum workspace per class) */
int mmdgan_gather_gemm_tiles(int Nimg, int Hg, int Wg);

/* Direct 3x3 / stride-1 / SAME convolution on the CUDA cores for layers with <= 4 channels on one side (the image layers:
 * tf.nn.conv2d at layer_func.py:912-916 for D's first and G's last layer, and their input gradients, my_sngan.py:301-304).
 * fp32 FMAs on the values reassembled from the planes (exact products: no plane-pair passes); weights are read from the
 * layer's canonical fp32 array [k][k][Cin][Cout] through strides, mirrored (flip = 1) for an input gradient.
 * Either Cout <= 4 and Cin % 16 == 0 (<= 128), or Cin <= 4 and Cout % 16 == 0 (<= 128).  colsum / aux: Cout <= 4 only. */
typedef struct mmdgan_direct_desc {
    const mmdgan_bf16* src;
    long long src_plane;
    int src_npl, Cs, src_fmt, dst_fmt, aux_fmt, pad0;
    int N, H, W;
    int Cin, Cout;
    const float* w;
    long long w_tap, w_in, w_out;
    int flip;
    void* dst;
    long long dst_plane;
    int dst_npl, Cd, out_mode;
    float alpha_k;
    const float* sigma;
    const float* bias;
    int act;
    const mmdgan_bf16* aux;
    long long aux_plane;
    int aux_npl, aux_mode;
    float* colsum; /* [mmdgan_direct_conv_blocks()][Cd] */
    int* sat_flag; /* optional, as in mmdgan_gemm_desc */
} mmdgan_direct_desc;
int mmdgan_direct_conv(const mmdgan_direct_desc* d, void* stream);
/* The many -> few direction of the same two layers (forward of C -> 3, input gradient of 3 -> C; tf.nn.conv2d at
 * layer_func.py:912-916 and its input gradient) as a dense [C -> 27] product over 27 = 9 taps x 3 channels on mmdgan_gather_gemm
 * followed by this kernel: from T (fp32 [pixels][32], column tap * 3 + c) it forms
 * y[p][c] = act(alpha * sum_tap T[p +- off(tap)][tap * 3 + c] + bias[c]) (* act'(aux)), writes planes (channels >= 3 zero) or raw
 * fp32 and, optionally, per-block column sums [mmdgan_tapsum_blocks][Cd].  flip = 1: minus (input gradient). */
int mmdgan_tapsum3x3_small(const float* T, int N, int H, int W, int flip, float alpha_k, const float* sigma, const float* bias, int act,
                           const mmdgan_bf16* aux, long long aux_plane, int aux_npl, int aux_fmt, int aux_mode, void* dst, long long dst_plane,
                           int dst_npl, int dst_fmt, int Cd, int out_mode, float* colsum, int* sat_flag, void* stream);
int mmdgan_tapsum_blocks(int N, int H, int W);
int mmdgan_direct_conv_blocks(int N, int H, int W);

/* out[m][n] = alpha * sum_k a[m][k] * wt[n][k] + bias[n] for N in {4,8,16,32} output columns (the critic's score layer,
 * tf.matmul at layer_func.py:909-911 with 16 outputs): fp32 CUDA-core kernel on the values reassembled from npl planes of
 * the activation a and of the packed forward operand wt.  Split over K slices of 1024: `workspace` (caller-owned, at least
 * mmdgan_dense_small_workspace(rows, K, N) bytes) receives the slice partials, which a second small launch sums in a fixed order. */
int mmdgan_dense_small_fwd(const mmdgan_bf16* a, long long a_plane, int npl, int a_fmt, int rows, int K, const mmdgan_bf16* wt,
                           long long w_plane, int w_fmt, int kpad, int N, float alpha_k, const float* sigma, const float* bias, float* out, int ldo,
                           float* workspace, void* stream);
size_t mmdgan_dense_small_workspace(int rows, int K, int N);

/* Weight gradient: W[r][(t,c)] = sum_p P[p][r] * G[g(p,t)][c] (filter gradients of the ops above and the
 * d(sigma)/dW term of SpectralNorm, GeneralTools/math_func.py:661-672).  out: [splits][Cp][TH*TW*Cs]. */
typedef struct mmdgan_wgrad_desc {
    const mmdgan_bf16* plain; /* [planes][P][Cp] */
    long long plain_plane;
    long long P;
    int Cp;
    const mmdgan_bf16* g;     /* gathered activation planes [Nimg*Hs*Ws][Cs] */
    long long g_plane;
    int Nimg, Hs, Ws, Cs;
    int Hg, Wg, sy, sx, TH, TW, oy, ox;
    int splits;
    float* out;
    int bn, npass;      /* bn 64, 128 or 256; npass 3 or 1 */
    int p_fmt, g_fmt;   /* MMDGAN_FMT_* of plain / g (both bf16, or both fp16); the partial tiles carry the product of the format scales */
} mmdgan_wgrad_desc;
int mmdgan_wgrad_gemm(const mmdgan_wgrad_desc* d, void* stream);

/* split-K partials -> canonical gradient; canon index = base + r*sr + t*st + c*sc for column t*Cg + c.
 * dots (optional, [mmdgan_wgrad_reduce_blocks()] doubles) receives per-block partial <G, W>. */
typedef struct mmdgan_wred_desc {
    const float* partials;
    float scale;  /* the summed partials are multiplied by this (1 / the format scales of the GEMM operands) */
    int splits, R, NC, Cg, Cvalid, Rvalid;
    int r_perm_C, r_perm_HW, c_perm_C, c_perm_HW; /* optional NHWC-flatten -> NCHW-flatten permutation of r / c */
    long long base, sr, st, sc;
    const float* w;
    float* out;
    double* dots;
} mmdgan_wred_desc;
int mmdgan_wgrad_reduce(const mmdgan_wred_desc* d, void* stream);
int mmdgan_wgrad_reduce_blocks(long long total);
/* the same for every layer of a net in ONE launch: jobs_device = device array of njobs mmdgan_wred_desc, block_start_device =
 * njobs + 1 ints (job j owns blocks [start[j], start[j+1]) = mmdgan_wgrad_reduce_blocks(R * NC) of them, which is also the length
 * of its `dots`), total_blocks = start[njobs] */
int mmdgan_wgrad_reduce_batched(const void* jobs_device, const int* block_start_device, int njobs, int total_blocks, void* stream);
/* mmdgan_sn_grad_combine for several layers in one launch; a job is {float* g; const float* s; const double* dots; const float* sigma;
 * long long n; int ndots; float act_k;} (40 bytes) in device memory */
typedef struct mmdgan_sn_combine_job {
    float* g;
    const float* s;
    const double* dots;
    const float* sigma;
    long long n;
    int ndots;
    float act_k;
} mmdgan_sn_combine_job;
int mmdgan_sn_grad_combine_batched(const void* jobs_device, int njobs, int blocks, void* stream);
/* grad = m*G - (m/sigma)*<G,W>*S with m = act_k/sigma: gradient through kernel * act_k / SpectralNorm(kernel) */
int mmdgan_sn_grad_combine(float* g, const float* s, const double* dots, int ndots, const float* sigma, float act_k, long long n,
                           void* stream);
int mmdgan_scale_by_sigma(float* g, const float* sigma, float act_k, long long n, void* stream);
/* sigma = ||v||, out = v / (sigma + eps) as planes: SpectralNorm._l2_norm / _l2_normalize_ (math_func.py:639-659) */
int mmdgan_sn_normalize(const float* v, long long n, float eps, float* sigma_out, mmdgan_bf16* out, long long out_plane, int npl,
                        int fmt, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Reductions of per-tile partial sums (bias gradients; deterministic order, double accumulation) */
int mmdgan_reduce_tiles(const float* partials, int T, int C, float scale, float* out, void* stream);
int mmdgan_colsum_small(const float* x, int rows, int C, float* out, void* stream);
int mmdgan_colsum_planes(const mmdgan_bf16* x, long long plane, int npl, int rows, int C, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Batch normalisation: tf.layers.batch_normalization(axis=1, training, fused=True) (layer_func.py:953-966) */
/* bessel != 0: the moving variance is fed the Bessel-corrected batch variance (TF's fused kernel: rank-4 inputs); 0: the biased
 * one (rank-2 inputs, for which TF 1.8 silently falls back from fused=True to nn.moments) */
int mmdgan_bn_finalize(const float* psum, const float* psq, int T, int C, long long rows, float eps, float momentum, float* mean,
                       float* invstd, float* moving_mean, float* moving_var, int bessel, void* stream);
/* Inference-mode statistics, tf.layers.batch_normalization(training=False) (reference layer_func.py:953-966 with
 * is_training False, the eval_sampling graph of my_sngan.py:523-533): mean = moving_mean, invstd = 1/sqrt(moving_var + eps);
 * mmdgan_bn_apply then normalises with them. */
int mmdgan_bn_inference_stats(const float* moving_mean, const float* moving_var, int C, float eps, float* mean, float* invstd,
                              void* stream);
int mmdgan_bn_apply(const float* z, const float* mean, const float* invstd, const float* gamma, const float* beta, int C,
                    long long total, int act, mmdgan_bf16* out, long long out_plane, int npl, int fmt, int* sat_flag, void* stream);
int mmdgan_bn_bwd_reduce(const float* da, const float* z, const float* mean, const float* invstd, const float* gamma,
                         const float* beta, int C, long long rows, int rows_per_block, int act, float* psum, float* psumx,
                         void* stream);
int mmdgan_bn_bwd_apply(const float* da, const float* z, const float* mean, const float* invstd, const float* gamma,
                        const float* beta, const float* dbeta, const float* dgamma, int C, long long rows, int act,
                        mmdgan_bf16* out, long long out_plane, int npl, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Fused pairwise squared distance -> Gaussian kernel(s) -> rep / rmb / mmd_g / mgb losses + score gradients
 * (GeneralTools/math_func.py:767-858, 1048-1069, 1288-1473, 2160-2193, 2505-2550) and the t-distribution kernel mixture
 * mmd_t (1087-1184, 2263-2275).  Row-block form for data
 * parallelism: local rows are rows [row0, row0+b) of the Bg global rows. */
typedef struct mmdgan_mmd_desc {
    const float* gen_loc;
    const float* real_loc;
    const float* gen_all;
    const float* real_all;
    int b, Bg, row0, d;
    int n_sigma;
    float sigma[8];
    float cD[3];
    int bmode[3];
    float bval[3];
    int family;    /* 0: Gaussian kernels, sigma[] = bandwidths; 1: t-distribution kernels (mmd_t), sigma[] = the alphas */
    float beta;    /* t-distribution kernels only (math_func.py:2110) */
    float* sums;   /* [6] */
    float* losses; /* [2] loss_gen, loss_dis */
    float* dLg_dgen;
    float* dLg_dreal; /* may be null */
    float* dLd_dgen;
    float* dLd_dreal;
    void* workspace; /* mmdgan_mmd_workspace(b) bytes, zeroed once before the first call */
} mmdgan_mmd_desc;
/* fills n_sigma / sigma / cD / bmode / bval / family / beta for loss_type in {"rep","rmb","mmd_g","mgb","mmd_t"} and rep_weights (w0, w1);
 * returns MMDGAN_EINVAL for unknown types or w0 - w1 != 1 (the reference's assert, math_func.py:1340) */
int mmdgan_mmd_configure(mmdgan_mmd_desc* d, const char* loss_type, float w0, float w1);
size_t mmdgan_mmd_workspace(int b);
/* N(0, 1) samples on the device: replaces tf.random_normal([batch_size, code_size]) of SNGan.sample_codes (reference
 * DeepLearning/my_sngan.py:122-124).  Philox-4x32-10 keyed by `seed`, counter = (quadruple index, *draw_counter), Box-Muller.
 * draw_counter (device, may be null = 0) is read by the kernel, so a captured graph draws fresh codes on every replay once
 * mmdgan_incr_counter follows it.  raw_words (may be null) receives the four Philox words per quadruple (tests). */
int mmdgan_sample_normal(float* out, long long n, unsigned long long seed, const unsigned long long* draw_counter, unsigned int* raw_words,
                         void* stream);
int mmdgan_incr_counter(unsigned long long* counter, void* stream);
/* data-parallel step: the two losses from the six GLOBAL kernel sums (after their all-reduce):
 * loss_gen = e_gg + e_rr - 2 e_gr (math_func.py:1342), loss_dis = cD0 e_gg^b + cD1 e_gr^b + cD2 e_rr^b (math_func.py:1421) */
int mmdgan_losses_from_sums(const float* sums, float cD0, float cD1, float cD2, float* losses, void* stream);
int mmdgan_mmd_fwd_bwd(const mmdgan_mmd_desc* d, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * tf.train.AdamOptimizer(lr, beta1, beta2, eps).apply_gradients over one flat parameter buffer
 * (GeneralTools/graph_func.py:518-527; DeepLearning/my_sngan.py:424-426).  *step holds t for this update. */
int mmdgan_adam(float* w, float* m, float* v, const float* g, long long n, float lr, float beta1, float beta2, float eps,
                const int* step, void* stream);
/* Forward exchange of the data-parallel MMD loss through the same multicast mechanism (replaces the all-gather of the score
 * blocks): s_local [2b, d] (rows [0, b) real, [b, 2b) generated, my_sngan.py:279) is stored to rows
 * [rank * b, (rank + 1) * b) of real_all / gen_all on EVERY rank; `*_mc` are multicast addresses of two [world * b, d] fp32
 * buffers of a symmetric allocation.  The caller brackets the call with a cross-rank barrier on the stream. */
int mmdgan_scatter_scores_nvls(const float* s_local, int b, int d, int rank, float* gen_all_mc, float* real_all_mc, void* stream);
/* out[i] = sum over ranks of in[i] for a few floats (the six kernel sums e_gg, e_gr, e_rr, ... of the row-block MMD form,
 * math_func.py:1048-1069 evaluated per rank): `in_mc` is the multicast address of a per-rank slot of a symmetric allocation,
 * n a multiple of 4.  The caller places a cross-rank barrier between writing the slot and this call. */
int mmdgan_allreduce_small_nvls(float* out, const float* in_mc, int n, void* stream);
int mmdgan_incr_step(int* step, void* stream);
/* Data-parallel form of the same update (new functionality, SURVEY.md 8e: the reference is single-GPU): gradient all-reduce
 * FUSED with Adam through NVSwitch multicast.  The caller keeps g, w, m, v of one network in a symmetric allocation that is
 * mapped on every rank and bound to a multicast object; `w`, `m`, `v` are THIS rank's replicas, `*_mc` the multicast addresses
 * of the four buffers.  For elements [begin, end) (this rank's shard; multiples of 4, 16-byte aligned) the kernel loads the
 * sum of all ranks' gradients (multimem.ld_reduce), applies the update above and stores the new w, m, v to every replica
 * (multimem.st).  The caller brackets the call with a cross-rank barrier on the stream: all gradients written before, all
 * parameters visible after. */
int mmdgan_adam_allreduce_nvls(const float* w, const float* m, const float* v, const float* g_mc, float* w_mc, float* m_mc,
                               float* v_mc, long long begin, long long end, float lr, float beta1, float beta2, float eps,
                               const int* step, void* stream);
/* device-side replacement of the per-step `assert not any(isnan(loss))` (GeneralTools/graph_func.py:856) */
int mmdgan_nan_flag(const float* x, int n, int* flag, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMDGAN_B200_H */
