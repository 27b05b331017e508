// Parameter blocks of the two tensor-core kernels (shared between the kernels and the C-ABI wrappers in api.cu).
#pragma once
#include <stdint.h>

namespace mg {

// K elements per pipeline stage of the gather-GEMM = width of a channel chunk in its K order (channel chunk, tap, channel):
// 64 16-bit elements = one 128-byte swizzle row.  The packed weights (pack_value, elementwise.cu) use the same constant.
static constexpr int kGemmBK = 64;

// One "class" of a gather-GEMM launch (blockIdx.z).  A k4/s2 transposed convolution is four classes, one per output
// parity; everything else is one class.
struct GemmClass {
    int oy, ox;    // source pixel = (y*sy + a + oy, x*sx + b + ox) for tap (a, b)
    int ooy, oox;  // destination pixel = (y*osy + ooy, x*osx + oox)
    int wrow;      // first row of this class in the weight matrix
    int pad0, pad1, pad2;
};

// D[m][n] = act(alpha * sum_k A[m][k] * Bw[n][k] + bias[n]) (* act'(aux[m][n])),  m = (img, y, x) on an Hg x Wg grid,
// k = (tap, channel); A is gathered from an NHWC activation, Bw is a K-major weight matrix read by TMA.
// Operands are bf16 planes (tc_common.cuh); the result is written as bf16 planes or as raw fp32.
struct ConvGemmParams {
    const uint16_t* src;   // [planes][Nimg*Hs*Ws][Cs], 16-bit planes in format src_fmt
    long long src_plane;   // elements between consecutive planes
    int src_fmt, w_fmt;    // FMT_* of the gathered operand and of the packed weights (tc_common.cuh)
    int Nimg, Hs, Ws, Cs;  // Cs in {8, 16} or a multiple of 32
    int Hg, Wg, sy, sx, TH, TW;
    int M;       // Nimg*Hg*Wg
    int ksteps;  // padded K / 32
    void* dst;   // out_mode 0: bf16 planes [dst_npl][Nimg*Hd*Wd][Cd]; out_mode 2: fp32 [Nimg*Hd*Wd][Cd]
    long long dst_plane;
    int dst_npl, dst_fmt;
    int Hd, Wd, Cd, osy, osx;
    int Ncols;               // valid output columns (multiple of 4)
    float alpha_k;           // alpha = sigma ? alpha_k / *sigma : alpha_k
    const float* sigma;
    const float* bias;       // [Ncols] or null
    int act;                 // 0 linear, 1 lrelu(0.1), 2 relu, 3 tanh
    const uint16_t* aux;     // bf16 planes [aux_npl][rows][Cd]: activation whose derivative multiplies the result, or null
    long long aux_plane;
    int aux_npl, aux_fmt;
    int aux_mode;            // 1 lrelu', 2 relu' (sign of plane 0), 3 tanh' (all planes; evaluated from the layer OUTPUT)
    long long aux_wrap_at;   // destination rows >= aux_wrap_at read aux at row - aux_wrap_len
    long long aux_wrap_len;
    float* colsum;           // [gridDim.x*gridDim.z][Ncols] per-tile column sums of the written values, or null
    float* colsumsq;         // same for squares, or null
    long long colsum_rows;   // only destination rows < colsum_rows are counted
    int out_mode;            // 0: bf16 planes, 2: raw fp32
    unsigned int* err;       // device error flag (watchdog)
    int* sat_flag;           // set to 1 when a value written as fp16 planes saturates (or null)
    int tiles_m, tiles_n, classes;   // filled by the launcher: tile grid (linear tile index = ((tile_m [/ 2]) * classes + class) * tiles_n + tile_n)
    long long* prof;                 // profiling experiments only (MMDGAN_PROF=1): per-CTA wait-cycle counters, else null
    unsigned int* sched;             // filled by the launcher: ticket counter of the dynamic tile scheduler (zero between launches)
    int debug;               // profiling experiments only: bit0 = skip the A gather, bit1 = skip the MMAs
    GemmClass cls[4];
};

// W[r][(t, c)] = sum_p P[p][r] * G[g(p, t)][c]:  P plain [pixels][Cp] (TMA, MN-major), G gathered NHWC activation.
// blockIdx.z = split of the pixel range; every split writes its own partial tile.
struct WgradParams {
    const uint16_t* g;     // gathered activation, 16-bit planes [planes][Nimg*Hs*Ws][Cs]
    long long g_plane;
    int p_fmt, g_fmt;      // FMT_* of the plain and of the gathered operand (may differ: fp16 activations x bf16 gradients)
    int Nimg, Hs, Ws, Cs;
    int Hg, Wg, sy, sx, TH, TW, oy, ox;   // pixel p = (img, y, x) on the plain operand's Hg x Wg grid
    long long P;           // number of plain pixels = Nimg*Hg*Wg
    long long p_per_split; // pixels per blockIdx.z (multiple of 32)
    int Cp;                // plain channels (rows of the result), multiple of 8
    int Ncols;             // TH*TW*Cs
    float* out;            // [splits][Cp][Ncols]
    unsigned int* err;
};


// direct_conv.cu: 3x3 / stride 1 / SAME convolution with <= 4 channels on one side (the image layers), NHWC
struct DirectConvParams {
    const uint16_t* src;     // planes [src_npl][N*H*W][Cs] in format src_fmt
    long long src_plane;
    int src_npl, Cs, src_fmt, dst_fmt, aux_fmt, pad0;
    int N, H, W;
    int Cin, Cout;           // real channel counts of THIS convolution (for an input gradient: Cout_layer -> Cin_layer)
    const float* w;          // canonical fp32 weights of the layer, element (tap, in, out) at tap*w_tap + in*w_in + out*w_out
    long long w_tap, w_in, w_out;
    int flip;                // 1: taps mirrored (input gradient)
    void* dst;               // out_mode 0: bf16 planes [dst_npl][N*H*W][Cd]; 2: fp32 [N*H*W][Cd]
    long long dst_plane;
    int dst_npl, Cd, out_mode;
    float alpha_k;
    const float* sigma;
    const float* bias;       // [Cout] (padded) or null
    int act;
    const uint16_t* aux;     // planes [aux_npl][N*H*W][Cd] or null (LS only)
    long long aux_plane;
    int aux_npl, aux_mode;
    float* colsum;           // [blocks][Cd] per-block column sums of the written values, or null (LS only)
    int* sat_flag;           // set to 1 when a value written as fp16 planes saturates (or null)
};

// mmd.cu
struct MmdParams {
    const float* gen_loc;   // [b][d]
    const float* real_loc;  // [b][d]
    const float* gen_all;   // [Bg][d]
    const float* real_all;  // [Bg][d]
    int b, Bg, row0, d;
    int n_sigma;
    int family;        // 0: Gaussian kernels exp(-d / (2 sigma^2)); 1: t-distribution kernels (1 + d / (alpha beta))^-alpha
    float c_s[8];      // Gaussian: 1 / (2 sigma^2); t: alpha
    float c_t[8];      // t: 1 / (alpha beta)
    float inv_beta;    // t: 1 / beta
    float cD[3];       // loss_dis = cD[0] e_gg^b + cD[1] e_gr^b + cD[2] e_rr^b
    int bmode[3];      // 0 none, 1 lower bound (max), 2 upper bound (min)   [gg, gr, rr]
    float bval[3];
    float* sums;       // [6]  e_gg, e_gr, e_rr, e_gg^b, e_gr^b, e_rr^b over the LOCAL rows (already times 1/(Bg(Bg-1)))
    float* losses;     // [2]  loss_gen, loss_dis from the local sums (the global losses when b == Bg)
    float* dLg_dgen;   // [b][d]
    float* dLg_dreal;  // [b][d] (may be null)
    float* dLd_dgen;   // [b][d]
    float* dLd_dreal;  // [b][d]
    float* partials;   // [gridDim.x][6] workspace
    unsigned int* counter;  // workspace, zero before the first launch; the kernel leaves it zero
};

// elementwise.cu
// Canonical (reference) layouts: conv [k][k][Cin][Cout], transposed conv [k][k][Cout][Cin], dense [in][out].
// Packed operand: [planes][classes*rows_pad][kpad], element (class, row, col) as documented per mode.
struct PackParams {
    const float* w;   // canonical
    uint16_t* out;    // packed planes
    long long plane;  // elements between planes
    int npl;          // planes to write
    int fmt;          // FMT_BF16 or FMT_F16W
    int mode;         // PACK_* enum
    int k;            // spatial kernel size
    int Cin, Cout;    // of the layer op (dense: in / out features)
    int Cs;           // channels per tap in the packed K index (>= source channels; 8, 16 or a multiple of 32)
    int rows_pad, kpad, classes;
    int in_C, in_HW, out_C, out_HW;  // dense only: NCHW-flatten <-> NHWC-flatten permutation of features
};
struct RefreshJob {
    int kind;   // 0: pack weights, 1: permute / pad a per-feature vector
    int block_start;   // first block of the flat grid that works on this job (ascending over the job table)
    PackParams pack;
    const float* src;
    float* dst;
    int n, C, HW, inverse;
};
// Split-K weight-gradient partials [splits][R][NC] -> canonical layout, canon index = base + r*sr + t*st + c*sc with
// column = t*Cg + c.  Also emits per-block partial <G, W> (for the spectral-norm term) when dots != null.
struct WredParams {
    const float* partials;
    float scale;                     // the summed partials are multiplied by this (removes the fp16 plane scale)
    int splits, R, NC, Cg, Cvalid;   // channels c < Cvalid are real (Cg may be padded)
    int Rvalid;                      // rows r < Rvalid are real (R may be padded)
    int r_perm_C, r_perm_HW, c_perm_C, c_perm_HW;  // optional NHWC-flatten -> NCHW-flatten feature permutation (HW <= 1: none)
    long long base, sr, st, sc;
    const float* w;       // canonical weights (for the dot) or null
    float* out;           // canonical gradient
    double* dots;         // [gridDim.x] or null
};

}  // namespace mg
