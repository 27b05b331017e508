// C ABI of libmmdgan_b200.so (declared in include/mmdgan_b200.h): argument validation + kernel launches.
#include "../../include/mmdgan_b200.h"
#include "conv_gemm.cuh"
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

namespace mg {
// conv_gemm.cu / wgrad_gemm.cu
int launch_conv_gemm(const ConvGemmParams& p, const uint16_t* w, long long w_plane, long long w_rows, int kpad, int classes, int bn,
                     int npass, int pair, cudaStream_t st);
int launch_wgrad_gemm(const WgradParams& p, const uint16_t* plain, long long plain_plane, int splits, int bn, int npass,
                      cudaStream_t st);
// direct_conv.cu
int launch_direct_conv(const DirectConvParams& p, cudaStream_t st);
int direct_conv_blocks(int N, int H, int W);
// mmd.cu
int launch_mmd(const MmdParams& p, cudaStream_t st);
int mmd_grid_blocks(int b);
// elementwise.cu
int l_nchw_to_nhwc(const float*, uint16_t*, long long, int, int, int, int, int, int, int, cudaStream_t);
int l_nhwc_to_nchw(const uint16_t*, long long, int, int, float*, int, int, int, int, int, cudaStream_t);
int l_to_planes(const float*, uint16_t*, long long, int, int, long long, cudaStream_t);
int l_from_planes(const uint16_t*, long long, int, int, float*, long long, cudaStream_t);
int l_colsum_planes(const uint16_t*, long long, int, int, int, float*, cudaStream_t);
int l_convert_planes(const uint16_t*, long long, int, int, uint16_t*, long long, int, int, long long, cudaStream_t);
int l_pack_weights(const PackParams&, cudaStream_t);
int l_permute_features(const float*, float*, int, int, int, int, cudaStream_t);
int l_reduce_tiles(const float*, int, int, float, float*, cudaStream_t);
int l_colsum_small(const float*, int, int, float*, cudaStream_t);
int wgrad_reduce_blocks(long long total);
int l_wgrad_reduce(const WredParams&, cudaStream_t);
int l_sn_grad_combine(float*, const float*, const double*, int, const float*, float, long long, cudaStream_t);
int l_scale_by_sigma(float*, const float*, float, long long, cudaStream_t);
int l_sn_normalize(const float*, long long, float, float*, uint16_t*, long long, int, int, cudaStream_t);
int l_bn_finalize(const float*, const float*, int, int, long long, float, float, float*, float*, float*, float*, int, cudaStream_t);
int l_bn_inference_stats(const float*, const float*, int, float, float*, float*, cudaStream_t);
int l_bn_apply(const float*, const float*, const float*, const float*, const float*, int, long long, int, uint16_t*, long long, int, int,
               int*, cudaStream_t);
int l_bn_bwd_reduce(const float*, const float*, const float*, const float*, const float*, const float*, int, long long, int, int,
                    float*, float*, cudaStream_t);
int l_bn_bwd_apply(const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, int,
                   long long, int, uint16_t*, long long, int, cudaStream_t);
int l_adam(float*, float*, float*, const float*, long long, float, float, float, float, const int*, cudaStream_t);
int l_adam_allreduce_nvls(const float*, const float*, const float*, const float*, float*, float*, float*, long long, long long, float, float, float,
                          float, const int*, cudaStream_t);
int l_scatter_scores_nvls(const float*, int, int, int, float*, float*, cudaStream_t);
int l_allreduce_small_nvls(float*, const float*, int, cudaStream_t);
int l_incr_step(int*, cudaStream_t);
int l_wgrad_reduce_batched(const WredParams*, const int*, int, int, cudaStream_t);
int l_sn_grad_combine_batched(const void*, int, int, cudaStream_t);
int tapsum_blocks(int N, int H, int W);
int launch_tapsum27(const float*, int, int, int, int, float, const float*, const float*, int, const uint16_t*, long long, int, int, int, void*, long long,
                    int, int, int, int, float*, int*, cudaStream_t);
int l_sample_normal(float*, long long, unsigned long long, const unsigned long long*, uint32_t*, cudaStream_t);
int l_incr_u64(unsigned long long*, cudaStream_t);
int l_losses_from_sums(const float*, float, float, float, float*, cudaStream_t);
int l_refresh(const RefreshJob*, int, long long, cudaStream_t);
long long dense_small_workspace(int rows, int K, int N);
int l_dense_small_fwd(const uint16_t*, long long, int, int, int, int, const uint16_t*, long long, int, int, int, float, const float*, const float*,
                      float*, int, float*, cudaStream_t);
int l_nan_flag(const float*, int, int*, cudaStream_t);
}  // namespace mg

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
static int wrap(int rc, const char* what) {
    if (rc == 0) return MMDGAN_OK;
    if (rc == -4) {
        cudaError_t e = cudaGetLastError();
        return fail(MMDGAN_ECUDA, "%s: CUDA error (%s)", what, cudaGetErrorString(e));
    }
    return fail(rc, "%s: unsupported configuration (rc %d)", what, rc);
}
static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" {

const char* mmdgan_last_error(void) { return g_err; }
int mmdgan_version(void) { return 200; }

static inline bool chan_ok(int c) { return c == 8 || c == 16 || c == 32 || (c > 0 && (c & 63) == 0); }
static inline int npl_for(int npass) { return npass == 6 ? 3 : (npass == 3 ? 2 : 1); }
static inline bool fmt_ok(int fmt, int npl) { return fmt == 0 ? (npl >= 1 && npl <= 3) : ((fmt == 1 || fmt == 2) && npl >= 1 && npl <= 2); }

int mmdgan_check_device(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail(MMDGAN_ECUDA, "mmdgan_check_device: no CUDA device");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return fail(MMDGAN_ECUDA, "mmdgan_check_device: cannot query device");
    if (prop.major != 10) return fail(MMDGAN_EARCH, "mmdgan_check_device: sm_%d%d is not sm_100 (B200)", prop.major, prop.minor);
    return MMDGAN_OK;
}

int mmdgan_nchw_to_nhwc(const float* src, mmdgan_bf16* dst, long long dst_plane, int npl, int fmt, int N, int C, int H, int W, int Cpad,
                        void* stream) {
    if (!src || !dst) return fail(MMDGAN_EINVAL, "mmdgan_nchw_to_nhwc: null pointer");
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || Cpad < C || (Cpad & 3) || !fmt_ok(fmt, npl) || (npl > 1 && dst_plane <= 0))
        return fail(MMDGAN_ESHAPE, "mmdgan_nchw_to_nhwc: bad shape");
    return wrap(mg::l_nchw_to_nhwc(src, dst, dst_plane, npl, fmt, N, C, H, W, Cpad, S(stream)), "mmdgan_nchw_to_nhwc");
}
int mmdgan_nhwc_to_nchw(const mmdgan_bf16* src, long long src_plane, int npl, int fmt, float* dst, int N, int C, int H, int W, int Cpad,
                        void* stream) {
    if (!src || !dst) return fail(MMDGAN_EINVAL, "mmdgan_nhwc_to_nchw: null pointer");
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || Cpad < C || !fmt_ok(fmt, npl) || (npl > 1 && src_plane <= 0))
        return fail(MMDGAN_ESHAPE, "mmdgan_nhwc_to_nchw: bad shape");
    return wrap(mg::l_nhwc_to_nchw(src, src_plane, npl, fmt, dst, N, C, H, W, Cpad, S(stream)), "mmdgan_nhwc_to_nchw");
}
int mmdgan_to_planes(const float* x, mmdgan_bf16* dst, long long dst_plane, int npl, int fmt, long long n, void* stream) {
    if (!x || !dst) return fail(MMDGAN_EINVAL, "mmdgan_to_planes: null pointer");
    if (!fmt_ok(fmt, npl) || (npl > 1 && dst_plane < n)) return fail(MMDGAN_ESHAPE, "mmdgan_to_planes: bad plane layout");
    if (n <= 0) return MMDGAN_OK;
    return wrap(mg::l_to_planes(x, dst, dst_plane, npl, fmt, n, S(stream)), "mmdgan_to_planes");
}
int mmdgan_from_planes(const mmdgan_bf16* src, long long src_plane, int npl, int fmt, float* out, long long n, void* stream) {
    if (!src || !out) return fail(MMDGAN_EINVAL, "mmdgan_from_planes: null pointer");
    if (!fmt_ok(fmt, npl) || (npl > 1 && src_plane < n)) return fail(MMDGAN_ESHAPE, "mmdgan_from_planes: bad plane layout");
    if (n <= 0) return MMDGAN_OK;
    return wrap(mg::l_from_planes(src, src_plane, npl, fmt, out, n, S(stream)), "mmdgan_from_planes");
}

int mmdgan_convert_planes(const mmdgan_bf16* src, long long src_plane, int src_npl, int src_fmt, mmdgan_bf16* dst, long long dst_plane,
                          int dst_npl, int dst_fmt, long long n, void* stream) {
    if (!src || !dst) return fail(MMDGAN_EINVAL, "mmdgan_convert_planes: null pointer");
    if (!fmt_ok(src_fmt, src_npl) || !fmt_ok(dst_fmt, dst_npl) || (src_npl > 1 && src_plane < n) || (dst_npl > 1 && dst_plane < n) || (n & 3) ||
        !al16(src) || !al16(dst) || (src_plane & 3) || (dst_plane & 3))
        return fail(MMDGAN_ESHAPE, "mmdgan_convert_planes: bad plane layout");
    if (n <= 0) return MMDGAN_OK;
    return wrap(mg::l_convert_planes(src, src_plane, src_npl, src_fmt, dst, dst_plane, dst_npl, dst_fmt, n, S(stream)), "mmdgan_convert_planes");
}

int mmdgan_pack_weights(const mmdgan_pack_desc* d, void* stream) {
    if (!d || !d->w || !d->out) return fail(MMDGAN_EINVAL, "mmdgan_pack_weights: null pointer");
    if (d->mode < 0 || d->mode > 6) return fail(MMDGAN_EINVAL, "mmdgan_pack_weights: unknown mode %d", d->mode);
    if ((d->fmt != 0 && d->fmt != 2) || !fmt_ok(d->fmt, d->npl) || (d->npl > 1 && d->plane <= 0)) return fail(MMDGAN_ESHAPE, "mmdgan_pack_weights: bad plane layout");
    if (d->rows_pad <= 0 || d->kpad <= 0 || (d->kpad & 63) || d->classes < 1 || d->classes > 4 || !chan_ok(d->Cs))
        return fail(MMDGAN_ESHAPE, "mmdgan_pack_weights: bad padded shape rows_pad=%d kpad=%d classes=%d Cs=%d", d->rows_pad, d->kpad,
                    d->classes, d->Cs);
    mg::PackParams p;
    p.w = d->w; p.out = d->out; p.plane = d->plane; p.npl = d->npl; p.fmt = d->fmt; p.mode = d->mode; p.k = d->k; p.Cin = d->Cin; p.Cout = d->Cout; p.Cs = d->Cs;
    p.rows_pad = d->rows_pad; p.kpad = d->kpad; p.classes = d->classes;
    p.in_C = d->in_C; p.in_HW = d->in_HW; p.out_C = d->out_C; p.out_HW = d->out_HW;
    return wrap(mg::l_pack_weights(p, S(stream)), "mmdgan_pack_weights");
}
int mmdgan_permute_features(const float* src, float* dst, int n, int C, int HW, int inverse, void* stream) {
    if (!src || !dst) return fail(MMDGAN_EINVAL, "mmdgan_permute_features: null pointer");
    if (n <= 0) return MMDGAN_OK;
    return wrap(mg::l_permute_features(src, dst, n, C, HW, inverse, S(stream)), "mmdgan_permute_features");
}

int mmdgan_gather_gemm_tiles(int Nimg, int Hg, int Wg) {
    const long long m = static_cast<long long>(Nimg) * Hg * Wg;
    return static_cast<int>(((m + 127) / 128 + 1) / 2 * 2);   /* even: the CTA-pair kernel launches whole pairs */
}

int mmdgan_gather_gemm(const mmdgan_gemm_desc* d, void* stream) {
    if (!d || !d->src || !d->w || !d->dst) return fail(MMDGAN_EINVAL, "mmdgan_gather_gemm: null pointer");
    if (d->npass != 1 && d->npass != 3 && d->npass != 6) return fail(MMDGAN_EINVAL, "mmdgan_gather_gemm: npass must be 1, 3 or 6");
    if (d->bn != 16 && d->bn != 32 && d->bn != 64 && d->bn != 128 && d->bn != 256)
        return fail(MMDGAN_EINVAL, "mmdgan_gather_gemm: bn must be 16/32/64/128/256");
    if (d->Nimg <= 0 || d->Hs <= 0 || d->Ws <= 0 || !chan_ok(d->Cs) || d->Hg <= 0 || d->Wg <= 0 || d->TH <= 0 || d->TW <= 0)
        return fail(MMDGAN_ESHAPE, "mmdgan_gather_gemm: bad source shape");
    if (d->kpad <= 0 || (d->kpad & 63) || d->kpad < d->TH * d->TW * d->Cs)
        return fail(MMDGAN_ESHAPE, "mmdgan_gather_gemm: kpad %d does not cover %d taps x %d channels", d->kpad, d->TH * d->TW, d->Cs);
    if (d->classes < 1 || d->classes > 4 || d->w_rows <= 0) return fail(MMDGAN_ESHAPE, "mmdgan_gather_gemm: bad class count");
    if (d->Ncols <= 0 || (d->Ncols & 3) || d->Cd < d->Ncols || (d->Cd & 3)) return fail(MMDGAN_ESHAPE, "mmdgan_gather_gemm: bad output columns");
    if (!al16(d->src) || !al16(d->dst) || !al16(d->w) || (d->src_plane & 7) || (d->dst_plane & 7) || (d->w_plane & 7))
        return fail(MMDGAN_ESHAPE, "mmdgan_gather_gemm: pointers / plane offsets must be 16-byte aligned");
    if (d->src_fmt < 0 || d->src_fmt > 2 || d->w_fmt < 0 || d->w_fmt > 2 || (d->npass == 6 && (d->src_fmt || d->w_fmt)))
        return fail(MMDGAN_EINVAL, "mmdgan_gather_gemm: npass 6 is the bf16 three-plane mode; fp16 plane operands use npass 3 or 1");
    if ((d->src_fmt == 0) != (d->w_fmt == 0)) return fail(MMDGAN_EINVAL, "mmdgan_gather_gemm: both operands must be bf16 planes or both fp16 planes");
    if (d->npass > 1 && (d->src_plane <= 0 || d->w_plane <= 0)) return fail(MMDGAN_ESHAPE, "mmdgan_gather_gemm: npass %d needs %d operand planes", d->npass, npl_for(d->npass));
    if (d->out_mode != 0 && d->out_mode != 2) return fail(MMDGAN_EINVAL, "mmdgan_gather_gemm: bad out_mode");
    if (d->out_mode == 0 && (!fmt_ok(d->dst_fmt, d->dst_npl) || (d->dst_npl > 1 && d->dst_plane <= 0)))
        return fail(MMDGAN_ESHAPE, "mmdgan_gather_gemm: bad destination plane layout");
    if (d->aux && (!fmt_ok(d->aux_fmt, d->aux_npl) || (d->aux_npl > 1 && d->aux_plane <= 0) || !al16(d->aux)))
        return fail(MMDGAN_ESHAPE, "mmdgan_gather_gemm: bad aux plane layout");
    if (d->cta_pair && d->bn != 64 && d->bn != 128 && d->bn != 256) return fail(MMDGAN_EINVAL, "mmdgan_gather_gemm: cta_pair needs bn 64, 128 or 256");
    if (d->npass == 6 && d->bn == 256 && !d->cta_pair) return fail(MMDGAN_EINVAL, "mmdgan_gather_gemm: npass 6 with bn 256 needs cta_pair");
    const long long M = static_cast<long long>(d->Nimg) * d->Hg * d->Wg;
    if (M > 2000000000ll) return fail(MMDGAN_ESHAPE, "mmdgan_gather_gemm: too many rows");
    mg::ConvGemmParams p;
    memset(&p, 0, sizeof(p));
    p.src = d->src; p.src_plane = d->src_plane; p.src_fmt = d->src_fmt; p.w_fmt = d->w_fmt; p.Nimg = d->Nimg; p.Hs = d->Hs; p.Ws = d->Ws; p.Cs = d->Cs;
    p.Hg = d->Hg; p.Wg = d->Wg; p.sy = d->sy; p.sx = d->sx; p.TH = d->TH; p.TW = d->TW;
    p.M = static_cast<int>(M); p.ksteps = d->kpad / mg::kGemmBK;
    p.dst = d->dst; p.dst_plane = d->dst_plane; p.dst_npl = d->dst_npl; p.dst_fmt = d->dst_fmt; p.Hd = d->Hd; p.Wd = d->Wd; p.Cd = d->Cd; p.osy = d->osy; p.osx = d->osx;
    p.Ncols = d->Ncols; p.alpha_k = d->alpha_k; p.sigma = d->sigma; p.bias = d->bias; p.act = d->act;
    p.aux = d->aux; p.aux_plane = d->aux_plane; p.aux_npl = d->aux_npl; p.aux_fmt = d->aux_fmt; p.aux_mode = d->aux_mode;
    p.aux_wrap_at = d->aux_wrap_at > 0 ? d->aux_wrap_at : (1ll << 62); p.aux_wrap_len = d->aux_wrap_len;
    p.colsum = d->colsum; p.colsumsq = d->colsumsq; p.colsum_rows = d->colsum_rows > 0 ? d->colsum_rows : (1ll << 62);
    p.out_mode = d->out_mode; p.err = nullptr; p.sat_flag = d->sat_flag;
    {
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("MMDGAN_DEBUG"); dbg = e ? atoi(e) : 0; }
        p.debug = dbg;
    }
    for (int i = 0; i < 4; ++i) {
        p.cls[i].oy = d->cls[i].oy; p.cls[i].ox = d->cls[i].ox; p.cls[i].ooy = d->cls[i].ooy; p.cls[i].oox = d->cls[i].oox;
        p.cls[i].wrow = d->cls[i].wrow;
    }
    return wrap(mg::launch_conv_gemm(p, d->w, d->w_plane, d->w_rows, d->kpad, d->classes, d->bn, d->npass, d->cta_pair, S(stream)), "mmdgan_gather_gemm");
}

int mmdgan_tapsum_blocks(int N, int H, int W) { return mg::tapsum_blocks(N, H, W); }
int mmdgan_tapsum3x3_small(const float* T, int N, int H, int W, int flip, float alpha_k, const float* sigma, const float* bias, int act,
                           const mmdgan_bf16* aux, long long aux_plane, int aux_npl, int aux_fmt, int aux_mode, void* dst, long long dst_plane,
                           int dst_npl, int dst_fmt, int Cd, int out_mode, float* colsum, int* sat_flag, void* stream) {
    if (!T || !dst) return fail(MMDGAN_EINVAL, "mmdgan_tapsum3x3_small: null pointer");
    if (N <= 0 || H <= 0 || W <= 0 || Cd < 4 || (Cd & 3)) return fail(MMDGAN_ESHAPE, "mmdgan_tapsum3x3_small: bad shape");
    if (out_mode != 0 && out_mode != 2) return fail(MMDGAN_EINVAL, "mmdgan_tapsum3x3_small: bad out_mode");
    if (out_mode == 0 && (!fmt_ok(dst_fmt, dst_npl) || (dst_npl > 1 && dst_plane <= 0))) return fail(MMDGAN_ESHAPE, "mmdgan_tapsum3x3_small: bad destination plane layout");
    if (aux && (!fmt_ok(aux_fmt, aux_npl) || (aux_npl > 1 && aux_plane <= 0))) return fail(MMDGAN_ESHAPE, "mmdgan_tapsum3x3_small: bad aux plane layout");
    return wrap(mg::launch_tapsum27(T, N, H, W, flip, alpha_k, sigma, bias, act, aux, aux_plane, aux_npl, aux_fmt, aux_mode, dst, dst_plane, dst_npl,
                                    dst_fmt, Cd, out_mode, colsum, sat_flag, S(stream)), "mmdgan_tapsum3x3_small");
}
int mmdgan_direct_conv_blocks(int N, int H, int W) { return mg::direct_conv_blocks(N, H, W); }
int mmdgan_direct_conv(const mmdgan_direct_desc* d, void* stream) {
    if (!d || !d->src || !d->w || !d->dst) return fail(MMDGAN_EINVAL, "mmdgan_direct_conv: null pointer");
    if (d->N <= 0 || d->H <= 0 || d->W <= 0 || d->Cs <= 0 || (d->Cs & 7) || d->Cd <= 0 || (d->Cd & 3))
        return fail(MMDGAN_ESHAPE, "mmdgan_direct_conv: bad shape");
    const bool ls = d->Cout >= 1 && d->Cout <= 4 && d->Cin > 0 && d->Cin % 16 == 0 && d->Cin <= 128 && d->Cs >= d->Cin && d->Cd >= 4;
    const bool sl = d->Cin >= 1 && d->Cin <= 4 && d->Cout > 0 && d->Cout % 16 == 0 && d->Cout <= 128 && d->Cd >= d->Cout;
    if (!ls && !sl) return fail(MMDGAN_ESHAPE, "mmdgan_direct_conv: needs <= 4 channels on one side and a multiple of 16 (<= 128) on the other");
    if (sl && (d->aux || d->colsum)) return fail(MMDGAN_EINVAL, "mmdgan_direct_conv: aux / colsum need Cout <= 4");
    if (!fmt_ok(d->src_fmt, d->src_npl) || (d->src_npl > 1 && d->src_plane <= 0) || (d->src_plane & 7) || !al16(d->src) || !al16(d->dst))
        return fail(MMDGAN_ESHAPE, "mmdgan_direct_conv: bad source plane layout");
    if (d->out_mode != 0 && d->out_mode != 2) return fail(MMDGAN_EINVAL, "mmdgan_direct_conv: bad out_mode");
    if (d->out_mode == 0 && (!fmt_ok(d->dst_fmt, d->dst_npl) || (d->dst_npl > 1 && d->dst_plane <= 0)))
        return fail(MMDGAN_ESHAPE, "mmdgan_direct_conv: bad destination plane layout");
    if (d->aux && (!fmt_ok(d->aux_fmt, d->aux_npl) || (d->aux_npl > 1 && d->aux_plane <= 0)))
        return fail(MMDGAN_ESHAPE, "mmdgan_direct_conv: bad aux plane layout");
    mg::DirectConvParams p;
    memset(&p, 0, sizeof(p));
    p.src = d->src; p.src_plane = d->src_plane; p.src_npl = d->src_npl; p.Cs = d->Cs; p.src_fmt = d->src_fmt; p.dst_fmt = d->dst_fmt;
    p.aux_fmt = d->aux_fmt; p.N = d->N; p.H = d->H; p.W = d->W;
    p.Cin = d->Cin; p.Cout = d->Cout; p.w = d->w; p.w_tap = d->w_tap; p.w_in = d->w_in; p.w_out = d->w_out; p.flip = d->flip;
    p.dst = d->dst; p.dst_plane = d->dst_plane; p.dst_npl = d->dst_npl; p.Cd = d->Cd; p.out_mode = d->out_mode;
    p.alpha_k = d->alpha_k; p.sigma = d->sigma; p.bias = d->bias; p.act = d->act;
    p.aux = d->aux; p.aux_plane = d->aux_plane; p.aux_npl = d->aux_npl; p.aux_mode = d->aux_mode; p.colsum = d->colsum; p.sat_flag = d->sat_flag;
    return wrap(mg::launch_direct_conv(p, S(stream)), "mmdgan_direct_conv");
}

int mmdgan_wgrad_gemm(const mmdgan_wgrad_desc* d, void* stream) {
    if (!d || !d->plain || !d->g || !d->out) return fail(MMDGAN_EINVAL, "mmdgan_wgrad_gemm: null pointer");
    if (d->npass != 1 && d->npass != 3) return fail(MMDGAN_EINVAL, "mmdgan_wgrad_gemm: npass must be 1 or 3");
    if (d->bn != 64 && d->bn != 128 && d->bn != 256) return fail(MMDGAN_EINVAL, "mmdgan_wgrad_gemm: bn must be 64/128/256");
    if (d->P <= 0 || d->Cp <= 0 || (d->Cp & 7) || d->Cs <= 0 || (d->Cs & 7) || d->splits <= 0)
        return fail(MMDGAN_ESHAPE, "mmdgan_wgrad_gemm: bad shape");
    if (d->P != static_cast<long long>(d->Nimg) * d->Hg * d->Wg) return fail(MMDGAN_ESHAPE, "mmdgan_wgrad_gemm: P != Nimg*Hg*Wg");
    if (!al16(d->plain) || !al16(d->g) || !al16(d->out) || (d->plain_plane & 7) || (d->g_plane & 7))
        return fail(MMDGAN_ESHAPE, "mmdgan_wgrad_gemm: pointers / plane offsets must be 16-byte aligned");
    if (d->p_fmt < 0 || d->p_fmt > 2 || d->g_fmt < 0 || d->g_fmt > 2 || ((d->p_fmt == 0) != (d->g_fmt == 0)))
        return fail(MMDGAN_EINVAL, "mmdgan_wgrad_gemm: both operands must be bf16 planes or both fp16 planes");
    if (d->npass == 3 && (d->plain_plane <= 0 || d->g_plane <= 0)) return fail(MMDGAN_ESHAPE, "mmdgan_wgrad_gemm: npass 3 needs two operand planes");
    mg::WgradParams p;
    memset(&p, 0, sizeof(p));
    p.g = d->g; p.g_plane = d->g_plane; p.p_fmt = d->p_fmt; p.g_fmt = d->g_fmt; p.Nimg = d->Nimg; p.Hs = d->Hs; p.Ws = d->Ws; p.Cs = d->Cs;
    p.Hg = d->Hg; p.Wg = d->Wg; p.sy = d->sy; p.sx = d->sx; p.TH = d->TH; p.TW = d->TW; p.oy = d->oy; p.ox = d->ox;
    p.P = d->P;
    long long per = (d->P + d->splits - 1) / d->splits;
    per = (per + 31) / 32 * 32;
    p.p_per_split = per;
    p.Cp = d->Cp; p.Ncols = d->TH * d->TW * d->Cs; p.out = d->out; p.err = nullptr;
    return wrap(mg::launch_wgrad_gemm(p, d->plain, d->plain_plane, d->splits, d->bn, d->npass, S(stream)), "mmdgan_wgrad_gemm");
}

int mmdgan_wgrad_reduce_blocks(long long total) { return mg::wgrad_reduce_blocks(total); }
int mmdgan_wgrad_reduce_batched(const void* jobs_device, const int* block_start_device, int njobs, int total_blocks, void* stream) {
    if (!jobs_device || !block_start_device) return fail(MMDGAN_EINVAL, "mmdgan_wgrad_reduce_batched: null pointer");
    if (njobs <= 0 || total_blocks <= 0) return fail(MMDGAN_ESHAPE, "mmdgan_wgrad_reduce_batched: empty job table");
    static_assert(sizeof(mmdgan_wred_desc) == sizeof(mg::WredParams), "job layout");
    return wrap(mg::l_wgrad_reduce_batched(static_cast<const mg::WredParams*>(jobs_device), block_start_device, njobs, total_blocks, S(stream)),
                "mmdgan_wgrad_reduce_batched");
}
int mmdgan_sn_grad_combine_batched(const void* jobs_device, int njobs, int blocks, void* stream) {
    if (!jobs_device) return fail(MMDGAN_EINVAL, "mmdgan_sn_grad_combine_batched: null pointer");
    if (njobs <= 0 || blocks <= 0 || njobs > 65535) return fail(MMDGAN_ESHAPE, "mmdgan_sn_grad_combine_batched: bad job table");
    return wrap(mg::l_sn_grad_combine_batched(jobs_device, njobs, blocks, S(stream)), "mmdgan_sn_grad_combine_batched");
}
int mmdgan_wgrad_reduce(const mmdgan_wred_desc* d, void* stream) {
    if (!d || !d->partials || !d->out) return fail(MMDGAN_EINVAL, "mmdgan_wgrad_reduce: null pointer");
    if (d->splits <= 0 || d->R <= 0 || d->NC <= 0 || d->Cg <= 0 || d->NC % d->Cg) return fail(MMDGAN_ESHAPE, "mmdgan_wgrad_reduce: bad shape");
    mg::WredParams p;
    p.partials = d->partials; p.scale = d->scale != 0.f ? d->scale : 1.f; p.splits = d->splits; p.R = d->R; p.NC = d->NC; p.Cg = d->Cg; p.Cvalid = d->Cvalid;
    p.Rvalid = d->Rvalid; p.r_perm_C = d->r_perm_C; p.r_perm_HW = d->r_perm_HW; p.c_perm_C = d->c_perm_C; p.c_perm_HW = d->c_perm_HW;
    p.base = d->base; p.sr = d->sr; p.st = d->st; p.sc = d->sc; p.w = d->w; p.out = d->out; p.dots = d->dots;
    return wrap(mg::l_wgrad_reduce(p, S(stream)), "mmdgan_wgrad_reduce");
}
int mmdgan_sn_grad_combine(float* g, const float* s, const double* dots, int ndots, const float* sigma, float act_k, long long n,
                           void* stream) {
    if (!g || !s || !dots || !sigma) return fail(MMDGAN_EINVAL, "mmdgan_sn_grad_combine: null pointer");
    return wrap(mg::l_sn_grad_combine(g, s, dots, ndots, sigma, act_k, n, S(stream)), "mmdgan_sn_grad_combine");
}
int mmdgan_scale_by_sigma(float* g, const float* sigma, float act_k, long long n, void* stream) {
    if (!g || !sigma) return fail(MMDGAN_EINVAL, "mmdgan_scale_by_sigma: null pointer");
    return wrap(mg::l_scale_by_sigma(g, sigma, act_k, n, S(stream)), "mmdgan_scale_by_sigma");
}
int mmdgan_sn_normalize(const float* v, long long n, float eps, float* sigma_out, mmdgan_bf16* out, long long out_plane, int npl,
                        int fmt, void* stream) {
    if (!v || !out) return fail(MMDGAN_EINVAL, "mmdgan_sn_normalize: null pointer");
    if (n <= 0 || !fmt_ok(fmt, npl) || (npl > 1 && out_plane < n)) return fail(MMDGAN_ESHAPE, "mmdgan_sn_normalize: bad shape");
    return wrap(mg::l_sn_normalize(v, n, eps, sigma_out, out, out_plane, npl, fmt, S(stream)), "mmdgan_sn_normalize");
}
int mmdgan_reduce_tiles(const float* partials, int T, int C, float scale, float* out, void* stream) {
    if (!partials || !out) return fail(MMDGAN_EINVAL, "mmdgan_reduce_tiles: null pointer");
    if (T <= 0 || C <= 0) return fail(MMDGAN_ESHAPE, "mmdgan_reduce_tiles: bad shape");
    return wrap(mg::l_reduce_tiles(partials, T, C, scale, out, S(stream)), "mmdgan_reduce_tiles");
}
int mmdgan_colsum_small(const float* x, int rows, int C, float* out, void* stream) {
    if (!x || !out) return fail(MMDGAN_EINVAL, "mmdgan_colsum_small: null pointer");
    return wrap(mg::l_colsum_small(x, rows, C, out, S(stream)), "mmdgan_colsum_small");
}
int mmdgan_colsum_planes(const mmdgan_bf16* x, long long plane, int npl, int rows, int C, float* out, void* stream) {
    if (!x || !out) return fail(MMDGAN_EINVAL, "mmdgan_colsum_planes: null pointer");
    if (rows <= 0 || C <= 0 || npl < 1 || npl > 3 || (npl > 1 && plane <= 0)) return fail(MMDGAN_ESHAPE, "mmdgan_colsum_planes: bad shape");
    return wrap(mg::l_colsum_planes(x, plane, npl, rows, C, out, S(stream)), "mmdgan_colsum_planes");
}
int mmdgan_bn_finalize(const float* psum, const float* psq, int T, int C, long long rows, float eps, float momentum, float* mean,
                       float* invstd, float* moving_mean, float* moving_var, int bessel, void* stream) {
    if (!psum || !psq || !mean || !invstd) return fail(MMDGAN_EINVAL, "mmdgan_bn_finalize: null pointer");
    if (T <= 0 || C <= 0 || rows <= 0) return fail(MMDGAN_ESHAPE, "mmdgan_bn_finalize: bad shape");
    return wrap(mg::l_bn_finalize(psum, psq, T, C, rows, eps, momentum, mean, invstd, moving_mean, moving_var, bessel, S(stream)), "mmdgan_bn_finalize");
}
int mmdgan_bn_inference_stats(const float* moving_mean, const float* moving_var, int C, float eps, float* mean, float* invstd,
                              void* stream) {
    if (!moving_mean || !moving_var || !mean || !invstd) return fail(MMDGAN_EINVAL, "mmdgan_bn_inference_stats: null pointer");
    if (C <= 0 || !(eps >= 0.0f)) return fail(MMDGAN_ESHAPE, "mmdgan_bn_inference_stats: bad shape");
    return wrap(mg::l_bn_inference_stats(moving_mean, moving_var, C, eps, mean, invstd, S(stream)), "mmdgan_bn_inference_stats");
}
int mmdgan_bn_apply(const float* z, const float* mean, const float* invstd, const float* gamma, const float* beta, int C,
                    long long total, int act, mmdgan_bf16* out, long long out_plane, int npl, int fmt, int* sat_flag, void* stream) {
    if (!z || !mean || !invstd || !gamma || !beta || !out) return fail(MMDGAN_EINVAL, "mmdgan_bn_apply: null pointer");
    if (C <= 0 || (C & 3) || total <= 0 || total % C || !fmt_ok(fmt, npl) || (npl > 1 && out_plane < total))
        return fail(MMDGAN_ESHAPE, "mmdgan_bn_apply: bad shape");
    return wrap(mg::l_bn_apply(z, mean, invstd, gamma, beta, C, total, act, out, out_plane, npl, fmt, sat_flag, S(stream)), "mmdgan_bn_apply");
}
int mmdgan_bn_bwd_reduce(const float* da, const float* z, const float* mean, const float* invstd, const float* gamma,
                         const float* beta, int C, long long rows, int rows_per_block, int act, float* psum, float* psumx,
                         void* stream) {
    if (!da || !z || !mean || !invstd || !gamma || !beta || !psum || !psumx) return fail(MMDGAN_EINVAL, "mmdgan_bn_bwd_reduce: null pointer");
    if (C <= 0 || rows <= 0 || rows_per_block <= 0) return fail(MMDGAN_ESHAPE, "mmdgan_bn_bwd_reduce: bad shape");
    return wrap(mg::l_bn_bwd_reduce(da, z, mean, invstd, gamma, beta, C, rows, rows_per_block, act, psum, psumx, S(stream)), "mmdgan_bn_bwd_reduce");
}
int mmdgan_bn_bwd_apply(const float* da, const float* z, const float* mean, const float* invstd, const float* gamma,
                        const float* beta, const float* dbeta, const float* dgamma, int C, long long rows, int act,
                        mmdgan_bf16* out, long long out_plane, int npl, void* stream) {
    if (!da || !z || !mean || !invstd || !gamma || !beta || !dbeta || !dgamma || !out) return fail(MMDGAN_EINVAL, "mmdgan_bn_bwd_apply: null pointer");
    if (C <= 0 || rows <= 0 || npl < 1 || npl > 3 || (npl > 1 && out_plane < rows * C)) return fail(MMDGAN_ESHAPE, "mmdgan_bn_bwd_apply: bad shape");
    return wrap(mg::l_bn_bwd_apply(da, z, mean, invstd, gamma, beta, dbeta, dgamma, C, rows, act, out, out_plane, npl, S(stream)), "mmdgan_bn_bwd_apply");
}

int mmdgan_mmd_configure(mmdgan_mmd_desc* d, const char* loss_type, float w0, float w1) {
    if (!d || !loss_type) return fail(MMDGAN_EINVAL, "mmdgan_mmd_configure: null pointer");
    for (int i = 0; i < 8; ++i) d->sigma[i] = 0.f;
    for (int i = 0; i < 3; ++i) { d->bmode[i] = 0; d->bval[i] = 0.f; }
    d->n_sigma = 1;
    d->sigma[0] = 1.0f;
    d->family = 0;
    d->beta = 2.0f;
    const bool rep = !strcmp(loss_type, "rep") || !strcmp(loss_type, "rep_mmd_g");
    const bool rmb = !strcmp(loss_type, "rmb") || !strcmp(loss_type, "rep_b") || !strcmp(loss_type, "rep_mmd_b");
    if (rep || rmb) {
        if (w0 - w1 != 1.0f) return fail(MMDGAN_EINVAL, "w[0]-w[1] must be 1");  /* math_func.py:1340 */
        /* loss_dis = w0 e_gr - e_gg - w1 e_rr  (math_func.py:1342, 1421) */
        d->cD[0] = -1.0f; d->cD[1] = w0; d->cD[2] = -w1;
        if (rmb) {
            d->bmode[0] = 1; d->bval[0] = 0.25f;                 /* k_xx_b: lower bound (math_func.py:1386) */
            d->bmode[1] = 0;                                     /* e_kxy_b is never the bounded kernel (1387-1390 vs 1402) */
            if (w1 > 0) { d->bmode[2] = 1; d->bval[2] = 0.25f; } /* math_func.py:1391-1392 */
            else { d->bmode[2] = 2; d->bval[2] = 4.0f; }         /* math_func.py:1393-1394 */
        }
        return MMDGAN_OK;
    }
    if (!strcmp(loss_type, "mmd_g") || !strcmp(loss_type, "fixed_g")) { /* math_func.py:2160-2173, sigma list 2108 */
        d->n_sigma = 5;
        d->sigma[0] = 1.0f; d->sigma[1] = sqrtf(2.0f); d->sigma[2] = 2.0f; d->sigma[3] = sqrtf(8.0f); d->sigma[4] = 4.0f;
        d->cD[0] = -1.0f; d->cD[1] = 2.0f; d->cD[2] = -1.0f;
        return MMDGAN_OK;
    }
    if (!strcmp(loss_type, "mmd_t") || !strcmp(loss_type, "fixed_t")) { /* math_func.py:2263-2275, alpha list 2109, beta 2110 */
        d->family = 1;
        d->n_sigma = 5;
        d->sigma[0] = 0.2f; d->sigma[1] = 0.5f; d->sigma[2] = 1.0f; d->sigma[3] = 2.0f; d->sigma[4] = 5.0f;
        d->beta = 2.0f;
        d->cD[0] = -1.0f; d->cD[1] = 2.0f; d->cD[2] = -1.0f;
        return MMDGAN_OK;
    }
    if (!strcmp(loss_type, "mgb")) { /* math_func.py:2175-2193 */
        d->cD[0] = -1.0f; d->cD[1] = 2.0f; d->cD[2] = -1.0f;
        d->bmode[0] = 1; d->bval[0] = 0.25f;
        d->bmode[1] = 2; d->bval[1] = 4.0f;
        d->bmode[2] = 1; d->bval[2] = 0.25f;
        return MMDGAN_OK;
    }
    return fail(MMDGAN_EINVAL, "Not implemented.");  /* math_func.py:2651 */
}
int mmdgan_sample_normal(float* out, long long n, unsigned long long seed, const unsigned long long* draw_counter, unsigned int* raw_words,
                         void* stream) {
    if (!out || n <= 0) return fail(MMDGAN_EINVAL, "mmdgan_sample_normal: null pointer / empty request");
    return wrap(mg::l_sample_normal(out, n, seed, draw_counter, raw_words, S(stream)), "mmdgan_sample_normal");
}
int mmdgan_incr_counter(unsigned long long* counter, void* stream) {
    if (!counter) return fail(MMDGAN_EINVAL, "mmdgan_incr_counter: null pointer");
    return wrap(mg::l_incr_u64(counter, S(stream)), "mmdgan_incr_counter");
}
int mmdgan_losses_from_sums(const float* sums, float cD0, float cD1, float cD2, float* losses, void* stream) {
    if (!sums || !losses) return fail(MMDGAN_EINVAL, "mmdgan_losses_from_sums: null pointer");
    return wrap(mg::l_losses_from_sums(sums, cD0, cD1, cD2, losses, S(stream)), "mmdgan_losses_from_sums");
}
size_t mmdgan_mmd_workspace(int b) {
    if (b <= 0) return 0;
    return static_cast<size_t>(mg::mmd_grid_blocks(b)) * 6 * sizeof(float) + 16;
}
int mmdgan_mmd_fwd_bwd(const mmdgan_mmd_desc* d, void* stream) {
    if (!d || !d->gen_loc || !d->real_loc || !d->gen_all || !d->real_all || !d->sums || !d->losses || !d->dLg_dgen || !d->dLd_dgen ||
        !d->dLd_dreal || !d->workspace)
        return fail(MMDGAN_EINVAL, "mmdgan_mmd_fwd_bwd: null pointer");
    if (d->b < 1 || d->Bg < 2 || d->row0 < 0 || d->row0 + d->b > d->Bg) return fail(MMDGAN_ESHAPE, "mmdgan_mmd_fwd_bwd: bad batch (b=%d Bg=%d row0=%d)", d->b, d->Bg, d->row0);
    if (d->d != 4 && d->d != 8 && d->d != 16 && d->d != 32 && d->d != 64)
        return fail(MMDGAN_ESHAPE, "mmdgan_mmd_fwd_bwd: score size %d not in {4,8,16,32,64} (zero-pad the scores)", d->d);
    if (d->n_sigma < 1 || d->n_sigma > 8) return fail(MMDGAN_EINVAL, "mmdgan_mmd_fwd_bwd: n_sigma out of range");
    mg::MmdParams p;
    memset(&p, 0, sizeof(p));
    p.gen_loc = d->gen_loc; p.real_loc = d->real_loc; p.gen_all = d->gen_all; p.real_all = d->real_all;
    p.b = d->b; p.Bg = d->Bg; p.row0 = d->row0; p.d = d->d; p.n_sigma = d->n_sigma;
    if (d->family != 0 && d->family != 1) return fail(MMDGAN_EINVAL, "mmdgan_mmd_fwd_bwd: unknown kernel family");
    if (d->family == 1 && !(d->beta > 0.f)) return fail(MMDGAN_EINVAL, "mmdgan_mmd_fwd_bwd: beta must be positive");
    p.family = d->family;
    p.inv_beta = d->family == 1 ? 1.0f / d->beta : 0.f;
    for (int i = 0; i < d->n_sigma; ++i) {
        if (!(d->sigma[i] > 0.f)) return fail(MMDGAN_EINVAL, "mmdgan_mmd_fwd_bwd: sigma / alpha must be positive");
        if (d->family == 0) {
            p.c_s[i] = 1.0f / (2.0f * d->sigma[i] * d->sigma[i]);
        } else {
            p.c_s[i] = d->sigma[i];
            p.c_t[i] = 1.0f / (d->sigma[i] * d->beta);
        }
    }
    for (int i = 0; i < 3; ++i) { p.cD[i] = d->cD[i]; p.bmode[i] = d->bmode[i]; p.bval[i] = d->bval[i]; }
    p.sums = d->sums; p.losses = d->losses; p.dLg_dgen = d->dLg_dgen; p.dLg_dreal = d->dLg_dreal; p.dLd_dgen = d->dLd_dgen;
    p.dLd_dreal = d->dLd_dreal;
    p.counter = reinterpret_cast<unsigned int*>(d->workspace);
    p.partials = reinterpret_cast<float*>(reinterpret_cast<char*>(d->workspace) + 16);
    return wrap(mg::launch_mmd(p, S(stream)), "mmdgan_mmd_fwd_bwd");
}

int mmdgan_refresh(const mmdgan_refresh_job* jobs_device, int njobs, long long total_blocks, void* stream) {
    if (!jobs_device) return fail(MMDGAN_EINVAL, "mmdgan_refresh: null pointer");
    if (njobs <= 0) return MMDGAN_OK;
    if (njobs > 256 || total_blocks <= 0) return fail(MMDGAN_ESHAPE, "mmdgan_refresh: bad job count");
    static_assert(sizeof(mmdgan_refresh_job) == sizeof(mg::RefreshJob), "job layout");
    return wrap(mg::l_refresh(reinterpret_cast<const mg::RefreshJob*>(jobs_device), njobs, total_blocks, S(stream)), "mmdgan_refresh");
}
int mmdgan_dense_small_fwd(const mmdgan_bf16* a, long long a_plane, int npl, int a_fmt, int rows, int K, const mmdgan_bf16* wt,
                           long long w_plane, int w_fmt, int kpad, int N, float alpha_k, const float* sigma, const float* bias, float* out, int ldo,
                           float* workspace, void* stream) {
    if (!a || !wt || !out || !workspace) return fail(MMDGAN_EINVAL, "mmdgan_dense_small_fwd: null pointer");
    if (rows <= 0 || K <= 0 || (K & 3) || kpad < K || (kpad & 3) || ldo < N) return fail(MMDGAN_ESHAPE, "mmdgan_dense_small_fwd: bad shape");
    if (!fmt_ok(a_fmt, npl) || !fmt_ok(w_fmt, npl) || (npl > 1 && (a_plane <= 0 || w_plane <= 0))) return fail(MMDGAN_ESHAPE, "mmdgan_dense_small_fwd: bad plane layout");
    if (N != 8 && N != 16 && N != 32) return fail(MMDGAN_ESHAPE, "mmdgan_dense_small_fwd: N must be 8/16/32");
    return wrap(mg::l_dense_small_fwd(a, a_plane, npl, a_fmt, rows, K, wt, w_plane, w_fmt, kpad, N, alpha_k, sigma, bias, out, ldo, workspace, S(stream)), "mmdgan_dense_small_fwd");
}
size_t mmdgan_dense_small_workspace(int rows, int K, int N) {
    if (rows <= 0 || K <= 0 || N <= 0) return 0;
    return static_cast<size_t>(mg::dense_small_workspace(rows, K, N));
}

int mmdgan_adam(float* w, float* m, float* v, const float* g, long long n, float lr, float beta1, float beta2, float eps,
                const int* step, void* stream) {
    if (!w || !m || !v || !g || !step) return fail(MMDGAN_EINVAL, "mmdgan_adam: null pointer");
    if (n <= 0) return MMDGAN_OK;
    return wrap(mg::l_adam(w, m, v, g, n, lr, beta1, beta2, eps, step, S(stream)), "mmdgan_adam");
}
int mmdgan_adam_allreduce_nvls(const float* w, const float* m, const float* v, const float* g_mc, float* w_mc, float* m_mc,
                               float* v_mc, long long begin, long long end, float lr, float beta1, float beta2, float eps,
                               const int* step, void* stream) {
    if (!w || !m || !v || !g_mc || !w_mc || !m_mc || !v_mc || !step)
        return fail(MMDGAN_EINVAL, "mmdgan_adam_allreduce_nvls: null pointer");
    if (begin < 0 || end < begin || (begin & 3) || (end & 3))
        return fail(MMDGAN_ESHAPE, "mmdgan_adam_allreduce_nvls: shard [%lld, %lld) must be a range of whole float4 groups", begin, end);
    if (!al16(w) || !al16(m) || !al16(v) || !al16(g_mc) || !al16(w_mc) || !al16(m_mc) || !al16(v_mc))
        return fail(MMDGAN_ESHAPE, "mmdgan_adam_allreduce_nvls: buffers must be 16-byte aligned");
    return wrap(mg::l_adam_allreduce_nvls(w, m, v, g_mc, w_mc, m_mc, v_mc, begin, end, lr, beta1, beta2, eps, step, S(stream)),
                "mmdgan_adam_allreduce_nvls");
}
int mmdgan_scatter_scores_nvls(const float* s_local, int b, int d, int rank, float* gen_all_mc, float* real_all_mc, void* stream) {
    if (!s_local || !gen_all_mc || !real_all_mc) return fail(MMDGAN_EINVAL, "mmdgan_scatter_scores_nvls: null pointer");
    if (b <= 0 || d <= 0 || (d & 3) || rank < 0) return fail(MMDGAN_ESHAPE, "mmdgan_scatter_scores_nvls: b %d, d %d (multiple of 4), rank %d", b, d, rank);
    if (!al16(s_local) || !al16(gen_all_mc) || !al16(real_all_mc)) return fail(MMDGAN_ESHAPE, "mmdgan_scatter_scores_nvls: buffers must be 16-byte aligned");
    return wrap(mg::l_scatter_scores_nvls(s_local, b, d, rank, gen_all_mc, real_all_mc, S(stream)), "mmdgan_scatter_scores_nvls");
}
int mmdgan_allreduce_small_nvls(float* out, const float* in_mc, int n, void* stream) {
    if (!out || !in_mc) return fail(MMDGAN_EINVAL, "mmdgan_allreduce_small_nvls: null pointer");
    if (n < 0 || (n & 3) || !al16(out) || !al16(in_mc)) return fail(MMDGAN_ESHAPE, "mmdgan_allreduce_small_nvls: n %d must be a multiple of 4, buffers 16-byte aligned", n);
    return wrap(mg::l_allreduce_small_nvls(out, in_mc, n, S(stream)), "mmdgan_allreduce_small_nvls");
}
int mmdgan_incr_step(int* step, void* stream) {
    if (!step) return fail(MMDGAN_EINVAL, "mmdgan_incr_step: null pointer");
    return wrap(mg::l_incr_step(step, S(stream)), "mmdgan_incr_step");
}
int mmdgan_nan_flag(const float* x, int n, int* flag, void* stream) {
    if (!x || !flag) return fail(MMDGAN_EINVAL, "mmdgan_nan_flag: null pointer");
    return wrap(mg::l_nan_flag(x, n, flag, S(stream)), "mmdgan_nan_flag");
}

}  // extern "C"
