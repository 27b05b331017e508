// HBM-bound helper kernels of the SNGan step: layout changes at the boundary (NCHW fp32 <-> NHWC bf16 planes), weight
// packing into the GEMM operand layouts, batch-norm statistics / apply / backward, reductions of per-tile partial
// sums (bias gradients, split-K weight gradients), the spectral-norm combine and the fused multi-tensor TF-Adam.
// All reductions use fixed orders (no floating-point atomics) so a step is bit-reproducible run to run.
//
// Reference call sites replaced: tf.layers.batch_normalization (GeneralTools/layer_func.py:953-966),
// tf.nn.bias_add gradients, SpectralNorm._l2_normalize_ (GeneralTools/math_func.py:653-659), the kernel * multiplier
// product and its gradient (layer_func.py:884-918), tf.train.AdamOptimizer.apply_gradients (graph_func.py:525-526).
#include "tc_common.cuh"
#include "conv_gemm.cuh"
#include <cuda_runtime.h>
#include <stdint.h>

namespace mg {

static inline int nblocks(long long n, int bs) { return static_cast<int>((n + bs - 1) / bs); }

// ------------------------------------------------------------------------------------------------ layout
// src NCHW [N][C][H][W] fp32 -> dst bf16 planes [npl][N][H][W][Cp] (channels >= C zero)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, bf16_t* __restrict__ dst, long long plane, int npl, int fmt, int N,
                                    int C, int H, int W, int Cp) {
    const long long total = static_cast<long long>(N) * H * W * Cp;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % Cp);
        long long r = i / Cp;
        const int w = static_cast<int>(r % W);
        r /= W;
        const int h = static_cast<int>(r % H);
        const int n = static_cast<int>(r / H);
        const float v = c < C ? src[((static_cast<long long>(n) * C + c) * H + h) * W + w] : 0.f;
        store_val(dst + i, plane, npl, fmt, v);
    }
}
// src planes [npl][N][H][W][Cp] -> dst NCHW [N][C][H][W] fp32 (the values the planes carry)
__global__ void nhwc_to_nchw_kernel(const bf16_t* __restrict__ src, long long plane, int npl, int fmt, float* __restrict__ dst, int N, int C,
                                    int H, int W, int Cp) {
    const long long total = static_cast<long long>(N) * C * H * W;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int w = static_cast<int>(i % W);
        long long r = i / W;
        const int h = static_cast<int>(r % H);
        r /= H;
        const int c = static_cast<int>(r % C);
        const int n = static_cast<int>(r / C);
        dst[i] = load_val(src, plane, npl, fmt, ((static_cast<long long>(n) * H + h) * W + w) * Cp + c);
    }
}
// fp32 [n] -> bf16 planes [npl][n] (same element order), and back
__global__ void to_planes_kernel(const float* __restrict__ x, bf16_t* __restrict__ dst, long long plane, int npl, int fmt, long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        store_val(dst + i, plane, npl, fmt, x[i]);
}
__global__ void from_planes_kernel(const bf16_t* __restrict__ src, long long plane, int npl, int fmt, float* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = load_val(src, plane, npl, fmt, i);
}

// planes in one format -> planes in another (same element order), four elements per thread.  The tensor core cannot mix an
// fp16 operand with a bf16 one, so the weight-gradient GEMM reads a bf16 re-split of the fp16 forward activations.
__global__ void convert_planes_kernel(const bf16_t* __restrict__ src, long long sp, int snpl, int sfmt, bf16_t* __restrict__ dst,
                                      long long dp, int dnpl, int dfmt, long long n) {
    for (long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 4; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x * 4)
        store_vals4(dst + i, dp, dnpl, dfmt, load_vals4(src, sp, snpl, sfmt, i));
}

// ------------------------------------------------------------------------------------------------ weight packing
enum { PACK_CONV_FWD = 0, PACK_CONV_DGRAD_S1 = 1, PACK_CONV_DGRAD_S2 = 2, PACK_TC_FWD = 3, PACK_TC_DGRAD = 4,
       PACK_DENSE_FWD = 5, PACK_DENSE_DGRAD = 6 };

__device__ __forceinline__ int perm_feature(int j, int C, int HW) {  // internal (hw*C + c) -> canonical (c*HW + hw)
    if (HW <= 1) return j;
    const int hw = j / C, c = j - hw * C;
    return c * HW + hw;
}

// value of packed element (class cls, row, col) read from the canonical weights
__device__ __forceinline__ float pack_value(const PackParams& p, int cls, int row, int col) {
    {
        // K order of the gather-GEMM: (channel chunk of CW = min(Cs, 64), tap, channel within the chunk)
        const int cw = p.Cs >= kGemmBK ? kGemmBK : p.Cs;
        const int ntaps = (p.mode == PACK_CONV_DGRAD_S2 || p.mode == PACK_TC_FWD) ? 4 : ((p.mode >= PACK_DENSE_FWD) ? 1 : p.k * p.k);
        const int grp = col / cw;
        const int cc = grp / ntaps;
        int tap = grp - cc * ntaps;
        const int ch = cc * cw + (col - grp * cw);
        if (ch >= p.Cs) tap = 1 << 20;   // K padding
        const int ph = cls >> 1, pw = cls & 1;
        float v = 0.f;
        switch (p.mode) {
            case PACK_CONV_FWD: {  // row = co, col = (kh*k+kw)*Cs + ci
                const int kh = tap / p.k, kw = tap - kh * p.k;
                if (row < p.Cout && tap < p.k * p.k && ch < p.Cin)
                    v = p.w[((static_cast<long long>(kh) * p.k + kw) * p.Cin + ch) * p.Cout + row];
                break;
            }
            case PACK_CONV_DGRAD_S1: {  // row = ci, col = (a*k+b)*Cs + co, kh = k-1-a
                const int a = tap / p.k, b = tap - a * p.k;
                if (row < p.Cin && tap < p.k * p.k && ch < p.Cout)
                    v = p.w[((static_cast<long long>(p.k - 1 - a) * p.k + (p.k - 1 - b)) * p.Cin + row) * p.Cout + ch];
                break;
            }
            case PACK_CONV_DGRAD_S2: {  // k4 s2, class (ph,pw): row = ci, col = (a*2+b)*Cs + co, kh = 3-ph-2a
                const int a = tap >> 1, b = tap & 1;
                if (row < p.Cin && tap < 4 && ch < p.Cout)
                    v = p.w[((static_cast<long long>(3 - ph - 2 * a) * 4 + (3 - pw - 2 * b)) * p.Cin + row) * p.Cout + ch];
                break;
            }
            case PACK_TC_FWD: {  // canon [k][k][Cout][Cin], class (ph,pw): row = co, col = (a*2+b)*Cs + ci
                const int a = tap >> 1, b = tap & 1;
                if (row < p.Cout && tap < 4 && ch < p.Cin)
                    v = p.w[((static_cast<long long>(3 - ph - 2 * a) * 4 + (3 - pw - 2 * b)) * p.Cout + row) * p.Cin + ch];
                break;
            }
            case PACK_TC_DGRAD: {  // row = ci, col = (kh*k+kw)*Cs + co
                const int kh = tap / p.k, kw = tap - kh * p.k;
                if (row < p.Cin && tap < p.k * p.k && ch < p.Cout)
                    v = p.w[((static_cast<long long>(kh) * p.k + kw) * p.Cout + ch) * p.Cin + row];
                break;
            }
            case PACK_DENSE_FWD: {  // row = out', col = in'
                if (row < p.Cout && col < p.Cin)
                    v = p.w[static_cast<long long>(perm_feature(col, p.in_C, p.in_HW)) * p.Cout + perm_feature(row, p.out_C, p.out_HW)];
                break;
            }
            case PACK_DENSE_DGRAD: {  // row = in', col = out'
                if (row < p.Cin && col < p.Cout)
                    v = p.w[static_cast<long long>(perm_feature(row, p.in_C, p.in_HW)) * p.Cout + perm_feature(col, p.out_C, p.out_HW)];
                break;
            }
        }
        return v;
    }
}
// ---- the same mapping split into its column part, its row part and the fetch, so that the integer divisions are done once per
// tile column / tile row instead of once per element
__device__ __forceinline__ void pack_col_info(const PackParams& p, int col, int& tap, int& ch) {
    if (p.mode >= PACK_DENSE_FWD) {
        const int lim = p.mode == PACK_DENSE_FWD ? p.Cin : p.Cout;
        tap = col < lim ? 0 : (1 << 20);
        ch = col < lim ? (p.mode == PACK_DENSE_FWD ? perm_feature(col, p.in_C, p.in_HW) : perm_feature(col, p.out_C, p.out_HW)) : 0;
        return;
    }
    const int cw = p.Cs >= kGemmBK ? kGemmBK : p.Cs;
    const int ntaps = (p.mode == PACK_CONV_DGRAD_S2 || p.mode == PACK_TC_FWD) ? 4 : p.k * p.k;
    const int grp = col / cw;
    const int cc = grp / ntaps;
    tap = grp - cc * ntaps;
    ch = cc * cw + (col - grp * cw);
    const int chlim = (p.mode == PACK_CONV_FWD || p.mode == PACK_TC_FWD) ? p.Cin : p.Cout;
    if (ch >= p.Cs || ch >= chlim) tap = 1 << 20;   // K padding / channel padding
}
// row part: (class, source row index, valid)
__device__ __forceinline__ void pack_row_info(const PackParams& p, int grow, int& cls, int& row) {
    cls = grow / p.rows_pad;
    row = grow - cls * p.rows_pad;
    const int rlim = (p.mode == PACK_CONV_FWD || p.mode == PACK_TC_FWD || p.mode == PACK_DENSE_FWD) ? p.Cout : p.Cin;
    if (row >= rlim) { row = -1; return; }
    if (p.mode == PACK_DENSE_FWD) row = perm_feature(row, p.out_C, p.out_HW);
    else if (p.mode == PACK_DENSE_DGRAD) row = perm_feature(row, p.in_C, p.in_HW);
}
__device__ __forceinline__ float pack_fetch(const PackParams& p, int cls, int row, int tap, int ch) {
    if (row < 0 || tap >= (1 << 20)) return 0.f;
    const int ph = cls >> 1, pw = cls & 1;
    switch (p.mode) {
        case PACK_CONV_FWD: {
            const int kh = tap / p.k, kw = tap - kh * p.k;
            return p.w[((static_cast<long long>(kh) * p.k + kw) * p.Cin + ch) * p.Cout + row];
        }
        case PACK_CONV_DGRAD_S1: {
            const int a = tap / p.k, b = tap - a * p.k;
            return p.w[((static_cast<long long>(p.k - 1 - a) * p.k + (p.k - 1 - b)) * p.Cin + row) * p.Cout + ch];
        }
        case PACK_CONV_DGRAD_S2: {
            const int a = tap >> 1, b = tap & 1;
            return p.w[((static_cast<long long>(3 - ph - 2 * a) * 4 + (3 - pw - 2 * b)) * p.Cin + row) * p.Cout + ch];
        }
        case PACK_TC_FWD: {
            const int a = tap >> 1, b = tap & 1;
            return p.w[((static_cast<long long>(3 - ph - 2 * a) * 4 + (3 - pw - 2 * b)) * p.Cout + row) * p.Cin + ch];
        }
        case PACK_TC_DGRAD: {
            const int kh = tap / p.k, kw = tap - kh * p.k;
            return p.w[((static_cast<long long>(kh) * p.k + kw) * p.Cout + ch) * p.Cin + row];
        }
        case PACK_DENSE_FWD: return p.w[static_cast<long long>(ch) * p.Cout + row];      // ch = permuted input feature, row = permuted output
        case PACK_DENSE_DGRAD: return p.w[static_cast<long long>(row) * p.Cout + ch];    // row = permuted input feature, ch = permuted output
    }
    return 0.f;
}

// One 32 x 64 tile (rows x K columns = one channel chunk) of a packed operand per block iteration, 256 threads.  The
// canonical layouts are contiguous along the packed ROW index for half of the modes (conv forward, transposed-conv input
// gradient, dense forward) and along the packed COLUMN index for the others; the tile is read along whichever index is
// contiguous in the source and written along the columns (contiguous in the destination) through shared memory -- eight
// consecutive K elements per thread, one 8-byte store per plane and quadruple -- so both sides of the repack are coalesced.
struct PackSmem {
    float tile[32][68];
    int ctap[64], cch[64], rcls[32], rrow[32];
};
__device__ __forceinline__ void pack_tiles(const PackParams& p, PackSmem& sm, int blk, int nblk) {
    const int t = threadIdx.x;
    const int rows_all = p.rows_pad * p.classes;
    const int tr = (rows_all + 31) >> 5, tc = (p.kpad + 63) >> 6;
    const bool row_contig = p.mode == PACK_CONV_FWD || p.mode == PACK_TC_DGRAD || p.mode == PACK_DENSE_FWD;
    for (int tidx = blk; tidx < tr * tc; tidx += nblk) {
        const int r0 = (tidx / tc) << 5, c0 = (tidx % tc) << 6;
        if (t < 64) {
            int tap = 1 << 20, ch = 0;
            if (c0 + t < p.kpad) pack_col_info(p, c0 + t, tap, ch);
            sm.ctap[t] = tap;
            sm.cch[t] = ch;
        } else if (t < 96) {
            int cls = 0, row = -1;
            if (r0 + t - 64 < rows_all) pack_row_info(p, r0 + t - 64, cls, row);
            sm.rcls[t - 64] = cls;
            sm.rrow[t - 64] = row;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            // read: consecutive lanes run along the source-contiguous index
            const int rr = row_contig ? (t & 31) : (t >> 6) + 4 * j;
            const int cc = row_contig ? (t >> 5) + 8 * j : (t & 63);
            sm.tile[rr][cc] = pack_fetch(p, sm.rcls[rr], sm.rrow[rr], sm.ctap[cc], sm.cch[cc]);
        }
        __syncthreads();
        {
            const int rr = t >> 3, cg = (t & 7) * 8, grow = r0 + rr, col = c0 + cg;
            if (grow < rows_all && col < p.kpad) {
                const float4 a = *reinterpret_cast<const float4*>(&sm.tile[rr][cg]);
                const float4 b = *reinterpret_cast<const float4*>(&sm.tile[rr][cg + 4]);
                bf16_t* o = p.out + static_cast<long long>(grow) * p.kpad + col;
                store_vals4(o, p.plane, p.npl, p.fmt, a);
                store_vals4(o + 4, p.plane, p.npl, p.fmt, b);
            }
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) pack_weights_kernel(const PackParams p) {
    __shared__ __align__(16) PackSmem sm;
    pack_tiles(p, sm, blockIdx.x, gridDim.x);
}

// out[j'] = src[perm(j')]: canonical per-feature vector (bias / gamma / beta) -> internal NHWC-flatten order
__global__ void permute_features_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, int C, int HW, int inverse) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    if (inverse) dst[perm_feature(j, C, HW)] = src[j];
    else dst[j] = src[perm_feature(j, C, HW)];
}

// One launch refreshes every parameter-derived buffer of a net after an update: weight packing jobs and feature
// permutation / padding jobs.  The jobs live in device memory (built once at start-up); job j owns the blocks
// [block_start_j, block_start_{j+1}) of a flat grid, sized to its own tile count.  (A 2-D grid of (largest job's blocks) x jobs
// launched 41 000 blocks for the discriminator, three quarters of them without a tile: 83 us of block scheduling.)
__global__ void __launch_bounds__(256) refresh_kernel(const RefreshJob* __restrict__ jobs, int njobs) {
    __shared__ __align__(16) PackSmem sm;
    const int bx = static_cast<int>(blockIdx.x);
    // the job of this block = number of jobs whose first block is <= blockIdx.x, minus one (njobs <= 256: one lookup per thread)
    const int j = __syncthreads_count(static_cast<int>(threadIdx.x) < njobs && bx >= jobs[threadIdx.x].block_start) - 1;
    const RefreshJob& job = jobs[j];
    const int blk = bx - job.block_start;
    const int nblk = (j + 1 < njobs ? jobs[j + 1].block_start : static_cast<int>(gridDim.x)) - job.block_start;
    if (job.kind == 0) {
        pack_tiles(job.pack, sm, blk, nblk);
    } else {
        for (int i = blk * blockDim.x + threadIdx.x; i < job.n; i += nblk * blockDim.x) {
            if (job.inverse) job.dst[perm_feature(i, job.C, job.HW)] = job.src[i];
            else job.dst[i] = job.src[perm_feature(i, job.C, job.HW)];
        }
    }
}

// out[m][n] = alpha * sum_k A[m][k] * Wt[n][k] + bias[n] for a handful of output columns (the 16 critic scores): fp32 FFMA on
// the values reassembled from the 16-bit planes.  The tensor-core tile would be 94 % padding here (N = 16 of 128 lanes x
// 8192 deep on 4 CTAs).  Split K: block (row group of 128 / N rows, K slice of 1024) -- every thread owns 4 consecutive k, issues
// all its 8-byte plane loads at once (A: 128 / N rows, W: N rows), forms the 128 partial products of its k quadruple, and the
// block reduces them with the recursive-halving lane transpose (124 shuffles for 128 values) + one shared-memory pass.  The
// K-slice partials [slices][rows][N] are summed in a fixed order by dense_small_finish_kernel (deterministic, no atomics).
// (The first version, one block per row sweeping all of K, re-read the 0.5 MB of weights 512 times and took ~90 us on the
// critical path between the discriminator forward and the loss.)
static constexpr int kDsSlice = 1024;
template <int N>
__global__ void __launch_bounds__(256) dense_small_fwd_kernel(const bf16_t* __restrict__ a, long long a_plane, int npl, int a_fmt, int rows,
                                                             int K, const bf16_t* __restrict__ wt, long long w_plane, int w_fmt, int kpad,
                                                             float* __restrict__ partials) {
    constexpr int kDsRows = 128 / N;               // rows per block: 16 / 8 / 4
    constexpr int V = kDsRows * N;                 // 128 partial products per thread
    __shared__ float red[8][V];
    const int row0 = blockIdx.x * kDsRows;
    const int slice = blockIdx.y;
    const int k = slice * kDsSlice + threadIdx.x * 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc[V];
    if (k < K) {
        float4 x[kDsRows];
#pragma unroll
        for (int r = 0; r < kDsRows; ++r)
            x[r] = row0 + r < rows ? load_vals4(a, a_plane, npl, a_fmt, static_cast<long long>(row0 + r) * K + k) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int n = 0; n < N; ++n) {
            const float4 w = load_vals4(wt, w_plane, npl, w_fmt, static_cast<long long>(n) * kpad + k);
#pragma unroll
            for (int r = 0; r < kDsRows; ++r)
                acc[r * N + n] = fmaf(x[r].x, w.x, fmaf(x[r].y, w.y, fmaf(x[r].z, w.z, x[r].w * w.w)));
        }
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = 0.f;
    }
    // lane transpose-reduce: after the steps lane l holds the warp totals of values [l * V/32, (l + 1) * V/32)
#pragma unroll
    for (int off = 16, n = V / 2; off >= 1; off >>= 1, n >>= 1) {
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const float keep = hi ? acc[i + n] : acc[i];
            const float send = hi ? acc[i] : acc[i + n];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    constexpr int PER = V / 32;
    // lane l ends up with the block of values whose index has bit pattern (b4 b3 b2 b1 b0) = l in its top five bits
#pragma unroll
    for (int i = 0; i < PER; ++i) red[warp][lane * PER + i] = acc[i];
    __syncthreads();
    for (int i = threadIdx.x; i < V; i += 256) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][i];
        const int r = i / N, n = i - r * N;
        if (row0 + r < rows) partials[(static_cast<long long>(slice) * rows + row0 + r) * N + n] = s;
    }
}
__global__ void dense_small_finish_kernel(const float* __restrict__ partials, int slices, int rows, int N, float alpha_k,
                                          const float* __restrict__ sigma, const float* __restrict__ bias, float* __restrict__ out, int ldo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * N) return;
    float s = 0.f;
    for (int t = 0; t < slices; ++t) s += partials[static_cast<long long>(t) * rows * N + i];
    const float alpha = sigma ? alpha_k / __ldg(sigma) : alpha_k;
    const int r = i / N, n = i - r * N;
    out[static_cast<long long>(r) * ldo + n] = fmaf(s, alpha, bias ? bias[n] : 0.f);
}

// ------------------------------------------------------------------------------------------------ reductions
// out[c] = scale * sum_t partials[t][C] over T tiles; block = 32 columns x 32 row-slices, double accumulation, fixed order
__global__ void reduce_tiles_kernel(const float* __restrict__ partials, int T, int C, float scale, float* __restrict__ out) {
    __shared__ double red[32][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    double s = 0.0;
    if (c < C)
        for (int t = threadIdx.y; t < T; t += 32 * 8) {      // eight loads in flight, summed in the original order
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = t + 32 * k < T ? partials[static_cast<long long>(t + 32 * k) * C + c] : 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) s += static_cast<double>(v[k]);
        }
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        double a = 0.0;
        for (int k = 0; k < 32; ++k) a += red[k][threadIdx.x];
        out[c] = static_cast<float>(a * scale);
    }
}
// column sums of x[rows][C] (small matrices: score gradients)
__global__ void colsum_small_kernel(const float* __restrict__ x, int rows, int C, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0;
    for (int r = 0; r < rows; ++r) s += static_cast<double>(x[static_cast<long long>(r) * C + c]);
    out[c] = static_cast<float>(s);
}

// column sums of a small bf16-plane matrix x[npl][rows][C] (the [B, F] gradient of a dense layer with a per-feature bias)
// block = 32 columns x 32 row slices, double accumulation, fixed order
__global__ void colsum_planes_kernel(const bf16_t* __restrict__ x, long long plane, int npl, int rows, int C, float* __restrict__ out) {
    __shared__ double red[32][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    double s = 0.0;
    if (c < C)
        for (int r = threadIdx.y; r < rows; r += 32) s += static_cast<double>(load_planes(x, plane, npl, static_cast<long long>(r) * C + c));
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        double a = 0.0;
        for (int k = 0; k < 32; ++k) a += red[k][threadIdx.x];
        out[c] = static_cast<float>(a);
    }
}

__device__ __forceinline__ void wgrad_reduce_body(const WredParams& p, int block, int nblocks_, double* red) {
    // four consecutive elements per thread (NC is a multiple of 8: same row, same tap), 16-byte loads, four partial tiles in
    // flight per accumulator set: the kernel is a latency-bound sum over up to 56 partial tiles
    const long long total = static_cast<long long>(p.R) * p.NC;
    double dot = 0.0;
    for (long long i = (block * static_cast<long long>(blockDim.x) + threadIdx.x) * 4; i < total;
         i += static_cast<long long>(nblocks_) * blockDim.x * 4) {
        const int r = static_cast<int>(i / p.NC);
        const int col = static_cast<int>(i - static_cast<long long>(r) * p.NC);
        const int t = col / p.Cg, c = col - t * p.Cg;
        if (c >= p.Cvalid || r >= p.Rvalid) continue;
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
        const float* base = p.partials + i;
        int z = 0;
        for (; z + 4 <= p.splits; z += 4) {     // the summation order stays fixed
            const float4 a = __ldcs(reinterpret_cast<const float4*>(base + static_cast<long long>(z) * total));
            const float4 b = __ldcs(reinterpret_cast<const float4*>(base + static_cast<long long>(z + 1) * total));
            const float4 cc = __ldcs(reinterpret_cast<const float4*>(base + static_cast<long long>(z + 2) * total));
            const float4 d = __ldcs(reinterpret_cast<const float4*>(base + static_cast<long long>(z + 3) * total));
            s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
            s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
            s2.x += cc.x; s2.y += cc.y; s2.z += cc.z; s2.w += cc.w;
            s3.x += d.x; s3.y += d.y; s3.z += d.z; s3.w += d.w;
        }
        for (; z < p.splits; ++z) {
            const float4 a = __ldcs(reinterpret_cast<const float4*>(base + static_cast<long long>(z) * total));
            s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
        }
        const float v[4] = {((s0.x + s1.x) + (s2.x + s3.x)) * p.scale, ((s0.y + s1.y) + (s2.y + s3.y)) * p.scale,
                            ((s0.z + s1.z) + (s2.z + s3.z)) * p.scale, ((s0.w + s1.w) + (s2.w + s3.w)) * p.scale};
        const long long rbase = p.base + perm_feature(r, p.r_perm_C, p.r_perm_HW) * p.sr + t * p.st;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (c + e >= p.Cvalid) break;
            const long long ci = rbase + perm_feature(c + e, p.c_perm_C, p.c_perm_HW) * p.sc;
            p.out[ci] = v[e];
            if (p.w) dot += static_cast<double>(v[e]) * static_cast<double>(p.w[ci]);
        }
    }
    if (p.dots) {
        red[threadIdx.x] = dot;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) p.dots[block] = red[0];
    }
}
__global__ void wgrad_reduce_kernel(const WredParams p) {
    __shared__ double red[256];
    wgrad_reduce_body(p, blockIdx.x, gridDim.x, red);
}
// Every layer of a net in ONE launch (after its last weight-gradient GEMM): job j owns blocks [start[j], start[j + 1]).  The 28
// per-layer launches of a step were latency bound (each sums up to 56 partials per element, 10-19 us for a few MB) and cost
// 0.3 ms of step time; batched, their latencies overlap.
__global__ void wgrad_reduce_batched_kernel(const WredParams* __restrict__ jobs, const int* __restrict__ start, int njobs) {
    __shared__ double red[256];
    int j = 0;
    while (j + 1 < njobs && static_cast<int>(blockIdx.x) >= start[j + 1]) ++j;
    wgrad_reduce_body(jobs[j], blockIdx.x - start[j], start[j + 1] - start[j], red);
}
struct SnCombineJob {
    float* g;
    const float* s;
    const double* dots;
    const float* sigma;
    long long n;
    int ndots;
    float act_k;
};
__device__ __forceinline__ void sn_grad_combine_body(float* g, const float* s, const double* dots, int ndots, const float* sigma, float act_k,
                                                     long long n, int block, int nblocks_, double* red) {
    double d = 0.0;
    for (int i = threadIdx.x; i < ndots; i += blockDim.x) d += dots[i];     // fixed order per thread, fixed tree below
    red[threadIdx.x] = d;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    const double dsum = red[0];
    const float sg = *sigma;
    const float m = act_k / sg;
    const float coef = static_cast<float>(static_cast<double>(m) / static_cast<double>(sg) * dsum);
    for (long long i = block * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(nblocks_) * blockDim.x)
        g[i] = m * g[i] - coef * s[i];
}
// blockIdx.y = job
__global__ void sn_grad_combine_batched_kernel(const SnCombineJob* __restrict__ jobs) {
    __shared__ double red[256];
    const SnCombineJob j = jobs[blockIdx.y];
    sn_grad_combine_body(j.g, j.s, j.dots, j.ndots, j.sigma, j.act_k, j.n, blockIdx.x, gridDim.x, red);
}
// grad = m * G - (m / sigma) * <G, W> * S,  m = act_k / sigma  (layer_func.py:884-918 differentiated)
__global__ void sn_grad_combine_kernel(float* __restrict__ g, const float* __restrict__ s, const double* __restrict__ dots, int ndots,
                                       const float* __restrict__ sigma, float act_k, long long n) {
    __shared__ double red[256];
    double d = 0.0;
    for (int i = threadIdx.x; i < ndots; i += blockDim.x) d += dots[i];     // fixed order per thread, fixed tree below
    red[threadIdx.x] = d;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    const double dsum = red[0];
    const float sg = *sigma;
    const float m = act_k / sg;
    const float coef = static_cast<float>(static_cast<double>(m) / static_cast<double>(sg) * dsum);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        g[i] = m * g[i] - coef * s[i];
}
__global__ void scale_by_sigma_kernel(float* __restrict__ g, const float* __restrict__ sigma, float act_k, long long n) {
    const float m = act_k / *sigma;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        g[i] *= m;
}

// ------------------------------------------------------------------------------------------------ spectral norm
// sigma = ||v||, out planes = v / (sigma + eps); one block (v has at most a few 10^4 elements)
__global__ void sn_normalize_kernel(const float* __restrict__ v, long long n, float eps, float* __restrict__ sigma_out,
                                    bf16_t* __restrict__ out, long long plane, int npl, int fmt) {
    __shared__ double red[1024];
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) s += static_cast<double>(v[i]) * static_cast<double>(v[i]);
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    const float nrm = static_cast<float>(sqrt(red[0]));
    if (threadIdx.x == 0 && sigma_out) *sigma_out = nrm;
    const float inv = 1.0f / (nrm + eps);
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        store_val(out + i, plane, npl, fmt, v[i] * inv);
    }
}

// ------------------------------------------------------------------------------------------------ batch norm
// statistics from per-tile partial sums; moving stats as tf.layers.batch_normalization(fused=True): biased variance
// normalises; the moving average (momentum 0.99) is fed the Bessel-corrected variance for rank-4 inputs (fused kernel) and the
// biased one for rank-2 inputs (TF 1.8 falls back to nn.moments there): `bessel`
// CPB channels per block, 1024 / CPB tile slices per channel: a layer with few channels and many tiles (64 channels x 2048
// tiles) gets 8 blocks of 128 slices instead of 2 blocks of 32 (48 us on the forward critical path).
template <int CPB>
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float* __restrict__ psum, const float* __restrict__ psq, int T, int C,
                                                           long long rows, float eps, float momentum, float* __restrict__ mean,
                                                           float* __restrict__ invstd, float* __restrict__ moving_mean,
                                                           float* __restrict__ moving_var, int bessel) {
    constexpr int RT = 1024 / CPB;
    __shared__ double rs[RT][CPB + 1], rq[RT][CPB + 1];
    const int c = blockIdx.x * CPB + threadIdx.x;
    double s = 0.0, q = 0.0;
    if (c < C)
        for (int t = threadIdx.y; t < T; t += RT * 4) {      // eight loads in flight, summed in a fixed order
            float a[4], b[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool ok = t + RT * k < T;
                a[k] = ok ? psum[static_cast<long long>(t + RT * k) * C + c] : 0.f;
                b[k] = ok ? psq[static_cast<long long>(t + RT * k) * C + c] : 0.f;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                s += static_cast<double>(a[k]);
                q += static_cast<double>(b[k]);
            }
        }
    rs[threadIdx.y][threadIdx.x] = s;
    rq[threadIdx.y][threadIdx.x] = q;
    __syncthreads();
    // fold the RT slices down to 32 in parallel, then one thread per channel adds the 32 in order
    for (int half = RT / 2; half >= 32; half >>= 1) {
        if (threadIdx.y < half) {
            rs[threadIdx.y][threadIdx.x] += rs[threadIdx.y + half][threadIdx.x];
            rq[threadIdx.y][threadIdx.x] += rq[threadIdx.y + half][threadIdx.x];
        }
        __syncthreads();
    }
    if (threadIdx.y != 0 || c >= C) return;
    s = 0.0; q = 0.0;
    for (int k = 0; k < 32; ++k) { s += rs[k][threadIdx.x]; q += rq[k][threadIdx.x]; }
    const double mu = s / rows;
    double var = q / rows - mu * mu;
    if (var < 0.0) var = 0.0;
    mean[c] = static_cast<float>(mu);
    invstd[c] = static_cast<float>(1.0 / sqrt(var + eps));
    if (moving_mean) {
        const double var_u = (bessel && rows > 1) ? var * (static_cast<double>(rows) / (rows - 1)) : var;
        moving_mean[c] = moving_mean[c] * momentum + static_cast<float>(mu) * (1.0f - momentum);
        moving_var[c] = moving_var[c] * momentum + static_cast<float>(var_u) * (1.0f - momentum);
    }
}
// inference-mode statistics (tf.layers.batch_normalization(training=False)): the moving averages normalise
__global__ void bn_inference_stats_kernel(const float* __restrict__ moving_mean, const float* __restrict__ moving_var, int C, float eps,
                                          float* __restrict__ mean, float* __restrict__ invstd) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    mean[c] = moving_mean[c];
    invstd[c] = static_cast<float>(1.0 / sqrt(static_cast<double>(moving_var[c]) + eps));
}
// a = act(gamma * (z - mean) * invstd + beta) -> bf16 planes; z [rows][C] raw fp32
__global__ void bn_apply_kernel(const float* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ invstd,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int C, long long total, int act,
                                bf16_t* __restrict__ out, long long plane, int npl, int fmt, int* __restrict__ sat_flag) {
    for (long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 4; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x * 4) {
        const int c = static_cast<int>(i % C);
        const float4 zv = *reinterpret_cast<const float4*>(z + i);
        const float4 mu = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
        const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
        float4 y;
        y.x = fmaf((zv.x - mu.x) * is.x, g.x, b.x);
        y.y = fmaf((zv.y - mu.y) * is.y, g.y, b.y);
        y.z = fmaf((zv.z - mu.z) * is.z, g.z, b.z);
        y.w = fmaf((zv.w - mu.w) * is.w, g.w, b.w);
        if (act == 2) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
        note_saturation4(sat_flag, fmt, y);
        store_vals4(out + i, plane, npl, fmt, y);
    }
}
// per-block partial sums of dy and dy*xhat per channel; dy = da * act'(bn output); block handles a row slab
__global__ void bn_bwd_reduce_kernel(const float* __restrict__ da, const float* __restrict__ z, const float* __restrict__ mean,
                                     const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                                     int C, long long rows, int rows_per_block, int act, float* __restrict__ psum,
                                     float* __restrict__ psumx) {
    const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > rows) r1 = rows;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float mu = mean[c], is = invstd[c], g = gamma[c], b = beta[c];
        float s = 0.f, sx = 0.f;
        for (long long r = r0; r < r1; ++r) {
            const float xh = (z[r * C + c] - mu) * is;
            float dy = da[r * C + c];
            if (act == 2 && fmaf(xh, g, b) <= 0.f) dy = 0.f;
            s += dy;
            sx = fmaf(dy, xh, sx);
        }
        psum[static_cast<long long>(blockIdx.x) * C + c] = s;
        psumx[static_cast<long long>(blockIdx.x) * C + c] = sx;
    }
}
// same sums, for C <= 1024: a thread owns four consecutive channels (float4 loads) of every (256 / (C/4))-th row of the
// slab, the row groups are combined through shared memory in a fixed order
__global__ void __launch_bounds__(256) bn_bwd_reduce_vec_kernel(const float* __restrict__ da, const float* __restrict__ z,
                                                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta, int C,
                                                                long long rows, int rows_per_block, int act, float* __restrict__ psum,
                                                                float* __restrict__ psumx) {
    __shared__ float4 rs[256], rx[256];
    const int cq = C >> 2, ng = 256 / cq;
    const int c4 = threadIdx.x % cq, rg = threadIdx.x / cq;
    const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > rows) r1 = rows;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), sx = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rg < ng) {
        const int c = c4 * 4;
        const float4 mu = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
        const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
        for (long long r = r0 + rg; r < r1; r += ng) {
            const float4 zv = *reinterpret_cast<const float4*>(z + r * C + c);
            float4 dy = *reinterpret_cast<const float4*>(da + r * C + c);
            const float4 xh = make_float4((zv.x - mu.x) * is.x, (zv.y - mu.y) * is.y, (zv.z - mu.z) * is.z, (zv.w - mu.w) * is.w);
            if (act == 2) {
                if (fmaf(xh.x, g.x, b.x) <= 0.f) dy.x = 0.f;
                if (fmaf(xh.y, g.y, b.y) <= 0.f) dy.y = 0.f;
                if (fmaf(xh.z, g.z, b.z) <= 0.f) dy.z = 0.f;
                if (fmaf(xh.w, g.w, b.w) <= 0.f) dy.w = 0.f;
            }
            s.x += dy.x; s.y += dy.y; s.z += dy.z; s.w += dy.w;
            sx.x = fmaf(dy.x, xh.x, sx.x); sx.y = fmaf(dy.y, xh.y, sx.y); sx.z = fmaf(dy.z, xh.z, sx.z); sx.w = fmaf(dy.w, xh.w, sx.w);
        }
    }
    rs[threadIdx.x] = s;
    rx[threadIdx.x] = sx;
    __syncthreads();
    if (rg == 0) {
        for (int k = 1; k < ng; ++k) {
            const float4 a = rs[k * cq + c4], bq = rx[k * cq + c4];
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
            sx.x += bq.x; sx.y += bq.y; sx.z += bq.z; sx.w += bq.w;
        }
        *reinterpret_cast<float4*>(psum + static_cast<long long>(blockIdx.x) * C + c4 * 4) = s;
        *reinterpret_cast<float4*>(psumx + static_cast<long long>(blockIdx.x) * C + c4 * 4) = sx;
    }
}
// dz = gamma * invstd * (dy - mean(dy) - xhat * mean(dy*xhat)) -> bf16 planes
__global__ void bn_bwd_apply_kernel(const float* __restrict__ da, const float* __restrict__ z, const float* __restrict__ mean,
                                    const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ dbeta, const float* __restrict__ dgamma, int C, long long rows, int act,
                                    bf16_t* __restrict__ out, long long plane, int npl) {
    const long long total = rows * C;
    const float inv_rows = 1.0f / static_cast<float>(rows);
    if ((C & 3) == 0) {      // four consecutive channels per thread: 16-byte loads, 8-byte plane stores
        for (long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 4; i < total;
             i += static_cast<long long>(gridDim.x) * blockDim.x * 4) {
            const int c = static_cast<int>(i % C);
            const float4 is = *reinterpret_cast<const float4*>(invstd + c), g = *reinterpret_cast<const float4*>(gamma + c);
            const float4 mu = *reinterpret_cast<const float4*>(mean + c), b = *reinterpret_cast<const float4*>(beta + c);
            const float4 db = *reinterpret_cast<const float4*>(dbeta + c), dg = *reinterpret_cast<const float4*>(dgamma + c);
            const float4 zv = *reinterpret_cast<const float4*>(z + i);
            float4 dy = *reinterpret_cast<const float4*>(da + i);
            const float4 xh = make_float4((zv.x - mu.x) * is.x, (zv.y - mu.y) * is.y, (zv.z - mu.z) * is.z, (zv.w - mu.w) * is.w);
            if (act == 2) {
                if (fmaf(xh.x, g.x, b.x) <= 0.f) dy.x = 0.f;
                if (fmaf(xh.y, g.y, b.y) <= 0.f) dy.y = 0.f;
                if (fmaf(xh.z, g.z, b.z) <= 0.f) dy.z = 0.f;
                if (fmaf(xh.w, g.w, b.w) <= 0.f) dy.w = 0.f;
            }
            float4 v;
            v.x = g.x * is.x * (dy.x - db.x * inv_rows - xh.x * dg.x * inv_rows);
            v.y = g.y * is.y * (dy.y - db.y * inv_rows - xh.y * dg.y * inv_rows);
            v.z = g.z * is.z * (dy.z - db.z * inv_rows - xh.z * dg.z * inv_rows);
            v.w = g.w * is.w * (dy.w - db.w * inv_rows - xh.w * dg.w * inv_rows);
            store_planes4(out + i, plane, npl, v);
        }
        return;
    }
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        const float is = invstd[c], g = gamma[c];
        const float xh = (z[i] - mean[c]) * is;
        float dy = da[i];
        if (act == 2 && fmaf(xh, g, beta[c]) <= 0.f) dy = 0.f;
        const float v = g * is * (dy - dbeta[c] * inv_rows - xh * dgamma[c] * inv_rows);
        store_planes(out + i, plane, npl, v);
    }
}

// ------------------------------------------------------------------------------------------------ Adam
// tf.train.AdamOptimizer: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMA; theta -= lr_t*m/(sqrt(v)+eps).
// step_ptr holds t (already incremented for this update) so that a captured CUDA graph can be replayed.
__global__ void adam_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g, long long n,
                            float lr, float b1, float b2, float eps, const int* __restrict__ step_ptr) {
    const double t = static_cast<double>(*step_ptr);
    const float lr_t = static_cast<float>(static_cast<double>(lr) * sqrt(1.0 - pow(static_cast<double>(b2), t)) /
                                          (1.0 - pow(static_cast<double>(b1), t)));
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float gi = g[i];
        const float mi = b1 * m[i] + (1.0f - b1) * gi;
        const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        w[i] -= lr_t * mi / (sqrtf(vi) + eps);
    }
}
__global__ void incr_step_kernel(int* step) { *step += 1; }
// sets flag[0] = 1 if any of the n values is NaN (the reference's per-step host assert, graph_func.py:856, kept on device)
__global__ void nan_flag_kernel(const float* __restrict__ x, int n, int* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && isnan(x[i])) atomicExch(flag, 1);
}

// ------------------------------------------------------------------------------------------------ launchers
static const int kBS = 256;
static inline int grid_for(long long n) {
    long long g = (n + kBS - 1) / kBS;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}
#define MG_CHECK_LAUNCH() (cudaGetLastError() == cudaSuccess ? 0 : -4)

int l_nchw_to_nhwc(const float* src, bf16_t* dst, long long plane, int npl, int fmt, int N, int C, int H, int W, int Cp, cudaStream_t st) {
    nchw_to_nhwc_kernel<<<grid_for(static_cast<long long>(N) * H * W * Cp), kBS, 0, st>>>(src, dst, plane, npl, fmt, N, C, H, W, Cp);
    return MG_CHECK_LAUNCH();
}
int l_nhwc_to_nchw(const bf16_t* src, long long plane, int npl, int fmt, float* dst, int N, int C, int H, int W, int Cp, cudaStream_t st) {
    nhwc_to_nchw_kernel<<<grid_for(static_cast<long long>(N) * C * H * W), kBS, 0, st>>>(src, plane, npl, fmt, dst, N, C, H, W, Cp);
    return MG_CHECK_LAUNCH();
}
int l_to_planes(const float* x, bf16_t* dst, long long plane, int npl, int fmt, long long n, cudaStream_t st) {
    to_planes_kernel<<<grid_for(n), kBS, 0, st>>>(x, dst, plane, npl, fmt, n);
    return MG_CHECK_LAUNCH();
}
int l_from_planes(const bf16_t* src, long long plane, int npl, int fmt, float* out, long long n, cudaStream_t st) {
    from_planes_kernel<<<grid_for(n), kBS, 0, st>>>(src, plane, npl, fmt, out, n);
    return MG_CHECK_LAUNCH();
}
int l_convert_planes(const bf16_t* src, long long sp, int snpl, int sfmt, bf16_t* dst, long long dp, int dnpl, int dfmt, long long n,
                     cudaStream_t st) {
    convert_planes_kernel<<<grid_for(n / 4), kBS, 0, st>>>(src, sp, snpl, sfmt, dst, dp, dnpl, dfmt, n);
    return MG_CHECK_LAUNCH();
}
int l_colsum_planes(const bf16_t* x, long long plane, int npl, int rows, int C, float* out, cudaStream_t st) {
    colsum_planes_kernel<<<nblocks(C, 32), dim3(32, 32), 0, st>>>(x, plane, npl, rows, C, out);
    return MG_CHECK_LAUNCH();
}
int l_pack_weights(const PackParams& p, cudaStream_t st) {
    pack_weights_kernel<<<grid_for(static_cast<long long>(p.rows_pad) * p.kpad * p.classes / 4), kBS, 0, st>>>(p);
    return MG_CHECK_LAUNCH();
}
int l_permute_features(const float* src, float* dst, int n, int C, int HW, int inverse, cudaStream_t st) {
    permute_features_kernel<<<nblocks(n, kBS), kBS, 0, st>>>(src, dst, n, C, HW, inverse);
    return MG_CHECK_LAUNCH();
}
int l_reduce_tiles(const float* partials, int T, int C, float scale, float* out, cudaStream_t st) {
    reduce_tiles_kernel<<<nblocks(C, 32), dim3(32, 32), 0, st>>>(partials, T, C, scale, out);
    return MG_CHECK_LAUNCH();
}
int l_colsum_small(const float* x, int rows, int C, float* out, cudaStream_t st) {
    colsum_small_kernel<<<nblocks(C, 128), 128, 0, st>>>(x, rows, C, out);
    return MG_CHECK_LAUNCH();
}
int wgrad_reduce_blocks(long long total) {
    long long g = (total + 256 * 2 - 1) / (256 * 2);
    if (g > 2048) g = 2048;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}
int l_wgrad_reduce(const WredParams& p, cudaStream_t st) {
    wgrad_reduce_kernel<<<wgrad_reduce_blocks(static_cast<long long>(p.R) * p.NC), 256, 0, st>>>(p);
    return MG_CHECK_LAUNCH();
}
int l_wgrad_reduce_batched(const WredParams* jobs, const int* start, int njobs, int total_blocks, cudaStream_t st) {
    wgrad_reduce_batched_kernel<<<total_blocks, 256, 0, st>>>(jobs, start, njobs);
    return MG_CHECK_LAUNCH();
}
int l_sn_grad_combine_batched(const void* jobs, int njobs, int blocks, cudaStream_t st) {
    sn_grad_combine_batched_kernel<<<dim3(blocks, njobs), kBS, 0, st>>>(static_cast<const SnCombineJob*>(jobs));
    return MG_CHECK_LAUNCH();
}
int l_sn_grad_combine(float* g, const float* s, const double* dots, int ndots, const float* sigma, float act_k, long long n,
                      cudaStream_t st) {
    sn_grad_combine_kernel<<<grid_for(n), kBS, 0, st>>>(g, s, dots, ndots, sigma, act_k, n);
    return MG_CHECK_LAUNCH();
}
int l_scale_by_sigma(float* g, const float* sigma, float act_k, long long n, cudaStream_t st) {
    scale_by_sigma_kernel<<<grid_for(n), kBS, 0, st>>>(g, sigma, act_k, n);
    return MG_CHECK_LAUNCH();
}
int l_sn_normalize(const float* v, long long n, float eps, float* sigma_out, bf16_t* out, long long plane, int npl, int fmt, cudaStream_t st) {
    sn_normalize_kernel<<<1, 1024, 0, st>>>(v, n, eps, sigma_out, out, plane, npl, fmt);
    return MG_CHECK_LAUNCH();
}
int l_bn_finalize(const float* psum, const float* psq, int T, int C, long long rows, float eps, float momentum, float* mean,
                  float* invstd, float* mm, float* mv, int bessel, cudaStream_t st) {
    if (T >= 256)
        bn_finalize_kernel<8><<<nblocks(C, 8), dim3(8, 128), 0, st>>>(psum, psq, T, C, rows, eps, momentum, mean, invstd, mm, mv, bessel);
    else
        bn_finalize_kernel<32><<<nblocks(C, 32), dim3(32, 32), 0, st>>>(psum, psq, T, C, rows, eps, momentum, mean, invstd, mm, mv, bessel);
    return MG_CHECK_LAUNCH();
}
int l_bn_inference_stats(const float* mm, const float* mv, int C, float eps, float* mean, float* invstd, cudaStream_t st) {
    bn_inference_stats_kernel<<<nblocks(C, 128), 128, 0, st>>>(mm, mv, C, eps, mean, invstd);
    return MG_CHECK_LAUNCH();
}
int l_bn_apply(const float* z, const float* mean, const float* invstd, const float* gamma, const float* beta, int C, long long total,
               int act, bf16_t* out, long long plane, int npl, int fmt, int* sat_flag, cudaStream_t st) {
    bn_apply_kernel<<<grid_for(total / 4), kBS, 0, st>>>(z, mean, invstd, gamma, beta, C, total, act, out, plane, npl, fmt, sat_flag);
    return MG_CHECK_LAUNCH();
}
int l_bn_bwd_reduce(const float* da, const float* z, const float* mean, const float* invstd, const float* gamma, const float* beta,
                    int C, long long rows, int rows_per_block, int act, float* psum, float* psumx, cudaStream_t st) {
    const int blocks = static_cast<int>((rows + rows_per_block - 1) / rows_per_block);
    if ((C & 3) == 0 && C <= 1024 && 256 % (C >> 2) == 0)
        bn_bwd_reduce_vec_kernel<<<blocks, 256, 0, st>>>(da, z, mean, invstd, gamma, beta, C, rows, rows_per_block, act, psum, psumx);
    else
        bn_bwd_reduce_kernel<<<blocks, 256, 0, st>>>(da, z, mean, invstd, gamma, beta, C, rows, rows_per_block, act, psum, psumx);
    return MG_CHECK_LAUNCH();
}
int l_bn_bwd_apply(const float* da, const float* z, const float* mean, const float* invstd, const float* gamma, const float* beta,
                   const float* dbeta, const float* dgamma, int C, long long rows, int act, bf16_t* out, long long plane, int npl,
                   cudaStream_t st) {
    bn_bwd_apply_kernel<<<grid_for((C & 3) == 0 ? rows * C / 4 : rows * C), kBS, 0, st>>>(da, z, mean, invstd, gamma, beta, dbeta, dgamma, C, rows, act, out,
                                                                                        plane, npl);
    return MG_CHECK_LAUNCH();
}
int l_adam(float* w, float* m, float* v, const float* g, long long n, float lr, float b1, float b2, float eps, const int* step,
           cudaStream_t st) {
    adam_kernel<<<grid_for(n), kBS, 0, st>>>(w, m, v, g, n, lr, b1, b2, eps, step);
    return MG_CHECK_LAUNCH();
}
int l_refresh(const RefreshJob* jobs, int njobs, long long total_blocks, cudaStream_t st) {
    if (njobs <= 0 || njobs > 256 || total_blocks <= 0) return -1;
    refresh_kernel<<<static_cast<unsigned>(total_blocks), kBS, 0, st>>>(jobs, njobs);
    return MG_CHECK_LAUNCH();
}
long long dense_small_workspace(int rows, int K, int N) {
    const long long slices = (K + kDsSlice - 1) / kDsSlice;
    return slices * rows * N * static_cast<long long>(sizeof(float));
}
int l_dense_small_fwd(const bf16_t* a, long long a_plane, int npl, int a_fmt, int rows, int K, const bf16_t* wt, long long w_plane, int w_fmt,
                      int kpad, int N, float alpha_k, const float* sigma, const float* bias, float* out, int ldo, float* workspace,
                      cudaStream_t st) {
    const int slices = (K + kDsSlice - 1) / kDsSlice;
    const int rpb = 128 / N;
    const dim3 grid((rows + rpb - 1) / rpb, slices, 1);
    if (N == 16) dense_small_fwd_kernel<16><<<grid, 256, 0, st>>>(a, a_plane, npl, a_fmt, rows, K, wt, w_plane, w_fmt, kpad, workspace);
    else if (N == 8) dense_small_fwd_kernel<8><<<grid, 256, 0, st>>>(a, a_plane, npl, a_fmt, rows, K, wt, w_plane, w_fmt, kpad, workspace);
    else if (N == 32) dense_small_fwd_kernel<32><<<grid, 256, 0, st>>>(a, a_plane, npl, a_fmt, rows, K, wt, w_plane, w_fmt, kpad, workspace);
    else return -1;
    dense_small_finish_kernel<<<nblocks(static_cast<long long>(rows) * N, 256), 256, 0, st>>>(workspace, slices, rows, N, alpha_k, sigma, bias, out, ldo);
    return MG_CHECK_LAUNCH();
}
// ------------------------------------------------------------------------------------------------ code sampling
// tf.random_normal([batch, code_size]) inside the training graph (reference my_sngan.py:122-124): N(0, 1) codes drawn ON THE
// DEVICE.  Philox-4x32-10 (Salmon et al., SC'11; the generator behind tf.random_normal and curand) keyed by the run seed,
// counter = (element quadruple index, 0, draw counter lo, draw counter hi) with the draw counter read from device memory, so
// a captured CUDA graph produces fresh codes on every replay; Box-Muller turns the four 32-bit words into four normals.
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
// u in (0, 1]: (x + 1) * 2^-32 computed in fp32 from the top 24 bits so that it never rounds to 0
__device__ __forceinline__ float u01(uint32_t x) { return (static_cast<float>(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }
__global__ void sample_normal_kernel(float* __restrict__ out, long long n, unsigned long long seed, const unsigned long long* __restrict__ draw,
                                     uint32_t* __restrict__ raw) {
    const long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;      // quadruple index
    if (q * 4 >= n) return;
    const unsigned long long d = draw ? *draw : 0ull;
    uint32_t c[4] = {static_cast<uint32_t>(q), static_cast<uint32_t>(q >> 32), static_cast<uint32_t>(d), static_cast<uint32_t>(d >> 32)};
    philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    float z[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float r = sqrtf(-2.0f * logf(u01(c[2 * h])));
        float s, co;
        sincospif(2.0f * u01(c[2 * h + 1]), &s, &co);
        z[2 * h] = r * co;
        z[2 * h + 1] = r * s;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (q * 4 + k < n) {
            out[q * 4 + k] = z[k];
            if (raw) raw[q * 4 + k] = c[k];
        }
}
__global__ void incr_u64_kernel(unsigned long long* c) { *c += 1ull; }
int l_sample_normal(float* out, long long n, unsigned long long seed, const unsigned long long* draw, uint32_t* raw, cudaStream_t st) {
    sample_normal_kernel<<<nblocks((n + 3) / 4, 256), 256, 0, st>>>(out, n, seed, draw, raw);
    return MG_CHECK_LAUNCH();
}
int l_incr_u64(unsigned long long* c, cudaStream_t st) {
    incr_u64_kernel<<<1, 1, 0, st>>>(c);
    return MG_CHECK_LAUNCH();
}
__global__ void losses_from_sums_kernel(const float* __restrict__ sums, float c0, float c1, float c2, float* __restrict__ losses) {
    losses[0] = sums[0] + sums[2] - 2.0f * sums[1];
    losses[1] = c0 * sums[3] + c1 * sums[4] + c2 * sums[5];
}
int l_losses_from_sums(const float* sums, float c0, float c1, float c2, float* losses, cudaStream_t st) {
    losses_from_sums_kernel<<<1, 1, 0, st>>>(sums, c0, c1, c2, losses);
    return MG_CHECK_LAUNCH();
}
int l_incr_step(int* step, cudaStream_t st) {
    incr_step_kernel<<<1, 1, 0, st>>>(step);
    return MG_CHECK_LAUNCH();
}
int l_nan_flag(const float* x, int n, int* flag, cudaStream_t st) {
    nan_flag_kernel<<<nblocks(n, 128), 128, 0, st>>>(x, n, flag);
    return MG_CHECK_LAUNCH();
}

}  // namespace mg
