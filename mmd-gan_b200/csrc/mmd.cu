// Fused pairwise squared-distance -> Gaussian kernel(s) -> repulsive-MMD losses + score gradients, one launch.
//
// Replaces the ~40 TF ops of GANLoss._repulsive_mmd_g_ / _repulsive_mmd_g_bounded_ / _mmd_g_ / _mmd_g_bound_
// (GeneralTools/math_func.py:2160-2193, 2505-2550): get_squared_dist (767-858, Gram trick with clamp at 0),
// matrix_mean_wo_diagonal (1048-1069, the i == j entry is dropped from ALL three matrices), mmd_g (1288-1352),
// mmd_g_bounded (1356-1431), mixture_mmd_g (1435-1473), mmd_t / mixture_mmd_t (1087-1184: t-distribution kernels
// k = exp(-alpha log(d / (alpha beta) + 1))) and the backward pass TF derives from them.
//
// A "row task" is one score row (first the local generated rows, then the local real rows) against every column of the
// (gathered) generated and real score matrices.  The gradient of a row needs only that row's kernel values
//   grad_i = (sum_j w_ij) x_i - sum_j w_ij y_j
// so no B x B matrix is ever written.
//
// Shape of the launch (round 2; the first version gave one warp a whole row and took ~14-22 us at B = 256, 4x a launch):
//   * the grid is sized to ONE wave: rows-per-block = ceil(2b / #SMs) rounded up to a power of two (4 at b = 256: 128 blocks of
//     256 threads), and the 8 / rpb warps of a row split its columns, so every SM works from the first cycle;
//   * the column tile (256 rows of both matrices) is staged ROW-major with a pitch of D + 4 floats: one thread loads one
//     score row (four 16-byte loads), keeps its squared norm, and a lane later reads "its" column as D / 4 conflict-free
//     LDS.128 -- used for the dot product AND the gradient accumulation (the transposed layout of the first version cost
//     2 D scalar shared-memory loads per pair);
//   * the 2 D gradient accumulators are reduced across a warp with the recursive-halving lane transpose (2 D - 1 shuffles
//     instead of 2 D x 5), the six scalars with butterflies, the warps of a row through shared memory;
//   * the last block sums the per-block kernel sums with 32 lanes per sum (fixed order: deterministic) and writes the losses.
// Kernel family / bandwidth count are template parameters for the hot case (one Gaussian bandwidth: rep, rmb, mgb).
//
// Row-block (multi-GPU) form: the local rows are rows [row0, row0 + b) of the global index space; "diagonal" means
// equal GLOBAL index and the normalisation uses the global batch.
#include "conv_gemm.cuh"
#include <cuda_runtime.h>
#include <stdint.h>

namespace mg {

static constexpr int kTJ = 256;          // columns (score rows of the gathered matrices) per shared-memory tile
static constexpr int kWarps = 8;
static constexpr int kMmdThreads = kWarps * 32;
static constexpr float kLog2e = 1.4426950408889634f;

// Sum over the lanes of a warp of V per-lane values: afterwards lane l holds the totals of values [l * V / 32, (l + 1) * V / 32)
// in v[0 .. V / 32) (V >= 32, a power of two).
template <int V>
__device__ __forceinline__ void warp_transpose_sum(float (&v)[V], int lane) {
#pragma unroll
    for (int off = 16, n = V / 2; off >= 1; off >>= 1, n >>= 1) {
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const float keep = hi ? v[i + n] : v[i];
            const float send = hi ? v[i] : v[i + n];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
}

// SIMPLE: Gaussian family with one bandwidth (rep / rmb / mgb)
template <int D, bool SIMPLE>
__global__ void __launch_bounds__(kMmdThreads) mmd_fused_kernel(const MmdParams p, const int rpb) {
    extern __shared__ __align__(16) float sm[];
    constexpr int PITCH = D + 4;                  // floats: 16-byte aligned rows, conflict-free LDS.128 for consecutive rows
    constexpr int V = 2 * D;                      // gradient accumulators per thread (generator-loss and discriminator-loss sums)
    constexpr int NR = V + 8;                     // reduced values per row task: V sums, aG, aD, four kernel sums (+2 pad)
    float* gS = sm;                               // [kTJ][PITCH]
    float* rS = gS + kTJ * PITCH;                 // [kTJ][PITCH]
    float* gN = rS + kTJ * PITCH;                 // [kTJ]
    float* rN = gN + kTJ;                         // [kTJ]
    float* red = rN + kTJ;                        // [kWarps][NR]
    float* tot = red + kWarps * NR;               // [rpb <= 8][NR]
    __shared__ bool is_last;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpr = kWarps / rpb;                 // warps per row task
    const int task = blockIdx.x * rpb + warp / wpr;
    const int sub = warp % wpr;                   // this warp's share of the columns
    const bool active = task < 2 * p.b;
    const bool is_g = task < p.b;
    const int li = is_g ? task : task - p.b;
    const int gi = p.row0 + li;
    const float c = 1.0f / (static_cast<float>(p.Bg) * (static_cast<float>(p.Bg) - 1.0f));

    float xi[D];
    float ni = 0.f;
    if (active) {
        const float* xr = (is_g ? p.gen_loc : p.real_loc) + static_cast<long long>(li) * D;
#pragma unroll
        for (int k = 0; k < D; k += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(xr + k));
            xi[k] = v.x; xi[k + 1] = v.y; xi[k + 2] = v.z; xi[k + 3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < D; ++k) ni = fmaf(xi[k], xi[k], ni);
    } else {
#pragma unroll
        for (int k = 0; k < D; ++k) xi[k] = 0.f;
    }

    float aG = 0.f, aD = 0.f;
    float tv[V];                                  // [0, D): sum w_G y ; [D, 2D): sum w_D y
#pragma unroll
    for (int k = 0; k < V; ++k) tv[k] = 0.f;
    float s_same_u = 0.f, s_same_b = 0.f, s_gr_u = 0.f, s_gr_b = 0.f;

    // the matrix a G-row meets in the "same" sweep is gg (index 0), an R-row meets rr (index 2)
    const int same_idx = is_g ? 0 : 2;
    const float cD_same = p.cD[same_idx], cD_gr = p.cD[1];
    const int bm_same = p.bmode[same_idx], bm_gr = p.bmode[1];
    const float bv_same = p.bval[same_idx], bv_gr = p.bval[1];
    const float cs0 = p.c_s[0];

    for (int j0 = 0; j0 < p.Bg; j0 += kTJ) {
        if (j0 > 0) __syncthreads();
        // ---- stage kTJ rows of both matrices: one thread, one row (D / 4 coalesced 16-byte loads), norm on the fly
        for (int e = threadIdx.x; e < 2 * kTJ; e += kMmdThreads) {
            const int which = e / kTJ, r = e - which * kTJ;
            const int j = j0 + r;
            float* dst = (which ? rS : gS) + r * PITCH;
            float n = 0.f;
            if (j < p.Bg) {
                const float* src = (which ? p.real_all : p.gen_all) + static_cast<long long>(j) * D;
#pragma unroll
                for (int k = 0; k < D; k += 4) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(src + k));
                    *reinterpret_cast<float4*>(dst + k) = v;
                    n = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, n))));
                }
            }
            (which ? rN : gN)[r] = n;
        }
        __syncthreads();
        if (!active) continue;

#pragma unroll 1
        for (int sweep = 0; sweep < 2; ++sweep) {
            // sweep 0: same-set matrix (gg for a G row, rr for an R row); sweep 1: the cross matrix gr
            const bool same = sweep == 0;
            const float* S = (same == is_g) ? gS : rS;
            const float* N = (same == is_g) ? gN : rN;
            const float cDm = same ? cD_same : cD_gr;
            const int bm = same ? bm_same : bm_gr;
            const float bv = same ? bv_same : bv_gr;
            // d(mean)/d(row): the row appears in two ordered pairs of a same-set matrix, once in the cross matrix
            const float mult = same ? 2.0f * c : c;
            const float cG = same ? 1.0f : -2.0f;   // loss_gen = e_gg + e_rr - 2 e_gr
            const int jend = min(kTJ, p.Bg - j0);
#pragma unroll 2
            for (int jj = sub * 32 + lane; jj < jend; jj += wpr * 32) {
                if (j0 + jj == gi) continue;
                float y[D];
#pragma unroll
                for (int k = 0; k < D; k += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(S + jj * PITCH + k);
                    y[k] = v.x; y[k + 1] = v.y; y[k + 2] = v.z; y[k + 3] = v.w;
                }
                float dot = 0.f;
#pragma unroll
                for (int k = 0; k < D; ++k) dot = fmaf(xi[k], y[k], dot);
                const float raw = ni - 2.0f * dot + N[jj];
                const float dist = fmaxf(raw, 0.0f);
                const float m0 = raw >= 0.0f ? 1.0f : 0.0f;       // tf.maximum passes the gradient on ties
                float distb = dist, mb = 1.0f;
                if (bm == 1) { distb = fmaxf(dist, bv); mb = dist >= bv ? 1.0f : 0.0f; }
                else if (bm == 2) { distb = fminf(dist, bv); mb = dist <= bv ? 1.0f : 0.0f; }
                float ku = 0.f, kb = 0.f, dku = 0.f, dkb = 0.f;   // kernel sums and -(dK/dd) sums over sigma
                if (SIMPLE) {
                    ku = exp2f(-dist * cs0 * kLog2e);
                    kb = (bm == 0) ? ku : exp2f(-distb * cs0 * kLog2e);
                    dku = cs0 * ku;
                    dkb = cs0 * kb;
                } else {
                    for (int s = 0; s < p.n_sigma; ++s) {
                        const float cs = p.c_s[s];
                        if (p.family == 0) {                 // Gaussian: k = exp(-cs d), -dk/dd = cs k
                            const float eu = exp2f(-dist * cs * kLog2e);
                            const float eb = (bm == 0) ? eu : exp2f(-distb * cs * kLog2e);
                            ku += eu; kb += eb;
                            dku = fmaf(cs, eu, dku);
                            dkb = fmaf(cs, eb, dkb);
                        } else {                             // t: k = u^-alpha with u = 1 + d / (alpha beta), -dk/dd = k / (beta u)
                            const float uu = fmaf(dist, p.c_t[s], 1.0f), ub = fmaf(distb, p.c_t[s], 1.0f);
                            const float eu = exp2f(-cs * log2f(uu));
                            const float eb = (bm == 0) ? eu : exp2f(-cs * log2f(ub));
                            ku += eu; kb += eb;
                            dku = fmaf(eu, p.inv_beta / uu, dku);
                            dkb = fmaf(eb, p.inv_beta / ub, dkb);
                        }
                    }
                }
                // d dist / d x_i = 2 (x_i - y_j);  dK/dd = -dk
                const float wG = -cG * mult * 2.0f * dku * m0;
                const float wD = -cDm * mult * 2.0f * dkb * m0 * mb;
                aG += wG; aD += wD;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    tv[k] = fmaf(wG, y[k], tv[k]);
                    tv[D + k] = fmaf(wD, y[k], tv[D + k]);
                }
                if (same) { s_same_u += ku; s_same_b += kb; }
                else if (is_g) { s_gr_u += ku; s_gr_b += kb; }
            }
        }
    }

    // ---- warp reductions: lane transpose for the 2 D gradient sums, butterflies for the six scalars
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        aG += __shfl_xor_sync(0xffffffffu, aG, o);
        aD += __shfl_xor_sync(0xffffffffu, aD, o);
        s_same_u += __shfl_xor_sync(0xffffffffu, s_same_u, o);
        s_same_b += __shfl_xor_sync(0xffffffffu, s_same_b, o);
        s_gr_u += __shfl_xor_sync(0xffffffffu, s_gr_u, o);
        s_gr_b += __shfl_xor_sync(0xffffffffu, s_gr_b, o);
    }
    float* rw = red + warp * NR;
    if (V >= 32) {
        warp_transpose_sum<V>(tv, lane);
        constexpr int PER = V / 32 > 0 ? V / 32 : 1;
#pragma unroll
        for (int i = 0; i < PER; ++i) rw[lane * PER + i] = tv[i];
    } else {
#pragma unroll
        for (int k = 0; k < V; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tv[k] += __shfl_xor_sync(0xffffffffu, tv[k], o);
        }
        if (lane == 0)
#pragma unroll
            for (int k = 0; k < V; ++k) rw[k] = tv[k];
    }
    if (lane == 0) {
        rw[V] = aG; rw[V + 1] = aD;
        // order: gg_u, gr_u, rr_u (unbounded) / gg_b, gr_b, rr_b are assembled per row below
        rw[V + 2] = s_same_u; rw[V + 3] = s_same_b; rw[V + 4] = s_gr_u; rw[V + 5] = s_gr_b;
    }
    __syncthreads();
    // ---- the warps of a row task -> one total per value
    for (int e = threadIdx.x; e < rpb * NR; e += kMmdThreads) {
        const int r = e / NR, k = e - r * NR;
        float s = 0.f;
        for (int w = 0; w < wpr; ++w) s += red[(r * wpr + w) * NR + k];
        tot[e] = s;
    }
    __syncthreads();
    // ---- gradients of the block's rows: grad = a x_i - sum w y
    for (int e = threadIdx.x; e < rpb * V; e += kMmdThreads) {
        const int r = e / V, k = e - r * V;
        const int tk = blockIdx.x * rpb + r;
        if (tk >= 2 * p.b) continue;
        const bool g_row = tk < p.b;
        const int row = g_row ? tk : tk - p.b;
        const bool dpart = k >= D;
        const int kk = dpart ? k - D : k;
        float* out = dpart ? (g_row ? p.dLd_dgen : p.dLd_dreal) : (g_row ? p.dLg_dgen : p.dLg_dreal);
        if (out == nullptr) continue;
        const float x = __ldg((g_row ? p.gen_loc : p.real_loc) + static_cast<long long>(row) * D + kk);
        const float a = tot[r * NR + V + (dpart ? 1 : 0)];
        out[static_cast<long long>(row) * D + kk] = a * x - tot[r * NR + k];
    }
    // ---- kernel sums of the block (order: gg_u, gr_u, rr_u, gg_b, gr_b, rr_b)
    if (threadIdx.x < 6) {
        float s = 0.f;
        for (int r = 0; r < rpb; ++r) {
            const int tk = blockIdx.x * rpb + r;
            if (tk >= 2 * p.b) continue;
            const bool g_row = tk < p.b;
            const float* t = tot + r * NR + V + 2;      // same_u, same_b, gr_u, gr_b
            float v = 0.f;
            switch (threadIdx.x) {
                case 0: v = g_row ? t[0] : 0.f; break;
                case 1: v = g_row ? t[2] : 0.f; break;
                case 2: v = g_row ? 0.f : t[0]; break;
                case 3: v = g_row ? t[1] : 0.f; break;
                case 4: v = g_row ? t[3] : 0.f; break;
                default: v = g_row ? 0.f : t[1]; break;
            }
            s += v;
        }
        p.partials[blockIdx.x * 6 + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(p.counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        // six sums over gridDim.x blocks: warp w < 6 owns sum w, its lanes stride over the blocks (fixed order)
        if (warp < 6) {
            float s = 0.f;
            for (unsigned int blk = lane; blk < gridDim.x; blk += 32) s += __ldcg(p.partials + blk * 6 + warp);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) {
                red[warp] = s * c;
                p.sums[warp] = s * c;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            p.losses[0] = red[0] + red[2] - 2.0f * red[1];
            p.losses[1] = p.cD[0] * red[3] + p.cD[1] * red[4] + p.cD[2] * red[5];
            *p.counter = 0u;
        }
    }
}

// rows per block: one wave on the device (<= #SMs blocks) as long as that needs <= 8 rows per block, a power of two so that the
// eight warps divide evenly
static int mmd_rows_per_block(int b) {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    int rpb = 1;
    while (rpb < 8 && (2 * b + rpb - 1) / rpb > sms) rpb *= 2;
    return rpb;
}

template <int D>
static int launch_mmd_d(const MmdParams& p, cudaStream_t st) {
    const size_t smem = (2 * kTJ * (D + 4) + 2 * kTJ + kWarps * (2 * D + 8) + 8 * (2 * D + 8)) * sizeof(float);
    static bool attr_done = false;
    if (!attr_done && smem > 48 * 1024) {
        if (cudaFuncSetAttribute(mmd_fused_kernel<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess ||
            cudaFuncSetAttribute(mmd_fused_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
            return -4;
        attr_done = true;
    }
    const int rpb = mmd_rows_per_block(p.b);
    const int grid = (2 * p.b + rpb - 1) / rpb;
    if (p.family == 0 && p.n_sigma == 1) mmd_fused_kernel<D, true><<<grid, kMmdThreads, smem, st>>>(p, rpb);
    else mmd_fused_kernel<D, false><<<grid, kMmdThreads, smem, st>>>(p, rpb);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// upper bound of the grid over every rows-per-block choice (workspace sizing: 6 floats per block)
int mmd_grid_blocks(int b) { return 2 * b; }

int launch_mmd(const MmdParams& p, cudaStream_t st) {
    switch (p.d) {
        case 4: return launch_mmd_d<4>(p, st);
        case 8: return launch_mmd_d<8>(p, st);
        case 16: return launch_mmd_d<16>(p, st);
        case 32: return launch_mmd_d<32>(p, st);
        case 64: return launch_mmd_d<64>(p, st);
        default: return -1;
    }
}

}  // namespace mg
