// Fused pairwise squared-distance -> Gaussian kernel(s) -> repulsive-MMD losses + score gradients, one launch.
//
// Replaces the ~40 TF ops of GANLoss._repulsive_mmd_g_ / _repulsive_mmd_g_bounded_ / _mmd_g_ / _mmd_g_bound_
// (GeneralTools/math_func.py:2160-2193, 2505-2550): get_squared_dist (767-858, Gram trick with clamp at 0),
// matrix_mean_wo_diagonal (1048-1069, the i == j entry is dropped from ALL three matrices), mmd_g (1288-1352),
// mmd_g_bounded (1356-1431), mixture_mmd_g (1435-1473), mmd_t / mixture_mmd_t (1087-1184: t-distribution kernels
// k = exp(-alpha log(d / (alpha beta) + 1))) and the backward pass TF derives from them.
//
// One warp owns one score row (a "row task": first the local generated rows, then the local real rows) and sweeps
// every column of the (gathered) generated and real score matrices, which are staged tile by tile, transposed, in
// shared memory with float4 coalesced loads.  The gradient of a row needs only that row's kernel values
//   grad_i = (sum_j w_ij) x_i - sum_j w_ij y_j
// so no B x B matrix is ever written.  Lanes split the columns, warp shuffles reduce, the last block to finish
// reduces the per-block kernel sums in a fixed order (deterministic) and writes the two losses.
//
// Row-block (multi-GPU) form: the local rows are rows [row0, row0 + b) of the global index space; "diagonal" means
// equal GLOBAL index and the normalisation uses the global batch.
#include "conv_gemm.cuh"
#include <cuda_runtime.h>
#include <stdint.h>

namespace mg {


static constexpr int kTJ = 128;
static constexpr int kWarps = 8;
static constexpr float kLog2e = 1.4426950408889634f;

template <int D>
__global__ void __launch_bounds__(kWarps * 32) mmd_fused_kernel(const MmdParams p) {
    extern __shared__ float sm[];
    constexpr int PITCH = kTJ + 1;
    float* gT = sm;                    // [D][PITCH]
    float* rT = gT + D * PITCH;        // [D][PITCH]
    float* gN = rT + D * PITCH;        // [kTJ]
    float* rN = gN + kTJ;              // [kTJ]
    float* wsum = rN + kTJ;            // [kWarps][6]
    __shared__ bool is_last;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int task = blockIdx.x * kWarps + warp;
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
            const float4 v = *reinterpret_cast<const float4*>(xr + k);
            xi[k] = v.x; xi[k + 1] = v.y; xi[k + 2] = v.z; xi[k + 3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < D; ++k) ni = fmaf(xi[k], xi[k], ni);
    } else {
#pragma unroll
        for (int k = 0; k < D; ++k) xi[k] = 0.f;
    }

    float aG = 0.f, aD = 0.f;
    float tG[D], tD[D];
#pragma unroll
    for (int k = 0; k < D; ++k) { tG[k] = 0.f; tD[k] = 0.f; }
    float s_same_u = 0.f, s_same_b = 0.f, s_gr_u = 0.f, s_gr_b = 0.f;

    // the matrix a G-row meets in the "same" sweep is gg (index 0), an R-row meets rr (index 2)
    const int same_idx = is_g ? 0 : 2;
    const float cD_same = p.cD[same_idx], cD_gr = p.cD[1];
    const int bm_same = p.bmode[same_idx], bm_gr = p.bmode[1];
    const float bv_same = p.bval[same_idx], bv_gr = p.bval[1];

    for (int j0 = 0; j0 < p.Bg; j0 += kTJ) {
        __syncthreads();
        // ---- stage kTJ rows of both matrices, transposed (float4 coalesced global reads)
        constexpr int QPR = D / 4;
        for (int e = threadIdx.x; e < 2 * kTJ * QPR; e += blockDim.x) {
            const int which = e / (kTJ * QPR);
            const int f = e - which * kTJ * QPR;
            const int r = f / QPR, q = f - r * QPR;
            const int j = j0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < p.Bg) v = *reinterpret_cast<const float4*>((which ? p.real_all : p.gen_all) + static_cast<long long>(j) * D + q * 4);
            float* T = which ? rT : gT;
            T[(q * 4 + 0) * PITCH + r] = v.x;
            T[(q * 4 + 1) * PITCH + r] = v.y;
            T[(q * 4 + 2) * PITCH + r] = v.z;
            T[(q * 4 + 3) * PITCH + r] = v.w;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < 2 * kTJ; e += blockDim.x) {
            const int which = e / kTJ, r = e - which * kTJ;
            const float* T = which ? rT : gT;
            float n = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) n = fmaf(T[k * PITCH + r], T[k * PITCH + r], n);
            (which ? rN : gN)[r] = n;
        }
        __syncthreads();
        if (!active) continue;

#pragma unroll 1
        for (int sweep = 0; sweep < 2; ++sweep) {
            // sweep 0: same-set matrix (gg for a G row, rr for an R row); sweep 1: the cross matrix gr
            const bool same = sweep == 0;
            const float* T = (same == is_g) ? gT : rT;
            const float* N = (same == is_g) ? gN : rN;
            const float cDm = same ? cD_same : cD_gr;
            const int bm = same ? bm_same : bm_gr;
            const float bv = same ? bv_same : bv_gr;
            // d(mean)/d(row): the row appears in two ordered pairs of a same-set matrix, once in the cross matrix
            const float mult = same ? 2.0f * c : c;
            const float cG = same ? 1.0f : -2.0f;   // loss_gen = e_gg + e_rr - 2 e_gr
#pragma unroll 1
            for (int jj = lane; jj < kTJ; jj += 32) {
                const int j = j0 + jj;
                if (j >= p.Bg || j == gi) continue;
                float dot = 0.f;
#pragma unroll
                for (int k = 0; k < D; ++k) dot = fmaf(xi[k], T[k * PITCH + jj], dot);
                const float raw = ni - 2.0f * dot + N[jj];
                const float dist = fmaxf(raw, 0.0f);
                const float m0 = raw >= 0.0f ? 1.0f : 0.0f;       // tf.maximum passes the gradient on ties
                float distb = dist, mb = 1.0f;
                if (bm == 1) { distb = fmaxf(dist, bv); mb = dist >= bv ? 1.0f : 0.0f; }
                else if (bm == 2) { distb = fminf(dist, bv); mb = dist <= bv ? 1.0f : 0.0f; }
                float ku = 0.f, kb = 0.f, dku = 0.f, dkb = 0.f;   // kernel sums and -(dK/dd) sums over sigma
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
                // d dist / d x_i = 2 (x_i - y_j);  dK/dd = -dk
                const float wG = -cG * mult * 2.0f * dku * m0;
                const float wD = -cDm * mult * 2.0f * dkb * m0 * mb;
                aG += wG; aD += wD;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const float y = T[k * PITCH + jj];
                    tG[k] = fmaf(wG, y, tG[k]);
                    tD[k] = fmaf(wD, y, tD[k]);
                }
                if (same) { s_same_u += ku; s_same_b += kb; }
                else if (is_g) { s_gr_u += ku; s_gr_b += kb; }
            }
        }
    }

    // ---- warp reductions
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        aG += __shfl_xor_sync(0xffffffffu, aG, o);
        aD += __shfl_xor_sync(0xffffffffu, aD, o);
        s_same_u += __shfl_xor_sync(0xffffffffu, s_same_u, o);
        s_same_b += __shfl_xor_sync(0xffffffffu, s_same_b, o);
        s_gr_u += __shfl_xor_sync(0xffffffffu, s_gr_u, o);
        s_gr_b += __shfl_xor_sync(0xffffffffu, s_gr_b, o);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            tG[k] += __shfl_xor_sync(0xffffffffu, tG[k], o);
            tD[k] += __shfl_xor_sync(0xffffffffu, tD[k], o);
        }
    }
    if (active && lane == 0) {
        float* og = is_g ? p.dLg_dgen : p.dLg_dreal;
        float* od = is_g ? p.dLd_dgen : p.dLd_dreal;
#pragma unroll
        for (int k = 0; k < D; k += 4) {
            if (og) *reinterpret_cast<float4*>(og + static_cast<long long>(li) * D + k) =
                make_float4(aG * xi[k] - tG[k], aG * xi[k + 1] - tG[k + 1], aG * xi[k + 2] - tG[k + 2], aG * xi[k + 3] - tG[k + 3]);
            if (od) *reinterpret_cast<float4*>(od + static_cast<long long>(li) * D + k) =
                make_float4(aD * xi[k] - tD[k], aD * xi[k + 1] - tD[k + 1], aD * xi[k + 2] - tD[k + 2], aD * xi[k + 3] - tD[k + 3]);
        }
    }
    if (lane == 0) {
        float* w = wsum + warp * 6;
        // order: gg_u, gr_u, rr_u, gg_b, gr_b, rr_b
        w[0] = (active && is_g) ? s_same_u : 0.f;
        w[1] = (active && is_g) ? s_gr_u : 0.f;
        w[2] = (active && !is_g) ? s_same_u : 0.f;
        w[3] = (active && is_g) ? s_same_b : 0.f;
        w[4] = (active && is_g) ? s_gr_b : 0.f;
        w[5] = (active && !is_g) ? s_same_b : 0.f;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float s = 0.f;
        for (int w = 0; w < kWarps; ++w) s += wsum[w * 6 + threadIdx.x];
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
        if (threadIdx.x < 6) {
            float s = 0.f;
            for (unsigned int blk = 0; blk < gridDim.x; ++blk) s += __ldcg(p.partials + blk * 6 + threadIdx.x);
            wsum[threadIdx.x] = s * c;
            p.sums[threadIdx.x] = s * c;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            p.losses[0] = wsum[0] + wsum[2] - 2.0f * wsum[1];
            p.losses[1] = p.cD[0] * wsum[3] + p.cD[1] * wsum[4] + p.cD[2] * wsum[5];
            *p.counter = 0u;
        }
    }
}

template <int D>
static int launch_mmd_d(const MmdParams& p, cudaStream_t st) {
    const size_t smem = (2 * D * (kTJ + 1) + 2 * kTJ + kWarps * 6) * sizeof(float);
    static bool attr_done = false;
    if (!attr_done && smem > 48 * 1024) {
        if (cudaFuncSetAttribute(mmd_fused_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
            return -4;
        attr_done = true;
    }
    const int grid = (2 * p.b + kWarps - 1) / kWarps;
    mmd_fused_kernel<D><<<grid, kWarps * 32, smem, st>>>(p);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

int mmd_grid_blocks(int b) { return (2 * b + kWarps - 1) / kWarps; }

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
