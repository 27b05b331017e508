// Weight-gradient GEMM on the sm_100a tensor cores (contraction over pixels, split over blockIdx.z).
//
//   W[r][(t, c)] = sum_p P[p][r] * G[g(p, t)][c]
//
// P is a plain [pixels][Cp] activation / gradient matrix, G an NHWC activation gathered tap by tap, both as bf16 planes
// (fp16 forward activations are re-split by convert_planes first: one MMA cannot mix fp16 with bf16).
// With NHWC storage the contracted dimension (pixels) is the strided one, so BOTH operands are MN-major: a shared-
// memory "chunk" is 32 pixel rows x 128 bytes (64 consecutive channels of one pixel per row, 128-byte swizzle); P
// chunks are staged by TMA (one box {64 channels, 32 pixels, planes} per chunk), G chunks by 16-byte cp.async with the
// same swizzle, and the tcgen05.mma instruction descriptor carries the two transpose bits.  NPASS == 3 multiplies the
// plane pairs {00, 01, 10} (weight gradients are linear in both operands: ~2^-17 products suffice), NPASS == 1 is the
// plain bf16 mode.  Every split writes its own fp32 partial tile; the reduction over splits is fused into the
// gradient-finalise kernel.
//
// GTMA: when a 32-pixel k-step is a whole number of image rows (W in {4, 8, 16, 32}) and the gathered operand has whole
// 64-channel chunks, the G chunks are staged by TMA too -- one 5-D box {64 channels, W, rows, images, planes} per
// (column chunk, k-step), strided and zero-filled by the tensor map -- so no thread touches operand bytes.
// PAIR (GTMA only, result rows a multiple of 256): two CTAs of a cluster form one 256 x BN tile with cta_group::2 MMAs;
// each CTA stages its own 128 result rows of P but only HALF of the G columns, which cuts the L2 -> SM operand traffic
// per MMA cycle by a third (the kernel is bound by that traffic: 64 B/clk/SM at the full MMA rate without pairing).
//
// Replaces the filter gradients TF derives for tf.nn.conv2d / conv2d_transpose / matmul
// (DeepLearning/my_sngan.py:301-304) and the d(sigma)/dW term of SpectralNorm (GeneralTools/math_func.py:661-672).
#include "conv_gemm.cuh"
#include "tc_common.cuh"
#include <stdio.h>

namespace mg {

int make_tmap_planes(CUtensorMap* m, const uint16_t* base, long long rows, long long cols, long long row_stride_elems,
                     long long plane_stride_elems, int planes, int box_cols, int box_rows, int box_planes, int swizzle);
int make_tmap_act(CUtensorMap* m, const uint16_t* base, long long plane_stride_elems, int npl, int C, int W, int H, int N, int bw,
                  int bh, int bn, int sx, int sy);

static constexpr int kWBM = 128;
static constexpr int kWPix = 32;   // pixels (K) per pipeline stage: 32 KB stages at BN = 128 / two planes, two CTAs per SM
static constexpr int kWProducers = 128;
static constexpr int kWThreads = 192;
static constexpr unsigned long long kWWatchdogNs = 4000000000ull;

__device__ __forceinline__ unsigned long long w_gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void w_mbar_wait(uint64_t* bar, uint32_t parity, unsigned int* err, unsigned code) {
    if (mbar_try_wait(bar, parity)) return;
    unsigned long long t0 = w_gtime_ns();
    while (!mbar_try_wait(bar, parity)) {
        if (w_gtime_ns() - t0 > kWWatchdogNs) {
            if (err) atomicExch(err, code);
            printf("mmdgan: wgrad mbarrier watchdog (code %u) block (%d,%d,%d) thread %d\n", code, blockIdx.x, blockIdx.y,
                   blockIdx.z, threadIdx.x);
            __trap();
        }
    }
}

template <int BN, int NPASS, bool PAIR = false>
struct WgradCfg {
    static constexpr int NPL = (NPASS == 3) ? 2 : 1;
    static constexpr int CHUNK_BYTES = kWPix * 128;             // one 64-channel chunk of one plane: kWPix pixels x 128 B
    static constexpr int MCH = kWBM / 64;                       // row chunks (M = 128)
    static constexpr int NCH = (PAIR ? BN / 2 : BN) / 64;       // column chunks THIS CTA stages (a pair splits the columns)
    // stage layout: [A chunk][plane] then [B chunk][plane]; the planes of one chunk are adjacent (one TMA box)
    static constexpr int A_BYTES = MCH * NPL * CHUNK_BYTES;
    static constexpr int B_BYTES = NCH * NPL * CHUNK_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES_RAW = (96 * 1024) / STAGE_BYTES;   // two CTAs per SM
    static constexpr int STAGES = STAGES_RAW > 6 ? 6 : (STAGES_RAW < 2 ? 2 : STAGES_RAW);
    static constexpr int LAG = STAGES - 1 > 3 ? 3 : STAGES - 1;
    static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
    static constexpr int SMEM_BYTES = PIPE_BYTES + 1024 + 256;
    static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

template <int BN, int NPASS, bool GTMA, bool PAIR>
__device__ __forceinline__ void wgrad_body(const CUtensorMap& tmP0, const CUtensorMap& tmG0, const WgradParams& p) {
    static_assert(!PAIR || GTMA, "the CTA-pair form has no cp.async gather");
    using Cfg = WgradCfg<BN, NPASS, PAIR>;
    constexpr int NPL = Cfg::NPL;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int NCH = Cfg::NCH;
    constexpr int MCH = Cfg::MCH;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::PIPE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* accum_bar = bars + 2 * STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tile_m = blockIdx.x;      // PAIR: the cluster spans two consecutive row tiles
    const int tile_n = blockIdx.y;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const long long p_begin = static_cast<long long>(blockIdx.z) * p.p_per_split;
    long long p_end = p_begin + p.p_per_split;
    if (p_end > p.P) p_end = p.P;
    const int ksteps = p_end > p_begin ? static_cast<int>((p_end - p_begin + kWPix - 1) / kWPix) : 0;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmP0);
        if (GTMA) tma_prefetch_desc(&tmG0);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], GTMA ? 1 : kWProducers + 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 5) {
        if (PAIR) {
            tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
        } else {
            tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();      // programmatic dependent launch: the prologue above overlapped the previous kernel's tail
    pdl_launch_dependents();

    // chunk c, plane pl of the A / B operand of stage s
    auto stage_a = [&](int s, int c, int pl) -> uint8_t* { return smem + s * Cfg::STAGE_BYTES + (c * NPL + pl) * Cfg::CHUNK_BYTES; };
    auto stage_b = [&](int s, int c, int pl) -> uint8_t* {
        return smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES + (c * NPL + pl) * Cfg::CHUNK_BYTES;
    };

    if (warp < 4) {
      if (!GTMA) {
        // ======================= G producers (gather, MN-major) =======================
        const int t = threadIdx.x;
        const int chunk = t & 7;   // 16-byte unit (8 channels) inside the 128-byte row
        const int rbase = t >> 3;  // pixel rows rbase + 16*i of the k-step
        const int upt = p.Cs >> 3; // 16-byte units per tap
        const int ntaps = p.TH * p.TW;
        int ta[NCH], tb[NCH], tcq[NCH];
        bool tok[NCH];
#pragma unroll
        for (int q = 0; q < NCH; ++q) {
            const int u = (tile_n * NCH + q) * 8 + chunk;
            const int tap = u / upt;
            tcq[q] = u - tap * upt;
            ta[q] = tap / p.TW;
            tb[q] = tap - ta[q] * p.TW;
            tok[q] = tap < ntaps;
        }
        const int HgWg = p.Hg * p.Wg;
        for (int j = 0; j < ksteps; ++j) {
            const int s = j % STAGES;
            const uint32_t ph = (j / STAGES) & 1;
            w_mbar_wait(&empty_bar[s], ph ^ 1, p.err, 11);
            const uint32_t b0 = smem_u32(stage_b(s, 0, 0));
#pragma unroll
            for (int i = 0; i < kWPix / 16; ++i) {
                const int r = rbase + 16 * i;
                const long long pp = p_begin + static_cast<long long>(j) * kWPix + r;
                const bool pok = pp < p_end;
                int n = 0, y = 0, x = 0;
                if (pok) {
                    n = static_cast<int>(pp / HgWg);
                    const int rem = static_cast<int>(pp - static_cast<long long>(n) * HgWg);
                    y = rem / p.Wg;
                    x = rem - y * p.Wg;
                }
                const int by = y * p.sy + p.oy;
                const int bx = x * p.sx + p.ox;
                // SWIZZLE_128B: 16-byte unit index ^= row index inside the 8-row atom
                const uint32_t drow = b0 + static_cast<uint32_t>(r * 128 + ((chunk ^ (r & 7)) << 4));
#pragma unroll
                for (int q = 0; q < NCH; ++q) {
                    const int yy = by + ta[q];
                    const int xx = bx + tb[q];
                    const bool ok = pok && tok[q] && yy >= 0 && yy < p.Hs && xx >= 0 && xx < p.Ws;
                    const long long off =
                        ok ? (static_cast<long long>(n * p.Hs * p.Ws + yy * p.Ws + xx) * p.Cs + tcq[q] * 8) : 0;
#pragma unroll
                    for (int pl = 0; pl < NPL; ++pl)
                        cp_async16(drow + (q * NPL + pl) * Cfg::CHUNK_BYTES, p.g + pl * p.g_plane + off, ok ? 16u : 0u);
                }
            }
            cp_async_mbar_arrive_noinc(&full_bar[s]);   // asynchronous publication, see conv_gemm.cu
        }
      }
    } else if (warp == 4) {
        // ======================= P producer (TMA, MN-major); with GTMA the G chunks as well =======================
        if (lane == 0) {
            // GTMA: column chunk q of this CTA is (tap, 64-channel chunk); a k-step is `rows` image rows of `nb` images
            int gch[NCH], gx[NCH], gy[NCH];
            int img = 0, y = 0, ystep = 0, nstep = 0;
            if (GTMA) {
                const int ntaps = p.TH * p.TW;
#pragma unroll
                for (int q = 0; q < NCH; ++q) {
                    const int col = tile_n * BN + (PAIR ? static_cast<int>(rank) * (BN / 2) : 0) + q * 64;
                    const int tap = col / p.Cs;
                    const int ta = tap / p.TW;
                    gch[q] = col - tap * p.Cs;
                    gx[q] = tap - ta * p.TW + p.ox;
                    gy[q] = tap < ntaps ? ta + p.oy : (1 << 20);      // a tap past the filter: the whole box is out of bounds (zeros)
                }
                const int rows = kWPix / p.Wg;
                ystep = rows < p.Hg ? rows : 0;                       // rows >= Hg: whole images per k-step
                nstep = rows < p.Hg ? 0 : rows / p.Hg;
                const int hw = p.Hg * p.Wg;
                img = static_cast<int>(p_begin / hw);
                y = static_cast<int>(p_begin - static_cast<long long>(img) * hw) / p.Wg;
            }
            for (int j = 0; j < ksteps; ++j) {
                const int s = j % STAGES;
                const uint32_t ph = (j / STAGES) & 1;
                w_mbar_wait(&empty_bar[s], ph ^ 1, p.err, 12);
                const int prow = static_cast<int>(p_begin + static_cast<long long>(j) * kWPix);
                if (PAIR) {
                    // both CTAs load their shares; all bytes are credited to the leader's barrier
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::STAGE_BYTES);
#pragma unroll
                    for (int mc = 0; mc < MCH; ++mc)
                        tma_load_3d_pair(smem_u32(stage_a(s, mc, 0)), &tmP0, &full_bar[s], tile_m * kWBM + mc * 64, prow, 0);
#pragma unroll
                    for (int q = 0; q < NCH; ++q)
                        tma_load_5d_pair(smem_u32(stage_b(s, q, 0)), &tmG0, &full_bar[s], gch[q], gx[q], y * p.sy + gy[q], img, 0);
                } else {
                    mbar_arrive_expect_tx(&full_bar[s], GTMA ? Cfg::STAGE_BYTES : Cfg::A_BYTES);
#pragma unroll
                    for (int mc = 0; mc < MCH; ++mc)   // one box = {64 channels, kWPix pixels, NPL planes}
                        tma_load_3d(smem_u32(stage_a(s, mc, 0)), &tmP0, &full_bar[s], tile_m * kWBM + mc * 64, prow, 0);
                    if (GTMA) {
#pragma unroll
                        for (int q = 0; q < NCH; ++q)
                            tma_load_5d(smem_u32(stage_b(s, q, 0)), &tmG0, &full_bar[s], gch[q], gx[q], y * p.sy + gy[q], img, 0);
                    }
                }
                if (GTMA) {
                    y += ystep;
                    if (y >= p.Hg) { y = 0; ++img; }
                    img += nstep;
                }
            }
        }
    } else if (!PAIR || rank == 0) {
        // ======================= MMA issuer (the pair's leader issues for both CTAs) =======================
        const uint32_t idesc = idesc_f16(PAIR ? 2 * kWBM : kWBM, BN, 1, 1, p.p_fmt, p.g_fmt);
        for (int j = 0; j < ksteps; ++j) {
            const int s = j % STAGES;
            const uint32_t ph = (j / STAGES) & 1;
            w_mbar_wait(&full_bar[s], ph, p.err, 13);
            fence_proxy_async_smem();
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int pass = 0; pass < NPASS; ++pass) {
                    const int pa = (pass == 1) ? 1 : 0;
                    const int pb = (pass == 2) ? 1 : 0;
                    const uint32_t abase = smem_u32(stage_a(s, 0, pa));
                    const uint32_t bbase = smem_u32(stage_b(s, 0, pb));
#pragma unroll
                    for (int kg = 0; kg < kWPix / 16; ++kg) {
                        // MN-major SWIZZLE_128B: one MMA covers K = 16 pixel rows = two 8-row atoms (SBO = 1024 bytes apart);
                        // consecutive 64-channel chunks of the same plane are LBO = NPL chunks apart
                        const uint64_t ad = smem_desc(abase + kg * 2048, NPL * Cfg::CHUNK_BYTES, 1024, 2u);
                        const uint64_t bd = smem_desc(bbase + kg * 2048, NPL * Cfg::CHUNK_BYTES, 1024, 2u);
                        const uint32_t acc = (j > 0 || pass > 0 || kg > 0) ? 1u : 0u;
                        if (PAIR) umma_bf16_pair(tmem_base, ad, bd, idesc, acc);
                        else umma_bf16(tmem_base, ad, bd, idesc, acc);
                    }
                }
                if (PAIR) umma_commit_pair(&empty_bar[s]); else umma_commit(&empty_bar[s]);
            }
            __syncwarp();
        }
        if (lane == 0) {
            if (PAIR) umma_commit_pair(accum_bar); else umma_commit(accum_bar);
        }
        __syncwarp();
    }

    // ======================= epilogue: partial tile -> global =======================
    if (warp < 4) {
        const int r_out = tile_m * kWBM + warp * 32 + lane;
        float* orow = p.out + (static_cast<long long>(blockIdx.z) * p.Cp + r_out) * p.Ncols;
        if (ksteps > 0) {
            w_mbar_wait(accum_bar, 0, p.err, 14);
            tc_fence_after();
        }
        float v[32];
#pragma unroll 1
        for (int cc = 0; cc < BN / 32; ++cc) {
            if (ksteps > 0) {
                tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + cc * 32, v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int q = 0; q < 32; ++q) v[q] = 0.f;
            }
            const int col0 = tile_n * BN + cc * 32;
            if (r_out < p.Cp) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int col = col0 + q * 4;
                    if (col < p.Ncols)
                        *reinterpret_cast<float4*>(orow + col) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                }
            }
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
        else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int BN, int NPASS, bool GTMA>
__global__ void __launch_bounds__(kWThreads, 2)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap tmP0, const __grid_constant__ CUtensorMap tmG0,
                  const __grid_constant__ WgradParams p) {
    wgrad_body<BN, NPASS, GTMA, false>(tmP0, tmG0, p);
}
template <int BN, int NPASS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWThreads, 2)
wgrad_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmP0, const __grid_constant__ CUtensorMap tmG0,
                       const __grid_constant__ WgradParams p) {
    wgrad_body<BN, NPASS, true, true>(tmP0, tmG0, p);
}

// a 32-pixel k-step is `hb` whole image rows of `nb` images, and the gathered operand has whole 64-channel chunks
static bool gtma_geometry(const WgradParams& p, int* hb, int* nb) {
    if (p.Cs % 64 != 0 || p.Wg <= 0 || kWPix % p.Wg != 0) return false;
    const int rows = kWPix / p.Wg;
    if (rows <= p.Hg) {
        if (p.Hg % rows != 0) return false;
        *hb = rows; *nb = 1;
    } else {
        if (rows % p.Hg != 0) return false;
        *hb = p.Hg; *nb = rows / p.Hg;
    }
    if (p.p_per_split % (static_cast<long long>(rows >= p.Hg ? p.Hg : rows) * p.Wg * (*nb)) != 0) return false;
    return p.Wg * p.sx <= 256 && *hb * p.sy <= 256;
}

template <int BN, int NPASS, bool GTMA, bool PAIR>
static int launch_wcfg(const WgradParams& p, const uint16_t* plain, long long plain_plane, int splits, int hb, int nb, cudaStream_t st) {
    using Cfg = WgradCfg<BN, NPASS, PAIR>;
    CUtensorMap t0, g0;
    if (make_tmap_planes(&t0, plain, p.P, p.Cp, p.Cp, plain_plane, Cfg::NPL, 64, kWPix, Cfg::NPL, 0)) return -4;
    if (GTMA) {
        if (make_tmap_act(&g0, p.g, p.g_plane, Cfg::NPL, p.Cs, p.Ws, p.Hs, p.Nimg, p.Wg * p.sx, hb * p.sy, nb, p.sx, p.sy)) return -4;
    } else {
        g0 = t0;
    }
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e;
        if constexpr (PAIR) e = cudaFuncSetAttribute(wgrad_gemm_pair_kernel<BN, NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        else e = cudaFuncSetAttribute(wgrad_gemm_kernel<BN, NPASS, GTMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return -4;
        attr_done = true;
    }
    dim3 grid((p.Cp + kWBM - 1) / kWBM, (p.Ncols + BN - 1) / BN, splits);
    cudaError_t le;
    if constexpr (PAIR) le = launch_pdl(wgrad_gemm_pair_kernel<BN, NPASS>, grid, dim3(kWThreads), Cfg::SMEM_BYTES, st, t0, g0, p);
    else le = launch_pdl(wgrad_gemm_kernel<BN, NPASS, GTMA>, grid, dim3(kWThreads), Cfg::SMEM_BYTES, st, t0, g0, p);
    return le == cudaSuccess && cudaGetLastError() == cudaSuccess ? 0 : -4;
}

int launch_wgrad_gemm(const WgradParams& p, const uint16_t* plain, long long plain_plane, int splits, int bn, int npass,
                      cudaStream_t st) {
    static int no_gtma = -1, no_pair = -1;
    if (no_gtma < 0) { const char* e = getenv("MMDGAN_WGRAD_NO_GTMA"); no_gtma = e ? atoi(e) : 0; }
    if (no_pair < 0) { const char* e = getenv("MMDGAN_WGRAD_NO_PAIR"); no_pair = e ? atoi(e) : 0; }
    int hb = 0, nb = 0;
    const bool gtma = !no_gtma && gtma_geometry(p, &hb, &nb);
    const bool pair = gtma && !no_pair && bn >= 128 && p.Cp % (2 * kWBM) == 0;
#define MG_CASE(B, N) \
    if (bn == B && npass == N) \
        return gtma ? launch_wcfg<B, N, true, false>(p, plain, plain_plane, splits, hb, nb, st) \
                    : launch_wcfg<B, N, false, false>(p, plain, plain_plane, splits, hb, nb, st);
#define MG_PAIR(B, N) \
    if (bn == B && npass == N && pair) return launch_wcfg<B, N, true, true>(p, plain, plain_plane, splits, hb, nb, st);
    MG_PAIR(128, 3) MG_PAIR(256, 3) MG_PAIR(128, 1) MG_PAIR(256, 1)
    MG_CASE(64, 3) MG_CASE(128, 3) MG_CASE(256, 3)
    MG_CASE(64, 1) MG_CASE(128, 1) MG_CASE(256, 1)
#undef MG_CASE
#undef MG_PAIR
    return -1;
}

}  // namespace mg
