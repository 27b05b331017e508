// Gather-GEMM on the sm_100a tensor cores: the conv / transposed-conv / dense forward and input-gradient passes.
//
//   D[m][n] = epilogue( alpha * sum_k A[m][k] * Bw[n][k] )
//
// A (activations, NHWC, 16-bit planes) is staged tap by tap into 64-byte-swizzled K-major shared-memory tiles -- by TMA
// as one 5-D box {32 channels, W, rows, images, planes} per (channel chunk, tap) when a tile is a whole number of image
// rows, else by four producer warps with 16-byte cp.async (zero fill implements SAME padding and tile tails); Bw
// (weights, K-major [planes][rows][Kpad]) is staged by one 3-D TMA box; one elected thread issues tcgen05.mma kind::f16
// (fp16 or bf16 in -- both operands the same type --, fp32 accumulate in TMEM).
// An fp32 value is carried as the sum of 16-bit planes (tc_common.cuh).  NPASS == 3 multiplies the plane pairs {00, 01,
// 10}: with two fp16 planes per operand (22 bits) that is fp32-grade -- the forward passes, whose errors the MMD loss
// amplifies --, with two bf16 planes (16 bits) it is ~2^-17 -- input gradients, linear in the operand.  NPASS == 6
// multiplies {00, 01, 10, 02, 20, 11} of three bf16 planes (fp32-grade without fp16's range limit: the spectral-norm
// adjoint, and the forward passes when MMDGAN_F16_FORWARD=0); NPASS == 1 is the plain bf16 speed mode.  The producer
// warps turn into the epilogue: TMEM -> registers -> shared staging -> coalesced stores with bias / activation /
// activation derivative (signs prefetched as bit masks during the main loop) / plane split and per-tile column sums
// (bias gradients, batch-norm statistics) fused in.
//
// Replaces the tf.matmul / tf.nn.conv2d / tf.nn.conv2d_transpose call sites of the reference
// (GeneralTools/layer_func.py:909-928) and their gradients (DeepLearning/my_sngan.py:301-304).
#include "conv_gemm.cuh"
#include "tc_common.cuh"
#include <stdio.h>
#include <stdlib.h>

namespace mg {

static constexpr int kBM = 128;
static constexpr int kSmemMax = 227 * 1024;   // opt-in dynamic shared memory per block on sm_100
// K per pipeline stage = 64 16-bit elements = one 128-byte swizzle row (SWIZZLE_128B).  With 64-byte rows (SWIZZLE_64B) the
// tensor core's operand reads of a K = 16 slice (32 bytes of each of 8 rows) hit every shared-memory bank twice: measured
// ~134 cycles per 256 x 128 x 16 MMA against the 64-cycle floor (role profiling, MMDGAN_PROF=1) -- the N <= 128 layers ran
// at half rate.  In a 128-byte-swizzled atom the same slice spreads over all 32 banks.
static constexpr int kBK = kGemmBK;    // 64
static constexpr int kRowBytes = 128;
static constexpr int kEpiThreads = 128;       // one epilogue group: warps 0-3 (group 0) / warps 6-9 (group 1); TMEM lane quarter = warp % 4
static constexpr int kProducerThreads = 128;  // warps 10-13: cp.async gather producers (ATMA = false only)
static constexpr int kThreadsTma = 320;       // epilogue group 0, TMA, MMA, epilogue group 1
static constexpr int kThreadsGather = 448;    // + four gather warps
static constexpr int kMaxStages = 8;
static constexpr unsigned long long kWatchdogNs = 4000000000ull;

__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// bounded wait: a pipeline bug must fail loudly instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity, unsigned int* err, unsigned code) {
    if (mbar_try_wait(bar, parity)) return;
    unsigned long long t0 = gtime_ns();
    while (!mbar_try_wait(bar, parity)) {
        if (gtime_ns() - t0 > kWatchdogNs) {
            if (err) atomicExch(err, code);
            printf("mmdgan: mbarrier watchdog (code %u) block %d thread %d\n", code, blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}

// ---- epilogue building blocks
// The epilogue is ROW-PER-THREAD and register-only: thread (warp % 4) * 32 + lane owns one accumulator row (= one output
// pixel), pulls 32 columns at a time out of TMEM, applies alpha / bias / activation / activation derivative, converts and
// writes its 64 (planes) or 128 (fp32) contiguous bytes per chunk straight to global memory.  No staging tile, no named
// barrier per column group, no shared memory except the 4 x BN column-sum partials.  (The staged form needed ~5 barriers
// and a serial 4-thread reduction per 32 columns: with MMAs switched off the N <= 128 launches ran no faster -- the
// epilogue, at ~17 k cycles per 128 x 64 tile, was the bottleneck, not the tensor pipe.)
__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == 1) return v > 0.f ? v : 0.1f * v;
    if (act == 2) return v > 0.f ? v : 0.f;
    if (act == 3) return tanhf(v);
    return v;
}
__device__ __forceinline__ float act_grad_from_output(float a, int mode) {
    if (mode == 1) return a > 0.f ? 1.f : 0.1f;
    if (mode == 2) return a > 0.f ? 1.f : 0.f;
    if (mode == 3) return 1.f - a * a;
    return 1.f;
}
// 256-bit global stores / loads (sm_100: STG.256 / LDG.256): a thread's 32 bytes fill one whole sector per instruction
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&w)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]),
                 "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}
__device__ __forceinline__ void epi_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// Sum over the 32 lanes of a warp of CW = 32 per-lane values, column by column: afterwards lane j holds the total of
// column j.  Recursive halving: 16 + 8 + 4 + 2 + 1 = 31 shuffles instead of 32 x 5.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float keep = hi ? v[i + off] : v[i];
            const float send = hi ? v[i] : v[i + off];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}
// CW = 16: lanes j and j + 16 both end with the total of column j
__device__ __forceinline__ float warp_colsum16(float (&v)[32], int lane) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float keep = hi ? v[i + off] : v[i];
            const float send = hi ? v[i] : v[i + off];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

struct EpiRow {
    const ConvGemmParams* p;
    long long prow;        // destination pixel of this thread's row (-1: beyond M)
    long long arow;        // row of the aux activation (3B-wrap applied)
    float alpha, negslope, auxslope;
    float* part;           // [4 warps][2][BN] column-sum partials of this epilogue group
    int warp4, lane, bn;
};
// One chunk of CW accumulator columns of one row.  OUT 0 = two fp16 planes of 16 x value (FMT_F16A), 1 = two bf16 planes,
// 2 = raw fp32, 3 = anything else (store_val per element); AUX 0 none, 1 sign bits of plane 0 (lrelu' / relu'), 2 tanh'
// from the value of the aux activation; GENERIC adds the run-time activation switch (tanh).
template <int CW, int AUX, bool COLSUM, int OUT, bool GENERIC>
__device__ __forceinline__ void epi_chunk(const EpiRow& c, float (&v)[32], const uint4 (&araw)[4], int col, int lcol) {
    const ConvGemmParams& p = *c.p;
    const int nvalid = p.Ncols - col;        // > 0, a multiple of 4
    const bool row_ok = c.prow >= 0;
    // ---- alpha, bias, activation
#pragma unroll
    for (int q = 0; q < CW / 4; ++q) {
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias && q * 4 < nvalid) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col + q * 4));
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float x = fmaf(v[q * 4 + e], c.alpha, bb[e]);
            if (GENERIC) x = apply_act(x, p.act);
            else x = fmaf(c.negslope, fminf(x, 0.f), fmaxf(x, 0.f));
            v[q * 4 + e] = x;
        }
    }
    // ---- activation derivative
    if (AUX == 1) {
        const uint32_t w[16] = {araw[0].x, araw[0].y, araw[0].z, araw[0].w, araw[1].x, araw[1].y, araw[1].z, araw[1].w,
                                araw[2].x, araw[2].y, araw[2].z, araw[2].w, araw[3].x, araw[3].y, araw[3].z, araw[3].w};
#pragma unroll
        for (int j = 0; j < CW; ++j) {
            const uint32_t h = (j & 1) ? (w[j >> 1] >> 16) : (w[j >> 1] & 0xFFFFu);
            v[j] *= (h - 1u < 0x7FFFu) ? 1.f : c.auxslope;      // 16-bit float > 0: sign clear, magnitude non-zero
        }
    } else if (AUX == 2) {
        if (row_ok) {
#pragma unroll
            for (int q = 0; q < CW / 4; ++q)
                if (q * 4 < nvalid) {
                    const float4 a4 = load_vals4(p.aux, p.aux_plane, p.aux_npl, p.aux_fmt, c.arow * p.Cd + col + q * 4);
                    v[q * 4] *= act_grad_from_output(a4.x, 3); v[q * 4 + 1] *= act_grad_from_output(a4.y, 3);
                    v[q * 4 + 2] *= act_grad_from_output(a4.z, 3); v[q * 4 + 3] *= act_grad_from_output(a4.w, 3);
                }
        }
    }
    // ---- store this row's CW columns
    if (row_ok) {
        const long long off = c.prow * p.Cd + col;
        const bool wide = (p.Cd & 15) == 0 && (col & 15) == 0;      // 32-byte aligned rows: 256-bit stores
        if (OUT == 2) {
            float* d = static_cast<float*>(p.dst) + off;
#pragma unroll
            for (int q8 = 0; q8 < CW / 8; ++q8) {
                if (wide && q8 * 8 + 8 <= nvalid) {
                    uint32_t w[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) w[k] = __float_as_uint(v[q8 * 8 + k]);
                    st_global_v8(d + q8 * 8, w);
                } else {
#pragma unroll
                    for (int q = 2 * q8; q < 2 * q8 + 2; ++q)
                        if (q * 4 < nvalid) *reinterpret_cast<float4*>(d + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                }
            }
        } else if (OUT == 3) {
#pragma unroll
            for (int q = 0; q < CW / 4; ++q)
                if (q * 4 < nvalid) {
                    const float4 x = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                    note_saturation4(p.sat_flag, p.dst_fmt, x);
                    store_vals4(static_cast<bf16_t*>(p.dst) + off + q * 4, p.dst_plane, p.dst_npl, p.dst_fmt, x);
                }
        } else {
            // two 16-bit planes, 8 elements (16 bytes) per store
            bf16_t* d0 = static_cast<bf16_t*>(p.dst) + off;
            bf16_t* d1 = d0 + p.dst_plane;
            float amax = 0.f;
#pragma unroll
            for (int o16 = 0; o16 < (CW + 15) / 16; ++o16) {
                uint32_t w0[8], w1[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int j = o16 * 16 + 2 * k;
                    const float xa = j < CW ? v[j < 32 ? j : 0] : 0.f, xb = j + 1 < CW ? v[j + 1 < 32 ? j + 1 : 0] : 0.f;
                    uint16_t a0, a1, b0, b1;
                    if (OUT == 0) {
                        f16_split2(xa * 16.f, a0, a1);
                        f16_split2(xb * 16.f, b0, b1);
                        amax = fmaxf(amax, fmaxf(fabsf(xa), fabsf(xb)));
                    } else {
                        a0 = f2bf(xa); a1 = f2bf(xa - bf2f(a0));
                        b0 = f2bf(xb); b1 = f2bf(xb - bf2f(b0));
                    }
                    w0[k] = pack16(a0, b0);
                    w1[k] = pack16(a1, b1);
                }
                const int e0 = o16 * 16;       // first element of this 32-byte group
                if (wide && e0 + 16 <= nvalid && e0 + 16 <= CW) {
                    st_global_v8(d0 + e0, w0);
                    st_global_v8(d1 + e0, w1);
                } else {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int e = e0 + 8 * h;
                        if (e >= CW) continue;
                        if (e + 4 < nvalid) {
                            *reinterpret_cast<uint4*>(d0 + e) = make_uint4(w0[4 * h], w0[4 * h + 1], w0[4 * h + 2], w0[4 * h + 3]);
                            *reinterpret_cast<uint4*>(d1 + e) = make_uint4(w1[4 * h], w1[4 * h + 1], w1[4 * h + 2], w1[4 * h + 3]);
                        } else if (e < nvalid) {
                            *reinterpret_cast<uint2*>(d0 + e) = make_uint2(w0[4 * h], w0[4 * h + 1]);
                            *reinterpret_cast<uint2*>(d1 + e) = make_uint2(w1[4 * h], w1[4 * h + 1]);
                        }
                    }
                }
            }
            if (OUT == 0 && p.sat_flag != nullptr && amax * 16.f > 65504.f) atomicExch(p.sat_flag, 1);
        }
    }
    // ---- per-tile column sums of the written values (bias gradients, batch-norm statistics)
    if (COLSUM) {
        const bool counted = row_ok && c.prow < p.colsum_rows;
        float sq[32];
#pragma unroll
        for (int j = 0; j < CW; ++j) {
            if (!counted) v[j] = 0.f;
            sq[j] = v[j] * v[j];
        }
        const float s1 = CW == 32 ? warp_colsum32(v, c.lane) : warp_colsum16(v, c.lane);
        float s2 = 0.f;
        if (p.colsumsq) s2 = CW == 32 ? warp_colsum32(sq, c.lane) : warp_colsum16(sq, c.lane);
        if (c.lane < CW) {
            c.part[(c.warp4 * 2 + 0) * c.bn + lcol + c.lane] = s1;
            c.part[(c.warp4 * 2 + 1) * c.bn + lcol + c.lane] = s2;
        }
    }
}


// profiling experiments (MMDGAN_PROF=1): cycles a role spends waiting, per CTA
__device__ __forceinline__ void mbar_wait_prof(uint64_t* bar, uint32_t parity, unsigned int* err, unsigned code, long long* acc) {
    if (acc == nullptr) { mbar_wait_wd(bar, parity, err, code); return; }
    const long long t0 = clock64();
    mbar_wait_wd(bar, parity, err, code);
    *acc += clock64() - t0;
}

// PERSISTENT kernel: one CTA (or CTA pair) per SM loops over output tiles; the fp32 accumulator is DOUBLE-BUFFERED in TMEM
// (2 x BN columns), so the epilogue warps drain tile i while the MMA warp already accumulates tile i + 1, the operand
// pipeline never drains between tiles, and barrier init / TMEM allocation / tensor-map fetch happen once per SM instead of
// once per tile.  The epilogue staging tile has its own shared memory (it no longer aliases the pipeline stages).
//
// PAIR = true: two CTAs of a cluster (one TPC) share one 256 x BN tile through tcgen05 cta_group::2 -- each CTA gathers
// its own 128 rows of A and stages HALF of the weight tile, so the L2 -> SM bytes per MMA drop by a quarter (BN = 128) to
// a half (BN = 256) and each SM's tensor core reads only half of B from its own shared memory.
template <int BN, int NPASS, bool PAIR>
struct GemmCfg {
    static constexpr int NPL = (NPASS == 6) ? 3 : ((NPASS == 3) ? 2 : 1);   // operand planes staged per k-block
    // A tcgen05.mma with both operands in shared memory takes >= ~128 cycles whatever its N (measured 115-133 cycles per MMA
    // for N = 64, 128 and 256 alike: role profiling, MMDGAN_PROF=1) -- only N = 256 runs at the full rate.  So for BN <= 128
    // the two weight planes are CONCATENATED along N: one MMA per activation plane, A_a x [B_0 | B_1] with N = 2 BN, writes
    // the two partial products into adjacent column blocks of the accumulator, and the epilogue adds the blocks.  Two MMAs of
    // N = 2 BN per K = 16 slice instead of three of N = BN (and all four plane-pair products instead of three).
    static constexpr bool CAT = NPASS == 3 && BN <= 128;
    static constexpr int NMMA = CAT ? NPL * BN : BN;      // N of one MMA
    // weight rows staged by this CTA per plane, and planes: a CAT pair gives each CTA ONE whole plane (CTA r = plane r = columns
    // [r BN, (r + 1) BN) of the concatenated operand); otherwise a pair splits every plane in halves
    static constexpr int BROWS = (PAIR && !CAT) ? BN / 2 : BN;
    static constexpr int BPL = (PAIR && CAT) ? 1 : NPL;   // weight planes staged by this CTA
    static constexpr int A_BYTES = kBM * kRowBytes;       // per plane
    static constexpr int B_BYTES = BROWS * kRowBytes;     // per plane
    static constexpr int STAGE_BYTES = NPL * A_BYTES + BPL * B_BYTES;
    static constexpr int NACC = CAT ? NPL : 1;            // column blocks of width BN the epilogue adds up
    static constexpr int BUF_COLS = (NACC * BN) < 32 ? 32 : NACC * BN;   // TMEM columns of one tile's accumulator
    static constexpr int TMEM_COLS = 2 * BUF_COLS;        // double-buffered: a power of two in [64, 512]
    // Every byte of shared memory that is not a pipeline stage costs operand bytes in flight (the L2 round trip is ~1.5 us; a
    // stage is consumed in 0.4 - 0.8 us at the full MMA rate): the register-only epilogue needs just the column-sum partials.
    static constexpr int EPI_BYTES = 4 * 2 * BN * 4 < 1024 ? 1024 : 4 * 2 * BN * 4;   // [4 warps][sum, sum of squares][BN], per epilogue group
    static constexpr int BAR_BYTES = 1024;
    static constexpr int FIXED_BYTES = 2 * EPI_BYTES + BAR_BYTES;
    static constexpr int STAGES_RAW = (kSmemMax - 1024 /*align slack*/ - FIXED_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > kMaxStages ? kMaxStages : (STAGES_RAW < 2 ? 2 : STAGES_RAW);
    static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
    static constexpr int SMEM_NEED = 1024 + PIPE_BYTES + FIXED_BYTES;
    // at least half of the SM's shared memory + 1 KB: ONE persistent CTA per SM (two of them must never queue for one SM's TMEM)
    static constexpr int SMEM_BYTES = SMEM_NEED > kSmemMax / 2 + 1024 ? SMEM_NEED : kSmemMax / 2 + 1024;
    static_assert(SMEM_BYTES <= kSmemMax, "shared-memory budget");
};

// ---- dynamic tile scheduler.  Tiles [0, nunits) are assigned statically (unit u starts with tile u); every further tile is a
// ticket drawn from a global counter by the scheduler thread of the unit (the TMA thread of the single CTA / of the pair's
// leader) and published to the other roles through a small ring in shared memory (pair: in both CTAs): tile_full[slot] says
// "ring[slot] holds the id of the unit's lt-th tile (-1: no more tiles)", tile_empty[slot] that every consumer has read it.
// A CTA that starts late (another stream's kernel still held its SM) simply draws fewer tickets: no tail imbalance.
static constexpr int kRing = 8;
template <bool PAIR>
__device__ __forceinline__ int ring_read(const int* ring, uint64_t* tile_full, uint64_t* tile_empty, int lt, unsigned int* err,
                                         uint32_t rank = 0) {
    // called by one lane; returns the tile id of this unit's lt-th tile.  Only the pair's PEER CTA needs cluster scope (its ring
    // copy is written from the leader, its "read" arrival goes to the leader's barrier); the leader's own consumers share a CTA
    // with the scheduler and use the plain CTA-scope wait / arrive on the same barriers.
    const int slot = lt % kRing;
    const uint32_t ph = (lt / kRing) & 1;
    const bool remote = PAIR && rank != 0;
    if (remote) {
        if (!mbar_try_wait_cluster(&tile_full[slot], ph)) {
            const unsigned long long t0 = gtime_ns();
            while (!mbar_try_wait_cluster(&tile_full[slot], ph))
                if (gtime_ns() - t0 > kWatchdogNs) {
                    printf("mmdgan: tile ring watchdog block %d thread %d\n", blockIdx.x, threadIdx.x);
                    __trap();
                }
        }
    } else {
        mbar_wait_wd(&tile_full[slot], ph, err, 9);
    }
    const int tile = *reinterpret_cast<const volatile int*>(ring + slot);
    if (remote) mbar_arrive_leader(&tile_empty[slot]); else mbar_arrive(&tile_empty[slot]);
    return tile;
}
// one lane reads, the warp gets the value
template <bool PAIR>
__device__ __forceinline__ int ring_read_warp(const int* ring, uint64_t* tile_full, uint64_t* tile_empty, int lt, unsigned int* err,
                                              uint32_t rank = 0) {
    int tile = 0;
    if ((threadIdx.x & 31) == 0) tile = ring_read<PAIR>(ring, tile_full, tile_empty, lt, err, rank);
    return __shfl_sync(0xffffffffu, tile, 0);
}


struct TileCoord {
    int tile_n, cls_idx, tile_m;
};
// linear tile index -> (N tile, class, M tile), N tile fastest: the CTAs that gather the same input rows (the other N tiles,
// the other output-parity classes) run at the same time on neighbouring SMs, so the re-reads hit L2 instead of HBM
template <bool PAIR>
__device__ __forceinline__ TileCoord decode_tile(int lin, const ConvGemmParams& p, uint32_t rank) {
    TileCoord c;
    c.tile_n = lin % p.tiles_n;
    lin /= p.tiles_n;
    c.cls_idx = lin % p.classes;
    lin /= p.classes;
    c.tile_m = PAIR ? lin * 2 + static_cast<int>(rank) : lin;
    return c;
}

// ATMA = true: the gathered operand is staged by TMA as well -- one 5-D box {32 channels, W, rows, images, planes} per
// (channel chunk, tap) lands exactly the 128 K-major rows of the tile, with out-of-bounds zero fill as SAME padding and
// the traversal stride as the convolution stride.  Used whenever a 128-row tile is a whole number of image rows;
// otherwise four producer warps gather with cp.async (ATMA = false).
template <int BN, int NPASS, bool PAIR, bool ATMA>
__device__ __forceinline__ void conv_gemm_body(const CUtensorMap& tmB0, const CUtensorMap& tmA0, const ConvGemmParams& p) {
    using Cfg = GemmCfg<BN, NPASS, PAIR>;
    constexpr int NPL = Cfg::NPL;
    constexpr int STAGES = Cfg::STAGES;
    constexpr bool RELAY = PAIR && !ATMA;         // the peer's cp.async arrivals must be forwarded to the leader

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // two epilogue groups of four warps alternate over the unit's tiles (group g drains accumulator g): every latency of the
    // epilogue (sign prefetch, tcgen05.ld, staging, stores) has two tile times to hide in
    const int egrp = warp >= 6 ? 1 : 0;
    float* part = reinterpret_cast<float*>(smem + Cfg::PIPE_BYTES + egrp * Cfg::EPI_BYTES);   // column-sum partials of this thread's group
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::PIPE_BYTES + 2 * Cfg::EPI_BYTES);
    uint64_t* full_bar = bars;                          // single: A arrivals + B bytes.  pair: this CTA's A arrivals only
    uint64_t* empty_bar = bars + kMaxStages;
    uint64_t* bfull_bar = bars + 2 * kMaxStages;        // pair, leader: weight bytes of BOTH CTAs
    uint64_t* peer_bar = bars + 3 * kMaxStages;         // pair, leader: "the peer's A rows have landed"
    uint64_t* accf_bar = bars + 4 * kMaxStages;         // [2] accumulator b is complete (MMA -> epilogue)
    uint64_t* acce_bar = bars + 4 * kMaxStages + 2;     // [2] accumulator b is drained (epilogue -> MMA; pair: leader's, both CTAs arrive)
    uint64_t* tile_full = bars + 4 * kMaxStages + 4;    // [kRing]
    uint64_t* tile_empty = bars + 4 * kMaxStages + 4 + kRing;   // [kRing] (pair: the leader's; consumers of both CTAs arrive)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * kMaxStages + 4 + 2 * kRing);
    int* ring = reinterpret_cast<int*>(tmem_slot + 2);          // [kRing]

    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int unit = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int nunits = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    const int total = p.tiles_n * p.classes * (PAIR ? p.tiles_m / 2 : p.tiles_m);
    const int ksteps = p.ksteps;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmB0);
        if (ATMA) tma_prefetch_desc(&tmA0);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], ATMA ? 1 : (PAIR ? kProducerThreads : kProducerThreads + 1));
            mbar_init(&empty_bar[s], 1);
            if (PAIR) {
                mbar_init(&bfull_bar[s], 1);
                mbar_init(&peer_bar[s], 1);
            }
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&accf_bar[b], 1);
            mbar_init(&acce_bar[b], PAIR ? 8 : 4);      // one arrival per epilogue warp (of both CTAs of a pair)
        }
        // consumers of a tile id: MMA warp + 8 epilogue warps (+ 4 gather warps); in the pair's peer CTA the TMA thread and the
        // relay warp read it too
        constexpr int kConsSelf = 1 + 8 + (ATMA ? 0 : 4);
        constexpr int kConsPeer = 1 + 8 + (ATMA ? 0 : 4 + 1);
        for (int r = 0; r < kRing; ++r) {
            mbar_init(&tile_full[r], 1);
            mbar_init(&tile_empty[r], PAIR ? kConsSelf + kConsPeer : kConsSelf);
        }
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
    // everything above touched only kernel parameters and on-chip state: from here on the previous kernel's results are read
    pdl_wait();
    pdl_launch_dependents();

    auto stage_a = [&](int s, int pl) -> uint8_t* { return smem + s * Cfg::STAGE_BYTES + pl * Cfg::A_BYTES; };
    auto stage_b = [&](int s, int pl) -> uint8_t* {
        return smem + s * Cfg::STAGE_BYTES + NPL * Cfg::A_BYTES + pl * Cfg::B_BYTES;      // pl < Cfg::BPL
    };

    if (warp >= 10) {
      if (!ATMA) {
        // ======================= A producers (gather) =======================
        const int t = threadIdx.x - 320;
        const int chunk = t & 7;   // 16-byte chunk of the 128-byte row
        const int rbase = t >> 3;  // rows rbase + 16*i
        // SWIZZLE_128B: 16-byte chunk index ^= bits [7,10) of the byte address = row & 7 (rows rbase + 16*i share it)
        const uint32_t swz_off = static_cast<uint32_t>((chunk ^ (rbase & 7)) << 4);
        const int HgWg = p.Hg * p.Wg;
        // K order: (channel chunk of CW = min(Cs, 64), tap, channel within chunk) -- taps innermost, so consecutive
        // k-steps re-read neighbouring pixels of the SAME 128-byte channel chunk
        const int cw4 = (p.Cs >= kBK ? kBK : p.Cs) >> 3;   // 16-byte units (8 bf16) per (chunk, tap) group
        const int ncc = p.Cs / (cw4 * 8);
        const int ntaps = p.TH * p.TW;
        int it = 0;
        for (int lt = 0;; ++lt) {
            const int tile = ring_read_warp<PAIR>(ring, tile_full, tile_empty, lt, p.err, rank);
            if (tile < 0) break;
            const TileCoord tc = decode_tile<PAIR>(tile, p, rank);
            const GemmClass cls = p.cls[tc.cls_idx];
            int by[8], bx[8], ib[8];  // by < -30000 marks an invalid row
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int m = tc.tile_m * kBM + rbase + 16 * i;
                if (m < p.M) {
                    int n = m / HgWg;
                    int rem = m - n * HgWg;
                    int y = rem / p.Wg;
                    int x = rem - y * p.Wg;
                    by[i] = y * p.sy + cls.oy;
                    bx[i] = x * p.sx + cls.ox;
                    ib[i] = n * p.Hs * p.Ws;
                } else {
                    by[i] = -1000000;
                    bx[i] = 0;
                    ib[i] = 0;
                }
            }
            for (int j = 0; j < ksteps; ++j, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait_wd(&empty_bar[s], ph ^ 1, p.err, 1);
                const int u = j * 8 + chunk;
                const int g = u / cw4;
                const int cc = g / ntaps;
                const int tap = g - cc * ntaps;
                const int cq = cc * cw4 + (u - g * cw4);
                const int a = tap / p.TW;
                const int b = tap - a * p.TW;
                const bool tap_ok = cc < ncc;
                const uint32_t a0 = smem_u32(stage_a(s, 0)) + swz_off;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int yy = by[i] + a;
                    const int xx = bx[i] + b;
                    const bool ok = tap_ok && yy >= 0 && yy < p.Hs && xx >= 0 && xx < p.Ws;
                    const long long off = ok ? (static_cast<long long>(ib[i] + yy * p.Ws + xx) * p.Cs + cq * 8) : 0;
                    const uint32_t dsta = a0 + static_cast<uint32_t>((rbase + 16 * i) * kRowBytes);
                    if (p.debug & 1) continue;
#pragma unroll
                    for (int pl = 0; pl < NPL; ++pl) cp_async16(dsta + pl * Cfg::A_BYTES, p.src + pl * p.src_plane + off, ok ? 16u : 0u);
                }
                // asynchronous publication: the stage's full barrier gets this thread's arrival when its copies land, so the
                // producers run ahead by as many stages as there are free slots and the MMA warp never waits on this loop
                cp_async_mbar_arrive_noinc(&full_bar[s]);
            }
        }
      }
    } else if (warp == 4) {
        // ======================= TMA producer: weights (and, with ATMA, the gathered operand) =======================
        if (lane == 0) {
            const int hw = p.Hg * p.Wg;
            const int ntaps = p.TH * p.TW;
            const uint32_t a_bytes = ATMA ? NPL * Cfg::A_BYTES : 0u;
            const bool prof = p.prof != nullptr;
            long long w_tma = 0, w_sched = 0;
            const long long t_begin = clock64();
            int it = 0;
            int next_tile = unit;          // scheduler: the ticket of tile lt + 1 is drawn while tile lt is being loaded
            unsigned int ticket = 0;
            bool drawn = false;
            for (int lt = 0;; ++lt) {
                int tile;
                if (PAIR && rank != 0) {
                    tile = ring_read<PAIR>(ring, tile_full, tile_empty, lt, p.err, rank);
                } else {
                    // ---- scheduler: next tile of this unit -> ring (of both CTAs of a pair)
                    const int slot = lt % kRing;
                    mbar_wait_wd(&tile_empty[slot], ((lt / kRing) & 1) ^ 1, p.err, 10);
                    if (drawn) {       // the ticket drawn one tile ago: its atomic round trip (~1-2 k cycles) ran under that tile's loads
                        if (ticket == static_cast<unsigned int>(total - 1)) atomicExch(p.sched, 0u);   // the last draw of the launch re-arms the counter
                        next_tile = static_cast<int>(ticket) + nunits;
                    }
                    tile = next_tile < total ? next_tile : -1;
                    ring[slot] = tile;
                    if (PAIR) {
                        st_shared_remote_u32(ring + slot, 1, static_cast<uint32_t>(tile));
                        mbar_arrive_remote(&tile_full[slot], 1);
                    }
                    mbar_arrive(&tile_full[slot]);
                    // the draw for the tile after this one is issued AFTER the publication and not looked at until the next
                    // iteration: neither the consumers nor this tile's first loads wait for the atomic
                    drawn = tile >= 0;
                    if (drawn) ticket = atomicAdd(p.sched, 1u);
                }
                if (tile < 0) break;
                const TileCoord tc = decode_tile<PAIR>(tile, p, rank);
                const GemmClass cls = p.cls[tc.cls_idx];
                const int row0 = cls.wrow + tc.tile_n * BN + ((PAIR && !Cfg::CAT) ? static_cast<int>(rank) * Cfg::BROWS : 0);
                // tile origin on the (image, row) grid; with ATMA a tile is a whole number of image rows
                const int pix0 = tc.tile_m * kBM;
                const int n0 = pix0 / hw;
                const int y0 = (pix0 - n0 * hw) / p.Wg;
                int cc = 0, tap = 0;
                for (int j = 0; j < ksteps; ++j, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait_prof(&empty_bar[s], ph ^ 1, p.err, 2, prof ? &w_tma : nullptr);
                    const int ta = tap / p.TW, tb = tap - ta * p.TW;
                    const int cx = tb + cls.ox, cy = y0 * p.sy + ta + cls.oy, cch = cc * kBK;
                    if (PAIR) {
                        // both CTAs load their operand shares; all bytes are credited to the leader's barrier
                        uint64_t* bar = ATMA ? &full_bar[s] : &bfull_bar[s];
                        if (rank == 0) mbar_arrive_expect_tx(bar, 2 * (Cfg::BPL * Cfg::B_BYTES + a_bytes));
                        // the planes are the outermost tensor-map dimension: one box brings all planes this CTA stages
                        // (CAT: the one plane `rank`, all BN rows; else both planes of this CTA's half of the rows)
                        tma_load_3d_pair(smem_u32(stage_b(s, 0)), &tmB0, bar, j * kBK, row0, Cfg::CAT ? static_cast<int>(rank) : 0);
                        if (ATMA) tma_load_5d_pair(smem_u32(stage_a(s, 0)), &tmA0, bar, cch, cx, cy, n0, 0);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[s], Cfg::BPL * Cfg::B_BYTES + a_bytes);
                        tma_load_3d(smem_u32(stage_b(s, 0)), &tmB0, &full_bar[s], j * kBK, row0, 0);
                        if (ATMA) tma_load_5d(smem_u32(stage_a(s, 0)), &tmA0, &full_bar[s], cch, cx, cy, n0, 0);
                    }
                    if (++tap == ntaps) { tap = 0; ++cc; }      // K order: (channel chunk, tap)
                }
            }
            if (prof) {
                p.prof[blockIdx.x * 8 + 0] = clock64() - t_begin;
                p.prof[blockIdx.x * 8 + 1] = w_tma;
                p.prof[blockIdx.x * 8 + 2] = w_sched;
            }
        }
    } else if (warp == 5) {
      if (PAIR && rank != 0) {
        if (RELAY) {
            // ======================= peer CTA: relay "my A rows have landed" to the leader =======================
            int it = 0;
            for (int lt = 0;; ++lt) {
                if (ring_read_warp<PAIR>(ring, tile_full, tile_empty, lt, p.err, rank) < 0) break;
                for (int j = 0; j < ksteps; ++j, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait_wd(&full_bar[s], ph, p.err, 5);
                    if (lane == 0) {
                        fence_proxy_async_smem();
                        mbar_arrive_remote(&peer_bar[s], 0);
                    }
                    __syncwarp();
                }
            }
        }
      } else {
        // ======================= MMA issuer (pair: the leader CTA issues for both) =======================
        const uint32_t idesc = idesc_f16(PAIR ? 2 * kBM : kBM, Cfg::NMMA, 0, 0, p.src_fmt, p.w_fmt);
        int it = 0;
        const bool prof = p.prof != nullptr;
        long long w_full = 0, w_acce = 0, w_ring = 0;
        int ntiles = 0;
        for (int lt = 0;; ++lt) {
            const long long tr0 = clock64();
            if (ring_read_warp<PAIR>(ring, tile_full, tile_empty, lt, p.err, rank) < 0) break;
            if (prof) w_ring += clock64() - tr0;
            ++ntiles;
            const int ab = lt & 1;
            const uint32_t aph = (lt >> 1) & 1;
            mbar_wait_prof(&acce_bar[ab], aph ^ 1, p.err, 8, prof ? &w_acce : nullptr);      // the epilogue has drained this accumulator (first use: free)
            tc_fence_after();
            const uint32_t tmem_buf = tmem_base + static_cast<uint32_t>(ab * Cfg::BUF_COLS);
            for (int j = 0; j < ksteps; ++j, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait_prof(&full_bar[s], ph, p.err, 3, prof ? &w_full : nullptr);
                if (RELAY) {
                    mbar_wait_wd(&peer_bar[s], ph, p.err, 6);
                    mbar_wait_wd(&bfull_bar[s], ph, p.err, 7);
                }
                fence_proxy_async_smem();   // cp.async wrote through the generic proxy; the MMA reads through the async proxy
                tc_fence_after();
                if (lane == 0) {
                    if (Cfg::CAT) {
                        // one MMA per activation plane and K = 16 slice: A_a x [B_0 | B_1], N = NPL * BN
#pragma unroll
                        for (int kk = 0; kk < kBK / 16; ++kk) {
                            const uint64_t bd = smem_desc(smem_u32(stage_b(s, 0)) + kk * 32, 16, 1024, 2u);
#pragma unroll
                            for (int pa = 0; pa < NPL; ++pa) {
                                if ((p.debug & 2) && (j > 0 || kk > 0 || pa > 0)) break;
                                const uint64_t ad = smem_desc(smem_u32(stage_a(s, pa)) + kk * 32, 16, 1024, 2u);
                                const uint32_t acc = (j > 0 || kk > 0 || pa > 0) ? 1u : 0u;
                                if (PAIR) umma_bf16_pair(tmem_buf, ad, bd, idesc, acc);
                                else umma_bf16(tmem_buf, ad, bd, idesc, acc);
                            }
                        }
                    } else {
#pragma unroll
                        for (int pass = 0; pass < NPASS; ++pass) {
                            if ((p.debug & 2) && (j > 0 || pass > 0)) break;
                            // plane pairs (a, b) in the order {00, 01, 10, 02, 20, 11}: NPASS 1 / 3 / 6 take a prefix
                            const int pa = (pass == 2 || pass == 5) ? 1 : (pass == 4 ? 2 : 0);
                            const int pb = (pass == 1 || pass == 5) ? 1 : (pass == 3 ? 2 : 0);
                            const uint32_t abase = smem_u32(stage_a(s, pa));
                            const uint32_t bbase = smem_u32(stage_b(s, pb));
#pragma unroll
                            for (int kk = 0; kk < kBK / 16; ++kk) {
                                // K-major, SWIZZLE_128B (layout type 2): 8-row groups are 1024 bytes apart; +32 bytes per K=16 slice
                                const uint64_t ad = smem_desc(abase + kk * 32, 16, 1024, 2u);
                                const uint64_t bd = smem_desc(bbase + kk * 32, 16, 1024, 2u);
                                const uint32_t acc = (j > 0 || pass > 0 || kk > 0) ? 1u : 0u;
                                if (PAIR) umma_bf16_pair(tmem_buf, ad, bd, idesc, acc);
                                else umma_bf16(tmem_buf, ad, bd, idesc, acc);
                            }
                        }
                    }
                    if (PAIR) umma_commit_pair(&empty_bar[s]); else umma_commit(&empty_bar[s]);
                }
                __syncwarp();
            }
            if (lane == 0) {
                if (PAIR) umma_commit_pair(&accf_bar[ab]); else umma_commit(&accf_bar[ab]);
            }
            __syncwarp();
        }
        if (prof && lane == 0) {
            p.prof[blockIdx.x * 8 + 3] = w_full;
            p.prof[blockIdx.x * 8 + 4] = w_acce;
            p.prof[blockIdx.x * 8 + 5] = w_ring;
            p.prof[blockIdx.x * 8 + 6] = ntiles;
        }
      }
    } else {
        // ======================= epilogue (two groups of four warps), one / two tiles behind the MMA warp =======================
        constexpr int CW = BN < 32 ? 16 : 32;         // accumulator columns per tcgen05.ld
        constexpr int NCH = BN < 32 ? 1 : BN / 32;
        const int t = egrp ? threadIdx.x - 192 : threadIdx.x;    // index inside the group
        const int bar_id = 1 + egrp;
        EpiRow er;
        er.p = &p;
        er.alpha = p.sigma ? p.alpha_k / __ldg(p.sigma) : p.alpha_k;
        er.negslope = p.act == 1 ? 0.1f : (p.act == 2 ? 0.f : 1.f);
        er.auxslope = p.aux_mode == 1 ? 0.1f : 0.f;      // derivative on the non-positive side
        er.part = part;
        er.warp4 = warp & 3;
        er.lane = lane;
        er.bn = BN;
        const int row = (warp & 3) * 32 + lane;                  // TMEM lane = accumulator row of this thread
        const bool aux_bits = p.aux != nullptr && p.aux_mode != 3;
        // launch kind -> specialisation of the per-chunk code (anything unusual takes the generic form)
        const bool fast = p.act != 3 && (p.aux == nullptr || aux_bits) && (p.out_mode == 2 || p.dst_npl == 2);
        const int out_kind = p.out_mode == 2 ? 2 : (p.dst_npl == 2 ? (p.dst_fmt == FMT_BF16 ? 1 : (p.dst_fmt == FMT_F16A ? 0 : 3)) : 3);
        const int kind = (fast && out_kind != 3) ? 1 + out_kind + 3 * (aux_bits ? 1 : 0) + 6 * (p.colsum ? 1 : 0) : 0;
        for (int lt = 0;; ++lt) {
            const int tile = ring_read_warp<PAIR>(ring, tile_full, tile_empty, lt, p.err, rank);
            if (tile < 0) break;
            if ((lt & 1) != egrp) continue;                    // the other group's tile
            const TileCoord tc = decode_tile<PAIR>(tile, p, rank);
            const GemmClass cls = p.cls[tc.cls_idx];
            const int ab = lt & 1;
            const uint32_t aph = (lt >> 1) & 1;
            // destination pixel of this thread's row
            {
                const int m = tc.tile_m * kBM + row;
                long long pr = -1;
                if (m < p.M) {
                    const int hw = p.Hg * p.Wg;
                    const int n = m / hw;
                    const int rem = m - n * hw;
                    const int y = rem / p.Wg;
                    const int x = rem - y * p.Wg;
                    pr = (static_cast<long long>(n) * p.Hd + (y * p.osy + cls.ooy)) * p.Wd + (x * p.osx + cls.oox);
                }
                er.prow = pr;
                er.arow = pr >= p.aux_wrap_at ? pr - p.aux_wrap_len : pr;
            }
            const int col0 = tc.tile_n * BN;
            // lrelu' / relu' need only the SIGN of the stored activation (plane 0): the 2 * CW bytes of this row's next chunk are
            // fetched one chunk ahead (chunk 0: before the accumulator is complete), so no load is waited for in the store path
            uint4 anext[4];
            auto fetch_aux = [&](int ch) {
                const int col = col0 + ch * CW;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    anext[k] = make_uint4(0u, 0u, 0u, 0u);
                    if (k * 8 < CW && er.prow >= 0 && col + k * 8 < p.Ncols)
                        anext[k] = __ldg(reinterpret_cast<const uint4*>(p.aux + er.arow * p.Cd + col + k * 8));
                }
            };
#pragma unroll
            for (int k = 0; k < 4; ++k) anext[k] = make_uint4(0u, 0u, 0u, 0u);
            if (aux_bits) fetch_aux(0);
            if (p.prof != nullptr && t == 0 && egrp == 0) {
                const long long te0 = clock64();
                mbar_wait_wd(&accf_bar[ab], aph, p.err, 4);
                p.prof[blockIdx.x * 8 + 7] += clock64() - te0;
            } else
                mbar_wait_wd(&accf_bar[ab], aph, p.err, 4);
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>(ab * Cfg::BUF_COLS);
#pragma unroll 1
            for (int ch = 0; ch < NCH; ++ch) {
                const int col = col0 + ch * CW;
                uint4 araw[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) araw[k] = anext[k];
                if (aux_bits && ch + 1 < NCH && col + CW < p.Ncols) fetch_aux(ch + 1);      // in flight while this chunk is processed
                float v[32];
                if (CW == 32) tmem_ld32(tmem_acc + ch * 32, v);
                else tmem_ld16(tmem_acc, v);
                if (Cfg::NACC > 1) {
                    // concatenated weight planes: column block a holds A x B_a; their sum is the result
#pragma unroll
                    for (int a = 1; a < Cfg::NACC; ++a) {
                        float u[32];
                        if (CW == 32) tmem_ld32(tmem_acc + a * BN + ch * 32, u);
                        else tmem_ld16(tmem_acc + a * BN, u);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < CW; ++j) v[j] += u[j];
                    }
                }
                tmem_ld_wait();
                if (ch == NCH - 1) {
                    // this warp has read its lanes of the accumulator: hand the buffer back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (PAIR && rank != 0) mbar_arrive_leader(&acce_bar[ab]); else mbar_arrive(&acce_bar[ab]);      // the leader arrives locally (a remote arrive on the own CTA is slow)
                    }
                }
                if (col >= p.Ncols) continue;            // padding columns of the last N tile (uniform over the group)
                const int lcol = ch * CW;
                switch (kind) {
                    case 1: epi_chunk<CW, 0, false, 0, false>(er, v, araw, col, lcol); break;     // forward conv -> fp16 planes
                    case 2: epi_chunk<CW, 0, false, 1, false>(er, v, araw, col, lcol); break;     // -> bf16 planes
                    case 3: epi_chunk<CW, 0, false, 2, false>(er, v, araw, col, lcol); break;     // -> raw fp32
                    case 4: epi_chunk<CW, 1, false, 0, false>(er, v, araw, col, lcol); break;
                    case 5: epi_chunk<CW, 1, false, 1, false>(er, v, araw, col, lcol); break;     // input gradient x activation derivative
                    case 6: epi_chunk<CW, 1, false, 2, false>(er, v, araw, col, lcol); break;
                    case 7: epi_chunk<CW, 0, true, 0, false>(er, v, araw, col, lcol); break;
                    case 8: epi_chunk<CW, 0, true, 1, false>(er, v, araw, col, lcol); break;
                    case 9: epi_chunk<CW, 0, true, 2, false>(er, v, araw, col, lcol); break;      // pre-batch-norm output + sum x, sum x^2
                    case 10: epi_chunk<CW, 1, true, 0, false>(er, v, araw, col, lcol); break;
                    case 11: epi_chunk<CW, 1, true, 1, false>(er, v, araw, col, lcol); break;     // input gradient + bias-gradient column sums
                    case 12: epi_chunk<CW, 1, true, 2, false>(er, v, araw, col, lcol); break;
                    default:
                        if (p.colsum) {
                            if (aux_bits) epi_chunk<CW, 1, true, 3, true>(er, v, araw, col, lcol);
                            else if (p.aux) epi_chunk<CW, 2, true, 3, true>(er, v, araw, col, lcol);
                            else epi_chunk<CW, 0, true, 3, true>(er, v, araw, col, lcol);
                        } else {
                            if (aux_bits) epi_chunk<CW, 1, false, 3, true>(er, v, araw, col, lcol);
                            else if (p.aux) epi_chunk<CW, 2, false, 3, true>(er, v, araw, col, lcol);
                            else epi_chunk<CW, 0, false, 3, true>(er, v, araw, col, lcol);
                        }
                        break;
                }
            }
            if (p.colsum) {
                // four per-warp partials per column -> the tile's column sums, fixed order
                epi_bar(bar_id);
                const long long tl = static_cast<long long>(tc.cls_idx) * p.tiles_m + tc.tile_m;
                for (int cc = t; cc < BN; cc += kEpiThreads) {
                    const int col = col0 + cc;
                    if (col < p.Ncols) {
                        float s1 = 0.f, s2 = 0.f;
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            s1 += part[(w * 2 + 0) * BN + cc];
                            s2 += part[(w * 2 + 1) * BN + cc];
                        }
                        p.colsum[tl * p.Ncols + col] = s1;
                        if (p.colsumsq) p.colsumsq[tl * p.Ncols + col] = s2;
                    }
                }
                epi_bar(bar_id);      // the partials are rewritten by the group's next tile
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

template <int BN, int NPASS, bool ATMA>
__global__ void __launch_bounds__(ATMA ? kThreadsTma : kThreadsGather, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmA0,
                 const __grid_constant__ ConvGemmParams p) {
    conv_gemm_body<BN, NPASS, false, ATMA>(tmB0, tmA0, p);
}

template <int BN, int NPASS, bool ATMA>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ATMA ? kThreadsTma : kThreadsGather, 1)
conv_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmA0,
                      const __grid_constant__ ConvGemmParams p) {
    conv_gemm_body<BN, NPASS, true, ATMA>(tmB0, tmA0, p);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || !ptr) return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// bf16 plane tensor map [planes][rows][cols] (cols contiguous): dims {cols, rows, planes}, box {box_cols, box_rows,
// box_planes}, zero OOB fill.  swizzle: 0 = 128B, 2 = 64B
int make_tmap_planes(CUtensorMap* m, const uint16_t* base, long long rows, long long cols, long long row_stride_elems,
                     long long plane_stride_elems, int planes, int box_cols, int box_rows, int box_planes, int swizzle) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return -1;
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(planes)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(row_stride_elems) * 2, static_cast<cuuint64_t>(plane_stride_elems > 0 ? plane_stride_elems : rows * row_stride_elems) * 2};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows), static_cast<cuuint32_t>(box_planes)};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapSwizzle sw = swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<uint16_t*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

// 5-D bf16 tensor map over NHWC activation planes: dims {C, W, H, N, planes}; box {64, bw, bh, bn, npl} with traversal
// strides {1, sx, sy, 1, 1} (the box loads ceil(b/s) elements per strided dimension), 128-byte swizzle, zero OOB fill
int make_tmap_act(CUtensorMap* m, const uint16_t* base, long long plane_stride_elems, int npl, int C, int W, int H, int N, int bw,
                         int bh, int bn, int sx, int sy) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return -1;
    cuuint64_t dims[5] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N),
                          static_cast<cuuint64_t>(npl)};
    const cuuint64_t img = static_cast<cuuint64_t>(H) * W * C * 2;
    cuuint64_t strides[4] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2, img,
                             npl > 1 ? static_cast<cuuint64_t>(plane_stride_elems) * 2 : img * N};
    cuuint32_t box[5] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh), static_cast<cuuint32_t>(bn),
                         static_cast<cuuint32_t>(npl)};
    cuuint32_t estr[5] = {1, static_cast<cuuint32_t>(sx), static_cast<cuuint32_t>(sy), 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<uint16_t*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

// a 128-row tile is `hb` whole image rows of `nb` images: the gathered operand can be one TMA box per (chunk, tap)
static bool atma_geometry(const ConvGemmParams& p, int* hb, int* nb) {
    if (p.Cs % kBK != 0 || p.Wg <= 0 || kBM % p.Wg != 0) return false;   // whole 32-channel chunks, whole image rows
    int rows = kBM / p.Wg;
    if (rows <= p.Hg) {
        if (p.Hg % rows != 0) return false;
        *hb = rows; *nb = 1;
    } else {
        if (rows % p.Hg != 0) return false;
        *hb = p.Hg; *nb = rows / p.Hg;
    }
    if (p.Wg * p.sx > 256 || *hb * p.sy > 256 || *nb > 256) return false;
    return true;
}

// ticket counters of the dynamic tile scheduler: one 4-byte slot per launch, handed out round-robin from a per-device pool.
// A kernel leaves its slot at zero (the last ticket drawn re-arms it), so a slot can be reused by any LATER launch; a CUDA
// graph keeps the slot its kernel node was captured with.  The pool is far larger than the number of launches that can be
// in flight at once, so two concurrently running kernels never share a slot.
static constexpr unsigned kSchedSlots = 1u << 16;
static unsigned int* sched_slot() {
    static unsigned int* pool[64] = {};
    static unsigned seq[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!pool[dev]) {
        // first use happens on an eager launch (the engine's first step is never captured), so allocating here is legal
        if (cudaMalloc(&pool[dev], kSchedSlots * sizeof(unsigned int)) != cudaSuccess) return nullptr;
        if (cudaMemset(pool[dev], 0, kSchedSlots * sizeof(unsigned int)) != cudaSuccess) return nullptr;
    }
    return pool[dev] + (seq[dev]++ % kSchedSlots);
}
static int num_sms() {
    static int n[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!n[dev] && cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n[dev] = 148;
    return n[dev];
}

template <int BN, int NPASS, bool PAIR, bool ATMA>
static int launch_cfg(const ConvGemmParams& p, const uint16_t* w, long long w_plane, long long w_rows, int kpad, int classes, int hb,
                      int nb, cudaStream_t st) {
    using Cfg = GemmCfg<BN, NPASS, PAIR>;
    constexpr int THREADS = ATMA ? kThreadsTma : kThreadsGather;
    CUtensorMap t0, a0;
    if (make_tmap_planes(&t0, w, w_rows, kpad, kpad, w_plane, Cfg::NPL, kBK, Cfg::BROWS, Cfg::BPL, 0)) return -4;
    if (ATMA) {
        if (make_tmap_act(&a0, p.src, p.src_plane, Cfg::NPL, p.Cs, p.Ws, p.Hs, p.Nimg, p.Wg * p.sx, hb * p.sy, nb, p.sx, p.sy)) return -4;
    } else {
        a0 = t0;
    }
    static int max_units = 0;      // persistent CTAs (CTA pairs) that are co-resident on the device: one per SM (TPC)
    if (!max_units) {
        cudaError_t e;
        if constexpr (PAIR) e = cudaFuncSetAttribute(conv_gemm_pair_kernel<BN, NPASS, ATMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        else e = cudaFuncSetAttribute(conv_gemm_kernel<BN, NPASS, ATMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return -4;
        const int sms = num_sms();
        if constexpr (PAIR) {
            int clusters = 0;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(static_cast<unsigned>(sms / 2 * 2), 1, 1);
            cfg.blockDim = dim3(THREADS, 1, 1);
            cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
            cudaLaunchAttribute attr;
            attr.id = cudaLaunchAttributeClusterDimension;
            attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
            cfg.attrs = &attr;
            cfg.numAttrs = 1;
            if (cudaOccupancyMaxActiveClusters(&clusters, conv_gemm_pair_kernel<BN, NPASS, ATMA>, &cfg) != cudaSuccess || clusters <= 0) {
                (void)cudaGetLastError();
                clusters = sms / 2;
            }
            max_units = clusters < sms / 2 ? clusters : sms / 2;
        } else {
            max_units = sms;
        }
    }
    ConvGemmParams q = p;
    q.tiles_m = (p.M + kBM - 1) / kBM;
    if (PAIR) q.tiles_m = (q.tiles_m + 1) / 2 * 2;
    q.tiles_n = (p.Ncols + BN - 1) / BN;
    q.classes = classes;
    q.sched = sched_slot();
    if (!q.sched) return -4;
    static int prof_on = -1;
    if (prof_on < 0) { const char* e = getenv("MMDGAN_PROF"); prof_on = e ? atoi(e) : 0; }
    static long long* prof_buf = nullptr;
    q.prof = nullptr;
    if (prof_on && p.M >= 4096) {
        if (!prof_buf) cudaMalloc(&prof_buf, 512 * 8 * sizeof(long long));
        cudaMemsetAsync(prof_buf, 0, 512 * 8 * sizeof(long long), st);
        q.prof = prof_buf;
    }
    const int total = q.tiles_n * classes * (PAIR ? q.tiles_m / 2 : q.tiles_m);
    const int units = total < max_units ? total : max_units;
    dim3 grid(static_cast<unsigned>(PAIR ? 2 * units : units), 1, 1);
    cudaError_t le;
    if constexpr (PAIR) le = launch_pdl(conv_gemm_pair_kernel<BN, NPASS, ATMA>, grid, dim3(THREADS), Cfg::SMEM_BYTES, st, t0, a0, q);
    else le = launch_pdl(conv_gemm_kernel<BN, NPASS, ATMA>, grid, dim3(THREADS), Cfg::SMEM_BYTES, st, t0, a0, q);
    if (le != cudaSuccess) return -4;
    if (q.prof) {       // profiling experiment: blocks the host; never on in the product path
        static long long host[512 * 8];
        cudaStreamSynchronize(st);
        cudaMemcpy(host, prof_buf, sizeof(host), cudaMemcpyDeviceToHost);
        const int nb_ = static_cast<int>(grid.x);
        double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int lead = 0;
        for (int b = 0; b < nb_ && b < 512; ++b) {
            if (PAIR && (b & 1)) { a[0] += 0; continue; }
            ++lead;
            for (int k = 0; k < 8; ++k) a[k] += static_cast<double>(host[b * 8 + k]);
        }
        fprintf(stderr, "PROF bn=%d npass=%d pair=%d atma=%d stages=%d M=%d N=%d ksteps=%d classes=%d grid=%d tiles/unit=%.2f | cycles/unit: total %.0f tma_wait_empty %.0f sched %.0f | mma_wait_full %.0f mma_wait_acce %.0f mma_ring %.0f | epi0_wait_accf %.0f\n",
                BN, NPASS, PAIR ? 1 : 0, ATMA ? 1 : 0, Cfg::STAGES, p.M, p.Ncols, p.ksteps, classes, nb_, a[6] / lead, a[0] / lead, a[1] / lead, a[2] / lead,
                a[3] / lead, a[4] / lead, a[5] / lead, a[7] / lead);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// w: [planes][w_rows][kpad]; w_rows = classes * rows_per_class (rows_per_class a multiple of the N tile)
int launch_conv_gemm(const ConvGemmParams& p, const uint16_t* w, long long w_plane, long long w_rows, int kpad, int classes,
                     int bn, int npass, int pair, cudaStream_t st) {
    int hb = 0, nb = 0;
    static int no_atma = -1;
    if (no_atma < 0) { const char* e = getenv("MMDGAN_NO_ATMA"); no_atma = e ? atoi(e) : 0; }
    const bool atma = !no_atma && atma_geometry(p, &hb, &nb);
#define MG_CASE(B, N) \
    if (bn == B && npass == N && !pair) \
        return atma ? launch_cfg<B, N, false, true>(p, w, w_plane, w_rows, kpad, classes, hb, nb, st) \
                    : launch_cfg<B, N, false, false>(p, w, w_plane, w_rows, kpad, classes, hb, nb, st);
#define MG_PAIR(B, N) \
    if (bn == B && npass == N && pair) \
        return atma ? launch_cfg<B, N, true, true>(p, w, w_plane, w_rows, kpad, classes, hb, nb, st) \
                    : launch_cfg<B, N, true, false>(p, w, w_plane, w_rows, kpad, classes, hb, nb, st);
    MG_CASE(16, 6) MG_CASE(32, 6) MG_CASE(64, 6) MG_CASE(128, 6)
    MG_CASE(16, 3) MG_CASE(32, 3) MG_CASE(64, 3) MG_CASE(128, 3) MG_CASE(256, 3)
    MG_CASE(16, 1) MG_CASE(32, 1) MG_CASE(64, 1) MG_CASE(128, 1) MG_CASE(256, 1)
    MG_PAIR(64, 6) MG_PAIR(128, 6) MG_PAIR(256, 6) MG_PAIR(64, 3) MG_PAIR(128, 3) MG_PAIR(256, 3) MG_PAIR(64, 1) MG_PAIR(128, 1) MG_PAIR(256, 1)
#undef MG_CASE
#undef MG_PAIR
    return -1;
}

}  // namespace mg
