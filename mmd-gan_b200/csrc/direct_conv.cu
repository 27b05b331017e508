// Direct 3x3 / stride-1 / SAME convolution on the CUDA cores for the two image-channel layers of every shipped
// architecture (D's first conv 3 -> 64, G's last conv 64 -> 3) and their input gradients.
//
// As implicit GEMMs these layers are the worst case of the tensor-core kernel: with 3 (padded 8) channels on one side the
// MMA tile is > 90 % padding, and the 64-channel operand is re-fetched from L2 once per filter tap (ncu: 7.7 TB/s of
// L2 -> SM traffic for 0.3 GFLOP).  Here a block owns a 16 x 16 pixel tile, stages the input halo ONCE in shared memory
// as fp32 (the bf16 planes are summed on the way in), keeps the 1728 filter weights in shared memory and does plain
// fp32 FMAs -- exact products, so these launches need no plane-pair passes at all.
//   LS ("large -> small"): Cin a multiple of 16 (<= 128), Cout <= 4.   forward of 64 -> 3 (+ bias, tanh), input gradient of 3 -> 64
//                          (x act_k / sigma, x tanh'(x_gen), per-block column sums = bias gradient of the generator's last layer)
//   SL ("small -> large"): Cin <= 4, Cout a multiple of 16 (<= 128).   forward of 3 -> 64 (+ bias, lrelu), input gradient of 64 -> 3
// Weights are read from the reference's canonical layout [k][k][Cin][Cout] through (tap, in, out) strides, flipped for the
// input gradients; nothing is packed.
//
// Replaces tf.nn.conv2d (GeneralTools/layer_func.py:912-916) for those two layers and its input gradient
// (DeepLearning/my_sngan.py:301-304).
#include "conv_gemm.cuh"
#include "tc_common.cuh"

namespace mg {

static constexpr int kDT = 16;            // tile edge (pixels)
static constexpr int kDH = kDT + 2;       // halo edge
static constexpr int kDMaxC = 128;        // largest "large" channel count

__device__ __forceinline__ float d_act(float v, int act) {
    if (act == 1) return v > 0.f ? v : 0.1f * v;
    if (act == 2) return v > 0.f ? v : 0.f;
    if (act == 3) return tanhf(v);
    return v;
}
__device__ __forceinline__ float d_act_grad(float a, int mode) {
    if (mode == 1) return a > 0.f ? 1.f : 0.1f;
    if (mode == 2) return a > 0.f ? 1.f : 0.f;
    if (mode == 3) return 1.f - a * a;
    return 1.f;
}
__device__ __forceinline__ void add8(const uint4 q, int fmt, float* v) {
    if (fmt == FMT_BF16) {
        v[0] += __uint_as_float(q.x << 16); v[1] += __uint_as_float(q.x & 0xFFFF0000u);
        v[2] += __uint_as_float(q.y << 16); v[3] += __uint_as_float(q.y & 0xFFFF0000u);
        v[4] += __uint_as_float(q.z << 16); v[5] += __uint_as_float(q.z & 0xFFFF0000u);
        v[6] += __uint_as_float(q.w << 16); v[7] += __uint_as_float(q.w & 0xFFFF0000u);
    } else {
        const float4 a = unpack_f16x4(make_uint2(q.x, q.y)), b = unpack_f16x4(make_uint2(q.z, q.w));
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
}
// sum of the planes (fp16 formats: still carrying the format's scale)
__device__ __forceinline__ void load8_planes(const bf16_t* src, long long plane, int npl, int fmt, long long off, float* v) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    for (int pl = 0; pl < npl; ++pl) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(src + pl * plane + off));
        add8(q, fmt, v);
    }
}

// ------------------------------------------------------------------------------------------------ large -> small
__global__ void __launch_bounds__(256) conv3x3_ls_kernel(const DirectConvParams p) {
    constexpr int PITCH = 20;                                   // 16 channels + 4 pad floats: conflict-free 16-byte reads
    __shared__ __align__(16) float tile[kDH * kDH * PITCH];     // 25.3 KB
    __shared__ __align__(16) float wsm[9 * kDMaxC * 4];         // [tap][ci][4 outputs]  18 KB
    __shared__ float red[8][4];
    const int t = threadIdx.x;
    const int tiles_x = (p.W + kDT - 1) / kDT, tiles_y = (p.H + kDT - 1) / kDT;
    int b = blockIdx.x;
    const int txi = b % tiles_x; b /= tiles_x;
    const int tyi = b % tiles_y;
    const int n = b / tiles_y;
    const int x0 = txi * kDT, y0 = tyi * kDT;
    const int CI = p.Cin;
    for (int i = t; i < 9 * CI * 4; i += 256) {
        const int o = i & 3, ci = (i >> 2) % CI, tap = (i >> 2) / CI;
        wsm[i] = o < p.Cout ? p.w[(p.flip ? 8 - tap : tap) * p.w_tap + ci * p.w_in + o * p.w_out] : 0.f;
    }
    const int lx = t & 15, ly = t >> 4;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    // halo staging: 648 sixteen-byte units (324 pixels x 2 halves of the 16-channel chunk) over 256 threads = up to 3 per
    // thread; the raw plane words of the NEXT chunk are fetched into registers while the current chunk is being computed
    constexpr int UPT = (kDH * kDH * 2 + 255) / 256;
    long long uoff[UPT];
    uint4 raw[UPT][3];
#pragma unroll
    for (int k = 0; k < UPT; ++k) {
        const int u = t + k * 256, px = u >> 1;
        const int gy = y0 + px / kDH - 1, gx = x0 + px % kDH - 1;
        uoff[k] = (u < kDH * kDH * 2 && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)
                      ? (static_cast<long long>(n * p.H + gy) * p.W + gx) * p.Cs + (u & 1) * 8 : -1;
    }
    auto fetch = [&](int c0) {
#pragma unroll
        for (int k = 0; k < UPT; ++k)
#pragma unroll
            for (int pl = 0; pl < 3; ++pl) {
                raw[k][pl] = make_uint4(0u, 0u, 0u, 0u);
                if (uoff[k] >= 0 && pl < p.src_npl) raw[k][pl] = __ldg(reinterpret_cast<const uint4*>(p.src + pl * p.src_plane + uoff[k] + c0));
            }
    };
    fetch(0);
    for (int c0 = 0; c0 < CI; c0 += 16) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < UPT; ++k) {
            const int u = t + k * 256;
            if (u < kDH * kDH * 2) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = 0.f;
#pragma unroll
                for (int pl = 0; pl < 3; ++pl) add8(raw[k][pl], p.src_fmt, v);      // absent planes were fetched as zeros
                float4* d = reinterpret_cast<float4*>(tile + (u >> 1) * PITCH + (u & 1) * 8);
                d[0] = make_float4(v[0], v[1], v[2], v[3]);
                d[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
        __syncthreads();
        if (c0 + 16 < CI) fetch(c0 + 16);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const float* xp = tile + ((ly + tap / 3) * kDH + lx + tap % 3) * PITCH;
            const float4* wp = reinterpret_cast<const float4*>(wsm) + tap * CI + c0;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float4 x = *reinterpret_cast<const float4*>(xp + c4 * 4);
                const float4 w0 = wp[c4 * 4], w1 = wp[c4 * 4 + 1], w2 = wp[c4 * 4 + 2], w3 = wp[c4 * 4 + 3];
                acc0 = fmaf(x.x, w0.x, acc0); acc1 = fmaf(x.x, w0.y, acc1); acc2 = fmaf(x.x, w0.z, acc2); acc3 = fmaf(x.x, w0.w, acc3);
                acc0 = fmaf(x.y, w1.x, acc0); acc1 = fmaf(x.y, w1.y, acc1); acc2 = fmaf(x.y, w1.z, acc2); acc3 = fmaf(x.y, w1.w, acc3);
                acc0 = fmaf(x.z, w2.x, acc0); acc1 = fmaf(x.z, w2.y, acc1); acc2 = fmaf(x.z, w2.z, acc2); acc3 = fmaf(x.z, w2.w, acc3);
                acc0 = fmaf(x.w, w3.x, acc0); acc1 = fmaf(x.w, w3.y, acc1); acc2 = fmaf(x.w, w3.z, acc2); acc3 = fmaf(x.w, w3.w, acc3);
            }
        }
    }
    // ---- epilogue: alpha, bias, activation or activation derivative, planes / raw output, column sums
    const int gy = y0 + ly, gx = x0 + lx;
    const bool ok = gy < p.H && gx < p.W;
    const float alpha = (p.sigma ? p.alpha_k / __ldg(p.sigma) : p.alpha_k) * fmt_inv_scale(p.src_fmt);   // the staged inputs carry the source scale
    float v[4] = {acc0, acc1, acc2, acc3};
    const long long prow = static_cast<long long>(n * p.H + gy) * p.W + gx;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        v[o] = o < p.Cout ? d_act(fmaf(v[o], alpha, p.bias ? p.bias[o] : 0.f), p.act) : 0.f;
        if (p.aux && ok && o < p.Cout) v[o] *= d_act_grad(load_val(p.aux, p.aux_plane, p.aux_npl, p.aux_fmt, prow * p.Cd + o), p.aux_mode);
        if (!ok) v[o] = 0.f;
    }
    if (ok) {
        const float4 lo = make_float4(v[0], v[1], v[2], v[3]), z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.out_mode == 0) {
            bf16_t* o = static_cast<bf16_t*>(p.dst) + prow * p.Cd;
            note_saturation4(p.sat_flag, p.dst_fmt, lo);
            store_vals4(o, p.dst_plane, p.dst_npl, p.dst_fmt, lo);
            for (int c = 4; c < p.Cd; c += 4) store_vals4(o + c, p.dst_plane, p.dst_npl, p.dst_fmt, z4);
        } else {
            float* o = static_cast<float*>(p.dst) + prow * p.Cd;
            *reinterpret_cast<float4*>(o) = lo;
            for (int c = 4; c < p.Cd; c += 4) *reinterpret_cast<float4*>(o + c) = z4;
        }
    }
    if (p.colsum) {          // fixed-order block reduction: warp shuffles, then the 8 warp sums
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) v[o] += __shfl_xor_sync(0xffffffffu, v[o], s);
        if ((t & 31) == 0)
#pragma unroll
            for (int o = 0; o < 4; ++o) red[t >> 5][o] = v[o];
        __syncthreads();
        if (t < p.Cd) {
            float s = 0.f;
            if (t < 4)
                for (int w = 0; w < 8; ++w) s += red[w][t];
            p.colsum[static_cast<long long>(blockIdx.x) * p.Cd + t] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------------ small -> large
__global__ void __launch_bounds__(256) conv3x3_sl_kernel(const DirectConvParams p) {
    __shared__ __align__(16) float tile[kDH * kDH * 4];          // 3 channels + pad, 5 KB
    __shared__ __align__(16) float wsm[9 * 4 * kDMaxC];          // [tap][ci (4)][co]  18 KB
    const int t = threadIdx.x;
    const int tiles_x = (p.W + kDT - 1) / kDT, tiles_y = (p.H + kDT - 1) / kDT;
    int b = blockIdx.x;
    const int txi = b % tiles_x; b /= tiles_x;
    const int tyi = b % tiles_y;
    const int n = b / tiles_y;
    const int x0 = txi * kDT, y0 = tyi * kDT;
    const int CO = p.Cout;
    for (int i = t; i < 9 * 4 * CO; i += 256) {
        const int co = i % CO, ci = (i / CO) & 3, tap = i / (4 * CO);
        wsm[i] = ci < p.Cin ? p.w[(p.flip ? 8 - tap : tap) * p.w_tap + ci * p.w_in + co * p.w_out] : 0.f;
    }
    for (int px = t; px < kDH * kDH; px += 256) {
        const int gy = y0 + px / kDH - 1, gx = x0 + px % kDH - 1;
        float v[8];
        if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)
            load8_planes(p.src, p.src_plane, p.src_npl, p.src_fmt, (static_cast<long long>(n * p.H + gy) * p.W + gx) * p.Cs, v);
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
        }
        *reinterpret_cast<float4*>(tile + px * 4) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
    __shared__ __align__(16) float stage[256 * 20];             // one 16-channel group of the tile, pitch 20 floats
    const int lx = t & 15, ly = t >> 4;
    float x[9][3];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const float4 q = *reinterpret_cast<const float4*>(tile + ((ly + tap / 3) * kDH + lx + tap % 3) * 4);
        x[tap][0] = q.x; x[tap][1] = q.y; x[tap][2] = q.z;
    }
    const float alpha = (p.sigma ? p.alpha_k / __ldg(p.sigma) : p.alpha_k) * fmt_inv_scale(p.src_fmt);   // the staged inputs carry the source scale
    for (int co0 = 0; co0 < CO; co0 += 16) {
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const float4* wp = reinterpret_cast<const float4*>(wsm + (tap * 4 + ci) * CO + co0);
                const float xv = x[tap][ci];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 w = wp[q];
                    acc[4 * q] = fmaf(xv, w.x, acc[4 * q]);
                    acc[4 * q + 1] = fmaf(xv, w.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(xv, w.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(xv, w.w, acc[4 * q + 3]);
                }
            }
        // stage the group in shared memory, then write it with four consecutive threads per pixel: every store instruction
        // fills whole 32-byte sectors (a thread writing its own pixel's 16 channels would touch each sector four times)
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(stage + t * 20 + 4 * q) = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = t + k * 256, px = idx >> 2, q = idx & 3;
            const int gy = y0 + (px >> 4), gx = x0 + (px & 15);
            if (gy >= p.H || gx >= p.W) continue;
            const float4 a = *reinterpret_cast<const float4*>(stage + px * 20 + 4 * q);
            float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) bq = *reinterpret_cast<const float4*>(p.bias + co0 + 4 * q);
            float4 v;
            v.x = d_act(fmaf(a.x, alpha, bq.x), p.act);
            v.y = d_act(fmaf(a.y, alpha, bq.y), p.act);
            v.z = d_act(fmaf(a.z, alpha, bq.z), p.act);
            v.w = d_act(fmaf(a.w, alpha, bq.w), p.act);
            const long long o = (static_cast<long long>(n * p.H + gy) * p.W + gx) * p.Cd + co0 + 4 * q;
            if (p.out_mode == 0) {
                note_saturation4(p.sat_flag, p.dst_fmt, v);
                store_vals4(static_cast<bf16_t*>(p.dst) + o, p.dst_plane, p.dst_npl, p.dst_fmt, v);
            } else *reinterpret_cast<float4*>(static_cast<float*>(p.dst) + o) = v;
        }
    }
}

int direct_conv_blocks(int N, int H, int W) { return N * ((H + kDT - 1) / kDT) * ((W + kDT - 1) / kDT); }

int launch_direct_conv(const DirectConvParams& p, cudaStream_t st) {
    const int blocks = direct_conv_blocks(p.N, p.H, p.W);
    if (p.Cout <= 4 && p.Cin % 16 == 0 && p.Cin <= kDMaxC) conv3x3_ls_kernel<<<blocks, 256, 0, st>>>(p);
    else if (p.Cin <= 4 && p.Cout % 16 == 0 && p.Cout <= kDMaxC) conv3x3_sl_kernel<<<blocks, 256, 0, st>>>(p);
    else return -1;
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// ================================================================================================ image layers as 1x1 GEMMs
// Round 2: the two image-channel layers cost 0.54 ms per step for 1 % of its FLOPs as direct convolutions.  The many -> few
// direction of a 3x3 / stride-1 convolution with 3 channels on one side (64 -> 3 forward, 3 -> 64 input gradient) is a DENSE
// product over 27 = 9 taps x 3 channels plus a shifted sum, and the dense product runs on the tcgen05 gather-GEMM like any layer:
//   T[p][(tap, c)] = sum_k x[p][k] W[k][(tap, c)]  ([pixels, 64] x [64, 27], fp32 out), then
//   y[p][c] = sum_tap T[p +- off(tap)][(tap, c)] with the layer's epilogue (tapsum27_kernel): 0.057 ms instead of 0.091.
// (The few -> many direction as im2col + dense [27 -> 64] was built too and measured slower than the direct kernel; removed.)
static constexpr int kImgK = 32;      // 27 columns padded to one 16-bit 64-byte row / 32 fp32

// y[p][c] = epilogue( alpha * sum_tap T[p + s * off(tap)][tap * 3 + c] ), c < 3; T fp32 [pixels][32]; s = +1 (forward of the
// many -> few convolution) or -1 (input gradient of the few -> many one).  Epilogue as the direct kernels': bias, activation
// or activation derivative from `aux`, 16-bit planes (channels 3 .. Cd-1 zero) or raw fp32, per-block column sums.
__global__ void __launch_bounds__(256) tapsum27_kernel(const float* __restrict__ T, int N, int H, int W, int flip, float alpha_k,
                                                      const float* __restrict__ sigma, const float* __restrict__ bias, int act,
                                                      const uint16_t* __restrict__ aux, long long aux_plane, int aux_npl, int aux_fmt,
                                                      int aux_mode, void* __restrict__ dst, long long dst_plane, int dst_npl,
                                                      int dst_fmt, int Cd, int out_mode, float* __restrict__ colsum, int* __restrict__ sat_flag) {
    __shared__ float red[8][4];
    const long long total = static_cast<long long>(N) * H * W;
    const long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const bool ok = p < total;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (ok) {
        const int x = static_cast<int>(p % W);
        const int y = static_cast<int>((p / W) % H);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            const int yy = flip ? y - dy : y + dy, xx = flip ? x - dx : x + dx;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                const float* t = T + (p + static_cast<long long>(yy - y) * W + (xx - x)) * kImgK + tap * 3;
                a0 += __ldg(t); a1 += __ldg(t + 1); a2 += __ldg(t + 2);
            }
        }
        const float alpha = sigma ? alpha_k / __ldg(sigma) : alpha_k;
        v[0] = d_act(fmaf(a0, alpha, bias ? bias[0] : 0.f), act);
        v[1] = d_act(fmaf(a1, alpha, bias ? bias[1] : 0.f), act);
        v[2] = d_act(fmaf(a2, alpha, bias ? bias[2] : 0.f), act);
        if (aux) {
#pragma unroll
            for (int o = 0; o < 3; ++o) v[o] *= d_act_grad(load_val(aux, aux_plane, aux_npl, aux_fmt, p * Cd + o), aux_mode);
        }
        const float4 lo = make_float4(v[0], v[1], v[2], 0.f), z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (out_mode == 0) {
            bf16_t* o = static_cast<bf16_t*>(dst) + p * Cd;
            note_saturation4(sat_flag, dst_fmt, lo);
            store_vals4(o, dst_plane, dst_npl, dst_fmt, lo);
            for (int c = 4; c < Cd; c += 4) store_vals4(o + c, dst_plane, dst_npl, dst_fmt, z4);
        } else {
            float* o = static_cast<float*>(dst) + p * Cd;
            *reinterpret_cast<float4*>(o) = lo;
            for (int c = 4; c < Cd; c += 4) *reinterpret_cast<float4*>(o + c) = z4;
        }
    }
    if (colsum) {          // fixed-order block reduction: warp shuffles, then the 8 warp sums
        const int t = threadIdx.x;
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) v[o] += __shfl_xor_sync(0xffffffffu, v[o], s);
        if ((t & 31) == 0)
#pragma unroll
            for (int o = 0; o < 4; ++o) red[t >> 5][o] = v[o];
        __syncthreads();
        if (t < Cd) {
            float s = 0.f;
            if (t < 4)
                for (int w = 0; w < 8; ++w) s += red[w][t];
            colsum[static_cast<long long>(blockIdx.x) * Cd + t] = s;
        }
    }
}

int tapsum_blocks(int N, int H, int W) { return static_cast<int>((static_cast<long long>(N) * H * W + 255) / 256); }

int launch_tapsum27(const float* T, int N, int H, int W, int flip, float alpha_k, const float* sigma, const float* bias, int act, const uint16_t* aux,
                    long long aux_plane, int aux_npl, int aux_fmt, int aux_mode, void* dst, long long dst_plane, int dst_npl, int dst_fmt, int Cd,
                    int out_mode, float* colsum, int* sat_flag, cudaStream_t st) {
    tapsum27_kernel<<<tapsum_blocks(N, H, W), 256, 0, st>>>(T, N, H, W, flip, alpha_k, sigma, bias, act, aux, aux_plane, aux_npl, aux_fmt, aux_mode,
                                                           dst, dst_plane, dst_npl, dst_fmt, Cd, out_mode, colsum, sat_flag);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace mg
