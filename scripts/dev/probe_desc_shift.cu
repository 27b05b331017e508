// Round-2 experiment (NOT part of the library): can one staged activation patch serve all filter taps of a convolution?
//
// Plan ("halo-resident gathered operand"): enumerate a tile's output pixels over the zero-padded row
// width, so that tap (dy, dx) is the SAME shared-memory patch read at a start address shifted by dy * (W + 2) + dx rows.
// That needs a K-major tcgen05 shared-memory descriptor whose start address is moved by an ARBITRARY number of operand
// rows.  Two layouts are probed, each against a CPU product of the logically shifted operand:
//
//   mode 0  SWIZZLE_64B, rows of 64 B (32 bf16), 8-row groups 512 B apart -- the layout conv_gemm.cu stages today.  The
//           data are written with the swizzle taken from ABSOLUTE address bits [7,9) (what TMA and the cp.async producers
//           do); the descriptor start moves by shift * 64 B, with the descriptor's base-offset field (bits [49,52)) either
//           0 or (start >> 7) & 7.
//   mode 1  no swizzle, "row-linear": for each 16-byte K chunk a contiguous array rows x 16 B (core matrices of 8 rows
//           = 128 contiguous bytes, SBO = 128, LBO = rows * 16), so that a row shift is start += shift * 16 B.
//           Both assignments of (LBO, SBO) are tried.
//   mode 2  SWIZZLE_128B, rows of 128 B (64 bf16), 8-row groups 1024 B apart: one operand row is exactly the 128-byte
//           granule the base-offset field counts, so if the hardware derives the swizzle phase from (row index relative
//           to the start address + base offset) rather than from absolute address bits, THIS layout still admits every
//           row shift (mode 0 would then admit even shifts only).  Same two base-offset variants as mode 0.
//
// Build and run on the GPU box:
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I mmd-gan_b200/csrc scripts/dev/probe_desc_shift.cu \
//        -o /tmp/probe_desc_shift -lcuda && /tmp/probe_desc_shift
// Output: one line per (mode, variant, shift) with the number of mismatching accumulator elements (0 = the shift works).
#include "tc_common.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

using namespace mg;

static constexpr int kRows = 192;   // staged operand rows (128 MMA rows + up to 64 rows of shift)
static constexpr int kN = 64;       // accumulator columns
static constexpr int kK = 64;       // bf16 per logical row (modes 0 / 1 use the first 32: two K = 16 MMAs; mode 2 all 64: four)

struct ProbeParams {
    const uint16_t* a;   // [kRows][kK] bf16 bits, logical order
    const uint16_t* b;   // [kN][kK]
    float* d;            // [128][kN]
    int mode;            // 0 = SWIZZLE_64B, 1 = no swizzle (row-linear), 2 = SWIZZLE_128B
    int variant;         // mode 0: 0 = base offset 0, 1 = base offset (start >> 7) & 7.  mode 1: 0 = (LBO = chunk stride, SBO = 128), 1 = swapped
    int shift;           // operand rows
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const ProbeParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;                          // kRows * 128 B = 24 KB at most
    uint8_t* sb = smem + kRows * 128;            // kN * 128 B = 8 KB at most (1024-aligned: 24 KB is a multiple of 1024)
    uint64_t* bar = reinterpret_cast<uint64_t*>(sb + kN * 128);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

    // stage both operands through the generic proxy, 16 bytes (one K chunk of one row) at a time
    const int cpr = p.mode == 2 ? 8 : 4;         // 16-byte chunks per staged row
    for (int u = t; u < (kRows + kN) * cpr; u += 128) {
        const bool is_a = u < kRows * cpr;
        const int row = (is_a ? u : u - kRows * cpr) / cpr;
        const int chunk = u % cpr;
        const uint4 v = *reinterpret_cast<const uint4*>((is_a ? p.a : p.b) + row * kK + chunk * 8);
        uint8_t* base = is_a ? sa : sb;
        const int rows = is_a ? kRows : kN;
        uint32_t off;
        if (p.mode == 0) off = row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);   // absolute-address swizzle (bases are 1024-aligned)
        else if (p.mode == 2) off = row * 128 + ((chunk ^ (row & 7)) << 4);
        else off = chunk * rows * 16 + row * 16;
        *reinterpret_cast<uint4*>(base + off) = v;
    }
    if (t == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, 64);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (t == 0) {
        const uint32_t idesc = idesc_bf16(128, kN, 0, 0);
        for (int kk = 0; kk < (p.mode == 2 ? 4 : 2); ++kk) {
            uint64_t ad, bd;
            if (p.mode == 2) {
                const uint32_t astart = smem_u32(sa) + p.shift * 128 + kk * 32;
                ad = smem_desc(astart, 16, 1024, 2u);
                if (p.variant == 1) ad |= static_cast<uint64_t>((astart >> 7) & 7u) << 49;
                bd = smem_desc(smem_u32(sb) + kk * 32, 16, 1024, 2u);
            } else if (p.mode == 0) {
                const uint32_t astart = smem_u32(sa) + p.shift * 64 + kk * 32;
                ad = smem_desc(astart, 16, 512, 4u);
                if (p.variant == 1) ad |= static_cast<uint64_t>((astart >> 7) & 7u) << 49;
                bd = smem_desc(smem_u32(sb) + kk * 32, 16, 512, 4u);
            } else {
                const uint32_t a_lbo = kRows * 16, b_lbo = kN * 16;
                const uint32_t astart = smem_u32(sa) + p.shift * 16 + kk * 2 * a_lbo;
                const uint32_t bstart = smem_u32(sb) + kk * 2 * b_lbo;
                ad = p.variant == 0 ? smem_desc(astart, a_lbo, 128, 0u) : smem_desc(astart, 128, a_lbo, 0u);
                bd = p.variant == 0 ? smem_desc(bstart, b_lbo, 128, 0u) : smem_desc(bstart, 128, b_lbo, 0u);
            }
            umma_bf16(tmem_base, ad, bd, idesc, kk > 0 ? 1u : 0u);
        }
        umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    tc_fence_after();
    float v[32];
    for (int cc = 0; cc < kN / 32; ++cc) {
        tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + cc * 32, v);
        tmem_ld_wait();
        for (int q = 0; q < 32; ++q) p.d[(warp * 32 + lane) * kN + cc * 32 + q] = v[q];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 64);
    }
}

static uint16_t bf16_bits(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    return static_cast<uint16_t>(u >> 16);   // small integers are exact in bf16
}

int main() {
    std::vector<float> a(kRows * kK), b(kN * kK);
    std::vector<uint16_t> ab(a.size()), bb(b.size());
    srand(1);
    for (size_t i = 0; i < a.size(); ++i) { a[i] = static_cast<float>(rand() % 9 - 4); ab[i] = bf16_bits(a[i]); }
    for (size_t i = 0; i < b.size(); ++i) { b[i] = static_cast<float>(rand() % 9 - 4); bb[i] = bf16_bits(b[i]); }
    uint16_t *da, *db;
    float* dd;
    cudaMalloc(&da, ab.size() * 2);
    cudaMalloc(&db, bb.size() * 2);
    cudaMalloc(&dd, 128 * kN * 4);
    cudaMemcpy(da, ab.data(), ab.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, bb.data(), bb.size() * 2, cudaMemcpyHostToDevice);
    const int smem_bytes = kRows * 128 + kN * 128 + 64 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    const int shifts[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 15, 16, 17, 18, 19, 33, 34, 35, 36, 64};
    std::vector<float> d(128 * kN);
    int failures = 0;
    for (int mode = 0; mode < 3; ++mode)
        for (int variant = 0; variant < 2; ++variant)
            for (int shift : shifts) {
                cudaMemset(dd, 0xFF, 128 * kN * 4);
                ProbeParams p{da, db, dd, mode, variant, shift};
                probe_kernel<<<1, 128, smem_bytes>>>(p);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) {
                    printf("mode %d variant %d shift %2d : CUDA error %s\n", mode, variant, shift, cudaGetErrorString(e));
                    return 2;   // a sticky error ends the probe: rerun with the failing case removed
                }
                cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
                int bad = 0;
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < kN; ++n) {
                        float ref = 0.f;
                        for (int k = 0; k < (mode == 2 ? 64 : 32); ++k) ref += a[(m + shift) * kK + k] * b[n * kK + k];
                        if (d[m * kN + n] != ref) ++bad;
                    }
                printf("mode %d (%s) variant %d shift %2d : %5d / %d mismatches%s\n", mode, mode == 0 ? "SWIZZLE_64B " : (mode == 1 ? "row-linear  " : "SWIZZLE_128B"),
                       variant, shift, bad, 128 * kN, bad ? "" : "  OK");
                if (bad && shift == 0 && mode != 1 && variant == 0) ++failures;   // the control case must pass
            }
    if (failures) printf("CONTROL CASE FAILED: the probe itself is wrong\n");
    return failures ? 1 : 0;
}
