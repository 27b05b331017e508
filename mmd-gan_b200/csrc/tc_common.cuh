// sm_100a primitives shared by the tensor-core kernels: mbarrier, TMA, tcgen05 (alloc / mma / commit / ld),
// shared-memory matrix descriptors, the bf16 instruction descriptor and the fp32 -> bf16 plane split.
// Inline PTX only; no CUTLASS.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdlib.h>

namespace mg {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---------------------------------------------------------------- proxies / cp.async
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// 16-byte global->shared copy; src_bytes == 0 zero-fills the destination (src must still be a valid address)
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
// same, but allocating in L1 (.ca): the gathered operand re-reads every input pixel once per filter tap, and with the
// taps innermost in the K order those re-reads are L1 hits instead of L2 -> SM traffic
__device__ __forceinline__ void cp_async16_ca(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one (pre-counted) arrival once every cp.async issued so far by this thread has landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------- TMA (tiled; the bf16 planes are the outermost dimension)
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            dst_smem),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst_smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
            dst_smem),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

// ---------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs (kind::f16), fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives when all tcgen05 ops previously issued by this thread are complete
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// warp w (w % 4 = TMEM lane quarter) reads its 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (sm_100 layout): start address [0,14) in 16-byte units, leading-dim byte offset
// [16,30), stride-dim byte offset [32,46), version (=1) [46,48), layout type [61,64) (2 = 128-byte swizzle).
// layout_type: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B (16-byte chunks XORed with the row index inside the 8-row group).
// K-major operands: rows of one swizzle span, 8-row groups SBO bytes apart, LBO unused.  MN-major operands
// (canonical layout ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) in bf16 elements): a 128-byte row holds 64 consecutive
// M/N elements of one K index, 8 K rows form an atom, atoms are LBO bytes apart along M/N and SBO bytes apart along K.
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    d |= static_cast<uint64_t>(layout_type) << 61;
    return d;
}
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return smem_desc(smem_addr, lbo_bytes, sbo_bytes, 2u);
}
// Plane formats (the `*_fmt` fields of the parameter blocks; include/mmdgan_b200.h):
//   FMT_BF16  bf16 planes of the value itself (three planes = fp32, two = 16 significand bits): gradients, and every
//             operand of the bf16-only modes
//   FMT_F16A  two fp16 planes of 16 x value   (activations / spectral-norm vectors feeding FORWARD launches)
//   FMT_F16W  two fp16 planes of 64 x value   (packed forward weights)
// fp16 carries 11 significand bits, so two planes hold 22 bits and the three products {00, 01, 10} are fp32-grade
// (~2^-21) at the FULL 16-bit tensor rate -- where bf16 needs three planes and six products.  The power-of-two scales keep
// the second plane of ordinary magnitudes in fp16's normal range (|x| >= 2^-7 for activations, 2^-9 for weights; smaller
// elements keep an ABSOLUTE accuracy of 2^-29 / 2^-31) and leave head-room up to |x| < 4094 / 1023; conversions
// saturate instead of overflowing.  The products carry the factor 16 * 64, removed through the epilogue alpha.
enum { FMT_BF16 = 0, FMT_F16A = 1, FMT_F16W = 2 };
__host__ __device__ constexpr float fmt_scale(int fmt) { return fmt == FMT_F16A ? 16.f : (fmt == FMT_F16W ? 64.f : 1.f); }
__host__ __device__ constexpr float fmt_inv_scale(int fmt) { return fmt == FMT_F16A ? 0.0625f : (fmt == FMT_F16W ? 0.015625f : 1.f); }

// Instruction descriptor for kind::f16 with fp32 accumulation: c_format F32 (1) at [4,6), a/b format (0 = F16, 1 = BF16) at
// [7,10)/[10,13) -- the two operands may differ --, a/b major (0 = K-major, 1 = MN-major) at 15/16, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_mn_major, int b_mn_major, int a_fmt, int b_fmt) {
    return (1u << 4) | ((a_fmt == FMT_BF16 ? 1u : 0u) << 7) | ((b_fmt == FMT_BF16 ? 1u : 0u) << 10) |
           (static_cast<uint32_t>(a_mn_major) << 15) | (static_cast<uint32_t>(b_mn_major) << 16) |
           (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return idesc_f16(M, N, a_mn_major, b_mn_major, FMT_BF16, FMT_BF16);
}

// ---------------------------------------------------------------- bf16 planes
// x = p0 + p1 + p2 with p0 = bf16_rn(x), p1 = bf16_rn(x - p0), p2 = bf16_rn(x - p0 - p1): three planes carry 24
// significand bits (the fp32 value, error <= 2^-25 |x|), two planes 16 bits.  The tensor-core kernels multiply plane
// pairs: npass 6 = {00, 01, 10, 02, 20, 11} (fp32-grade products, dropped terms <= 2^-24), npass 3 = {00, 01, 10}
// (error ~2^-17, used where the result is linear in the operand: input and weight gradients), npass 1 = {00}.
typedef uint16_t bf16_t;   // raw bf16 bits

__device__ __forceinline__ bf16_t f2bf(float x) { return __bfloat16_as_ushort(__float2bfloat16_rn(x)); }
__device__ __forceinline__ float bf2f(bf16_t b) { return __uint_as_float(static_cast<uint32_t>(b) << 16); }
__device__ __forceinline__ void bf16_split3(float x, bf16_t& p0, bf16_t& p1, bf16_t& p2) {
    p0 = f2bf(x);
    const float r1 = x - bf2f(p0);
    p1 = f2bf(r1);
    p2 = f2bf(r1 - bf2f(p1));
}
// store x as `npl` planes at dst[0], dst[plane], dst[2 * plane]
__device__ __forceinline__ void store_planes(bf16_t* dst, long long plane, int npl, float x) {
    bf16_t a, b, c;
    bf16_split3(x, a, b, c);
    dst[0] = a;
    if (npl > 1) dst[plane] = b;
    if (npl > 2) dst[2 * plane] = c;
}
// four consecutive elements (8-byte aligned): one 8-byte store per plane
__device__ __forceinline__ void store_planes4(bf16_t* dst, long long plane, int npl, float4 v) {
    bf16_t a[4], b[4], c[4];
    bf16_split3(v.x, a[0], b[0], c[0]);
    bf16_split3(v.y, a[1], b[1], c[1]);
    bf16_split3(v.z, a[2], b[2], c[2]);
    bf16_split3(v.w, a[3], b[3], c[3]);
    *reinterpret_cast<uint2*>(dst) = make_uint2(a[0] | (static_cast<uint32_t>(a[1]) << 16), a[2] | (static_cast<uint32_t>(a[3]) << 16));
    if (npl > 1)
        *reinterpret_cast<uint2*>(dst + plane) = make_uint2(b[0] | (static_cast<uint32_t>(b[1]) << 16), b[2] | (static_cast<uint32_t>(b[3]) << 16));
    if (npl > 2)
        *reinterpret_cast<uint2*>(dst + 2 * plane) = make_uint2(c[0] | (static_cast<uint32_t>(c[1]) << 16), c[2] | (static_cast<uint32_t>(c[3]) << 16));
}
__device__ __forceinline__ float load_planes(const bf16_t* src, long long plane, int npl, long long i) {
    float v = bf2f(src[i]);
    if (npl > 1) v += bf2f(src[i + plane]);
    if (npl > 2) v += bf2f(src[i + 2 * plane]);
    return v;
}
__device__ __forceinline__ float4 unpack_bf16x4(uint2 u) {
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xFFFF0000u));
}
__device__ __forceinline__ float4 load_planes4(const bf16_t* src, long long plane, int npl, long long i) {
    float4 v = unpack_bf16x4(*reinterpret_cast<const uint2*>(src + i));
    if (npl > 1) {
        const float4 m = unpack_bf16x4(*reinterpret_cast<const uint2*>(src + i + plane));
        v.x += m.x; v.y += m.y; v.z += m.z; v.w += m.w;
    }
    if (npl > 2) {
        const float4 l = unpack_bf16x4(*reinterpret_cast<const uint2*>(src + i + 2 * plane));
        v.x += l.x; v.y += l.y; v.z += l.z; v.w += l.w;
    }
    return v;
}

// ---------------------------------------------------------------- fp16 planes / format-generic access
__device__ __forceinline__ uint16_t f2h_sat(float x) {
    uint16_t h;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
    return h;
}
__device__ __forceinline__ float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }
// xs = scale * x (exact); p0 = fp16(xs), p1 = fp16(xs - p0)
__device__ __forceinline__ void f16_split2(float xs, uint16_t& p0, uint16_t& p1) {
    p0 = f2h_sat(xs);
    p1 = f2h_sat(xs - h2f(p0));
}
__device__ __forceinline__ uint32_t pack16(uint16_t a, uint16_t b) { return a | (static_cast<uint32_t>(b) << 16); }
// four consecutive elements in either format (npl planes; fp16 formats have at most two)
__device__ __forceinline__ void store_vals4(bf16_t* dst, long long plane, int npl, int fmt, float4 v) {
    if (fmt == FMT_BF16) {
        store_planes4(dst, plane, npl, v);
        return;
    }
    const float sc = fmt_scale(fmt);
    uint16_t a[4], b[4];
    f16_split2(v.x * sc, a[0], b[0]);
    f16_split2(v.y * sc, a[1], b[1]);
    f16_split2(v.z * sc, a[2], b[2]);
    f16_split2(v.w * sc, a[3], b[3]);
    *reinterpret_cast<uint2*>(dst) = make_uint2(pack16(a[0], a[1]), pack16(a[2], a[3]));
    if (npl > 1) *reinterpret_cast<uint2*>(dst + plane) = make_uint2(pack16(b[0], b[1]), pack16(b[2], b[3]));
}
// fp16 planes saturate at |scale * x| = 65504 instead of overflowing; a producer that can see large values reports it here so
// that the host can fail loudly (the precision claim of the fp16 formats holds below the saturation point only)
__device__ __forceinline__ void note_saturation4(int* flag, int fmt, float4 v) {
    if (flag == nullptr || fmt == FMT_BF16) return;
    const float m = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))) * fmt_scale(fmt);
    if (m > 65504.f) atomicExch(flag, 1);
}
__device__ __forceinline__ void store_val(bf16_t* dst, long long plane, int npl, int fmt, float x) {
    if (fmt == FMT_BF16) {
        store_planes(dst, plane, npl, x);
        return;
    }
    uint16_t a, b;
    f16_split2(x * fmt_scale(fmt), a, b);
    dst[0] = a;
    if (npl > 1) dst[plane] = b;
}
__device__ __forceinline__ float4 unpack_f16x4(uint2 u) {
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
// the VALUE (already unscaled) carried by the planes
__device__ __forceinline__ float4 load_vals4(const bf16_t* src, long long plane, int npl, int fmt, long long i) {
    if (fmt == FMT_BF16) return load_planes4(src, plane, npl, i);
    float4 v = unpack_f16x4(*reinterpret_cast<const uint2*>(src + i));
    if (npl > 1) {
        const float4 m = unpack_f16x4(*reinterpret_cast<const uint2*>(src + i + plane));
        v.x += m.x; v.y += m.y; v.z += m.z; v.w += m.w;
    }
    const float is = fmt_inv_scale(fmt);
    return make_float4(v.x * is, v.y * is, v.z * is, v.w * is);
}
__device__ __forceinline__ float load_val(const bf16_t* src, long long plane, int npl, int fmt, long long i) {
    if (fmt == FMT_BF16) return load_planes(src, plane, npl, i);
    float v = h2f(src[i]);
    if (npl > 1) v += h2f(src[i + plane]);
    return v * fmt_inv_scale(fmt);
}
// planes a launch reads from an operand in `fmt` for `npass` plane-pair products
__host__ __device__ constexpr int fmt_planes(int fmt, int npass) {
    return fmt == FMT_BF16 ? (npass == 6 ? 3 : (npass == 3 ? 2 : 1)) : (npass >= 3 ? 2 : 1);
}

// ---- cluster / cta_group::2 primitives (CTA-pair kernels)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
        "r"(rank)
        : "memory");
}
// TMA loads into THIS CTA's shared memory whose completion bytes are credited to the mbarrier of cluster CTA 0
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst_smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "{\n\t.reg .b32 rb;\n\t"
        "mapa.shared::cluster.u32 rb, %2, 0;\n\t"
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [rb];\n\t}" ::"r"(
            dst_smem),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst_smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3,
                                                 int c4) {
    asm volatile(
        "{\n\t.reg .b32 rb;\n\t"
        "mapa.shared::cluster.u32 rb, %2, 0;\n\t"
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [rb];\n\t}" ::"r"(
            dst_smem),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// completion of all prior MMAs of this thread arrives on the mbarrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
        "h"(static_cast<uint16_t>(3))
        : "memory");
}

// arrive on the mbarrier at this shared-memory offset in cluster CTA 0 (the pair's leader); rank 0 addresses itself
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) { mbar_arrive_remote(bar, 0); }
// wait with cluster-scope acquire: the data the barrier publishes was written by the OTHER CTA of the pair
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void st_shared_remote_u32(void* local_addr, uint32_t rank, uint32_t v) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "st.shared::cluster.u32 [ra], %2;\n\t}" ::"r"(smem_u32(local_addr)),
        "r"(rank), "r"(v)
        : "memory");
}

// ---- programmatic dependent launch.  A kernel launched with launch_pdl() may start (be scheduled, run its prologue: barrier
// initialisation, TMEM allocation, descriptor prefetch) while the kernel before it in the stream is still draining;
// pdl_wait() blocks until every prerequisite grid has COMPLETED and its memory is visible, so everything after it sees
// ordinary stream order.  pdl_launch_dependents() lets the NEXT kernel's CTAs be scheduled as soon as all of this grid's CTAs
// have passed it.  Both are no-ops for a launch without the attribute.  OPT-IN (MMDGAN_PDL=1): measured gain 0.01-0.02 ms per
// step (noise level) -- the persistent kernels leave no launch gap to hide -- and two of 16 bench runs with the attribute on
// hung (none of 50+ without it; not understood), so the attribute is not set by default.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#ifdef __CUDACC__
inline bool pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("MMDGAN_PDL"); on = e ? atoi(e) : 0; }
    return on != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

}  // namespace mg
