// Gradient all-reduce FUSED with the Adam update, through the NVSwitch (NVLink SHARP / "NVLS" multicast objects).
//
// Data-parallel training keeps the parameters replicated, so after the backward pass every rank needs
//   g = sum over ranks of its local gradient;  (w, m, v) <- Adam(w, m, v, g)            (my_sngan.py:424-426)
// The plain path is ncclAllReduce(g) followed by adam_kernel on every rank (N ranks each do the whole update).  Here each
// rank owns 1/N of the flat parameter buffer and, for its shard only,
//   * loads the SUM of all ranks' gradients with one `multimem.ld_reduce` per 16 bytes -- the switch adds the N replicas in
//     flight, nothing is staged --,
//   * applies tf.train.AdamOptimizer's update (graph_func.py:518-527) to its own w / m / v, and
//   * publishes the new w, m, v to ALL replicas with `multimem.st` (the switch fans the store out).
// Per rank and step that is n/N loads and 3n/N stores on NVLink instead of a 2(N-1)/N * n ring all-reduce plus N full Adam
// passes over HBM, and -- because exactly one rank computes each element -- the replicas are bit-identical by
// construction.  The four buffers live in ONE symmetric allocation per network ([g | w | m | v]); the host brackets the
// launch with the allocation's device-side barriers (all gradients written before / all parameters visible after).
//
// The same switch feature carries the forward exchange of the MMD loss (scatter_scores_nvls_kernel below).
//
// Opt-in (MMDGAN_NVLS_ADAM=1, world size > 1, multicast-capable fabric); the default multi-GPU path is NCCL.
#include <cuda_runtime.h>
#include <stdint.h>

namespace mg {

__device__ __forceinline__ float4 multimem_ld_reduce_add_f32x4(const float* mc_ptr) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(mc_ptr)
                 : "memory");
    return r;
}
__device__ __forceinline__ void multimem_st_f32x4(float* mc_ptr, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// w, m, v: this rank's replicas (plain device pointers into the symmetric allocation); g_mc, w_mc, m_mc, v_mc: the same four
// buffers through the multicast mapping.  Elements [begin, end) (multiples of 4) are this rank's shard.
// (no __restrict__ / read-only loads on w, m, v: the same memory is written through the multicast mapping)
__global__ void adam_allreduce_nvls_kernel(const float* w, const float* m, const float* v,
                                           const float* g_mc, float* w_mc, float* m_mc, float* v_mc, long long begin, long long end,
                                           float lr, float b1, float b2, float eps, const int* __restrict__ step_ptr) {
    const double t = static_cast<double>(*step_ptr);
    const float lr_t = static_cast<float>(static_cast<double>(lr) * sqrt(1.0 - pow(static_cast<double>(b2), t)) /
                                          (1.0 - pow(static_cast<double>(b1), t)));
    const long long n4 = (end - begin) >> 2;
    for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < n4;
         q += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long i = begin + (q << 2);
        const float4 gi = multimem_ld_reduce_add_f32x4(g_mc + i);
        const float4 wi = *reinterpret_cast<const float4*>(w + i);
        const float4 mi = *reinterpret_cast<const float4*>(m + i);
        const float4 vi = *reinterpret_cast<const float4*>(v + i);
        float4 mo, vo, wo;
        mo.x = b1 * mi.x + (1.0f - b1) * gi.x;  vo.x = b2 * vi.x + (1.0f - b2) * gi.x * gi.x;  wo.x = wi.x - lr_t * mo.x / (sqrtf(vo.x) + eps);
        mo.y = b1 * mi.y + (1.0f - b1) * gi.y;  vo.y = b2 * vi.y + (1.0f - b2) * gi.y * gi.y;  wo.y = wi.y - lr_t * mo.y / (sqrtf(vo.y) + eps);
        mo.z = b1 * mi.z + (1.0f - b1) * gi.z;  vo.z = b2 * vi.z + (1.0f - b2) * gi.z * gi.z;  wo.z = wi.z - lr_t * mo.z / (sqrtf(vo.z) + eps);
        mo.w = b1 * mi.w + (1.0f - b1) * gi.w;  vo.w = b2 * vi.w + (1.0f - b2) * gi.w * gi.w;  wo.w = wi.w - lr_t * mo.w / (sqrtf(vo.w) + eps);
        multimem_st_f32x4(m_mc + i, mo);
        multimem_st_f32x4(v_mc + i, vo);
        multimem_st_f32x4(w_mc + i, wo);
    }
}

// Score exchange of the data-parallel MMD loss (parallel.gather_scores: all-gather + two re-ordering copies) as ONE multicast
// kernel: s_local [2b, d] holds this rank's scores, rows [0, b) real and [b, 2b) generated (my_sngan.py:279); they become rows
// [rank * b, (rank + 1) * b) of real_all / gen_all on EVERY rank.  The host brackets the launch with cross-rank barriers.
__global__ void scatter_scores_nvls_kernel(const float* __restrict__ s_local, int b, int d, int rank, float* gen_all_mc, float* real_all_mc) {
    const int n4 = (2 * b * d) >> 2;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += gridDim.x * blockDim.x) {
        const int e = q << 2;
        const int row = e / d, col = e - row * d;
        const float4 v = *reinterpret_cast<const float4*>(s_local + e);
        float* dst = row < b ? real_all_mc + static_cast<long long>(rank * b + row) * d + col
                             : gen_all_mc + static_cast<long long>(rank * b + row - b) * d + col;
        multimem_st_f32x4(dst, v);
    }
}

int l_scatter_scores_nvls(const float* s_local, int b, int d, int rank, float* gen_all_mc, float* real_all_mc, cudaStream_t st) {
    const int n4 = (2 * b * d) >> 2;
    if (n4 <= 0) return 0;
    scatter_scores_nvls_kernel<<<(n4 + 255) / 256, 256, 0, st>>>(s_local, b, d, rank, gen_all_mc, real_all_mc);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// Sum all-reduce of a handful of floats (the six kernel sums of the MMD loss): out[i] = sum over ranks of in[i], every rank
// reading the switch-reduced value.  n is a multiple of 4; `in_mc` is the multicast address of the per-rank source slots.
__global__ void allreduce_small_nvls_kernel(float* __restrict__ out, const float* in_mc, int n) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q * 4 < n) *reinterpret_cast<float4*>(out + q * 4) = multimem_ld_reduce_add_f32x4(in_mc + q * 4);
}

int l_allreduce_small_nvls(float* out, const float* in_mc, int n, cudaStream_t st) {
    if (n <= 0) return 0;
    allreduce_small_nvls_kernel<<<(n / 4 + 63) / 64, 64, 0, st>>>(out, in_mc, n);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

int l_adam_allreduce_nvls(const float* w, const float* m, const float* v, const float* g_mc, float* w_mc, float* m_mc, float* v_mc,
                          long long begin, long long end, float lr, float b1, float b2, float eps, const int* step, cudaStream_t st) {
    const long long n4 = (end - begin) >> 2;
    if (n4 <= 0) return 0;
    long long blocks = (n4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;      // grid-stride over the shard: 8 resident blocks of 256 threads per SM
    adam_allreduce_nvls_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(w, m, v, g_mc, w_mc, m_mc, v_mc, begin, end, lr, b1, b2, eps, step);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace mg
