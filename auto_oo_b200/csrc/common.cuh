// Shared device/host helpers for the sm_100a kernels of the orbital-optimization hot path.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/oo_b200.h"

namespace oo {

// ---- error plumbing ---------------------------------------------------------
extern int g_last_cuda_error;
extern unsigned long long g_launch_count;   // kernels launched by this library (oo_launch_count)

inline int cuda_fail(cudaError_t e) {
    g_last_cuda_error = (int)e;
    return OO_ERR_CUDA;
}

#define OO_CUDA_CHECK(expr)                                   \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) return ::oo::cuda_fail(_e);    \
    } while (0)

#define OO_LAUNCH_CHECK()                      \
    do {                                       \
        ++::oo::g_launch_count;                \
        OO_CUDA_CHECK(cudaGetLastError());     \
    } while (0)

#define OO_REQUIRE(cond)                          \
    do {                                          \
        if (!(cond)) return OO_ERR_INVALID_ARG;   \
    } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

int sm_count();   // cached cudaDevAttrMultiProcessorCount of the current device

// Function attributes (opt-in dynamic shared memory) are per DEVICE: `once_per_device(mask)` is true the first time
// it is called with the calling thread's current device for a given call-site mask (devices 0..63; a process that
// drives several GPUs configures each of them).  Benign race: two threads may both configure the same device.
inline bool once_per_device(unsigned long long &mask) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return true;
    const unsigned long long bit = 1ull << dev;
    if (mask & bit) return false;
    mask |= bit;
    return true;
}

// ---- FP64 tensor-core MMA (DMMA.8x8x4) --------------------------------------
// Fragment ownership for lane = 4*g + t (g = 0..7, t = 0..3):
//   a = A[g][t]   (row-major 8x4),  b = B[t][g]  (col-major 4x8),
//   c0,c1 = C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---- shared-memory / mbarrier / TMA wrappers ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar_smem_addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_smem_addr) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// ---- reductions -----------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum; every thread gets the result.  `scratch` >= 32 doubles of smem.
__device__ __forceinline__ double block_sum(double v, double *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = (lane < nwarp) ? scratch[lane] : 0.0;
    r = warp_sum(r);
    return r;
}

// ---- views of the transformed two-electron integrals ----------------------------
// Every contraction of the hot path touches g' only through two access patterns with at
// least two indices m, n in I = occ + act:
//   coul(a,b,m,n) = g'[a,b,m,n]     (Coulomb class  (ab|mn))
//   exch(a,m,n,b) = g'[a,m,n,b]     (exchange class (am|nb))
// FullView reads them from the complete ld^4 tensor of the four-index transform;
// ClassView from the two class tensors of the partial transform (classes.cu):
//   J[m,n,a,b] = g'[a,b,m,n],  K[n,m,a,b] = g'[a,m,n,b],  each [nIp][nIp][ld][ld].
struct FullView {
    const double *g;
    int ld;
    int64_t batch_stride;
    __device__ __forceinline__ FullView at(int b) const { return {g + (int64_t)b * batch_stride, ld, batch_stride}; }
    __device__ __forceinline__ double coul(int a, int b, int m, int n) const {
        return g[(((int64_t)a * ld + b) * ld + m) * ld + n];
    }
    __device__ __forceinline__ double exch(int a, int m, int n, int b) const {
        return g[(((int64_t)a * ld + m) * ld + n) * ld + b];
    }
};

struct ClassView {
    const double *K, *J;      // K first: the Hessian's B operand is [K rows; J rows; h row]
    int ld, nIp;
    int64_t batch_stride;
    __device__ __forceinline__ ClassView at(int b) const {
        return {K + (int64_t)b * batch_stride, J + (int64_t)b * batch_stride, ld, nIp, batch_stride};
    }
    __device__ __forceinline__ double coul(int a, int b, int m, int n) const {
        return J[(((int64_t)m * nIp + n) * ld + a) * ld + b];
    }
    __device__ __forceinline__ double exch(int a, int m, int n, int b) const {
        return K[(((int64_t)n * nIp + m) * ld + a) * ld + b];
    }
};

// ---- host: TMA descriptor encode (driver entry point fetched at run time) -----
int encode_tmap_3d_f64(CUtensorMap *map, const void *base, uint64_t dim0, uint64_t dim1,
                       uint64_t dim2, uint64_t stride1_elems, uint64_t stride2_elems,
                       uint32_t box0, uint32_t box1);

}  // namespace oo
