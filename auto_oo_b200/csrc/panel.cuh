// Shared pieces of the panel GEMM: the stand-alone kernel of dgemm_small.cu and the single-launch expm chain of
// expm.cu run the same staging and the same DMMA loop (same summation order), and differ only in their epilogues.
#pragma once
#include "common.cuh"

namespace oo {
namespace panel {

constexpr int TS = 32;        // CTA tile (TS x TS)
constexpr int SPAD = TS + 4;  // smem row stride (doubles) of a k-major panel: (t*36 + g) mod 16 distinct per half-warp

// ---- K <= 256, 16-byte aligned operands
// The pipelined kernel above still pays one L2 round trip per 16-deep k-block (18 us at N = 256).  When the
// whole K extent of the two panels fits shared memory (<= 2 x 74 KB) every 16-byte chunk is requested at
// once with cp.async (zero-filled past the matrix edge), the CTA waits ONCE, and eight warps (four output
// quadrants x two K halves, combined through shared memory) run the DMMAs back to back: one memory latency
// plus ~1 us of tensor-pipe time per product.  Used for the N x N products of expm / C' = X C U / C^T h C.
constexpr int PK_MAX = 256;
constexpr int PTHREADS = 256;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
                 "r"(src_bytes)
                 : "memory");
}

// stage op(X)[f0 .. f0+32, 0 .. K) : "free-major" (fm = 1: smem[f][k], row stride K + 4) when the global rows run
// along k, "k-major" (smem[k][f], row stride 36) when they run along the free index
// smem row stride of a free-major panel: K rounded up to 16, + 4 (= 4 mod 16: conflict-free fragment loads)
__host__ __device__ inline int panel_ls(int K) { return ((K + 15) & ~15) + 4; }

__device__ __forceinline__ void stage_panel(double *smem, const double *__restrict__ X, int ldx, bool rows_along_k,
                                            int f0, int F, int K, int K4) {
    if (rows_along_k) {
        const int ls = panel_ls(K), chunks = K / 2;                // per row: K/2 chunks of 2 doubles
        for (int i = threadIdx.x; i < TS * chunks; i += PTHREADS) {
            const int f = i / chunks, c = i - f * chunks;
            const bool in = f0 + f < F;
            const double *src = X + (int64_t)(in ? f0 + f : 0) * ldx + 2 * c;
            cp_async16(smem + f * ls + 2 * c, src, in ? 16 : 0);
        }
        for (int i = threadIdx.x; i < TS * (K4 - K); i += PTHREADS)         // k in [K, K4): zero
            smem[(i / (K4 - K)) * ls + K + i % (K4 - K)] = 0.0;
    } else {
        for (int i = threadIdx.x; i < K * (TS / 2); i += PTHREADS) {
            const int k = i / (TS / 2), c = i - k * (TS / 2);
            const int f = f0 + 2 * c;
            const int bytes = f + 1 < F ? 16 : (f < F ? 8 : 0);
            const double *src = X + (int64_t)k * ldx + (bytes ? f : 0);
            cp_async16(smem + k * SPAD + 2 * c, src, bytes);
        }
        for (int i = threadIdx.x; i < (K4 - K) * TS; i += PTHREADS) smem[(K + i / TS) * SPAD + i % TS] = 0.0;
    }
}

// Product of two staged panels on eight warps (four 16 x 16 output quadrants x two K halves, combined through
// shared memory; `red` may alias the panels -- they are dead by then).  Returns true for the four warps that own
// the output; their acc holds the whole sum.  Two CTA barriers inside; callers that go on to stage another tile
// need one more before they overwrite `red`.
__device__ __forceinline__ bool tile_mma(const double *sA, const double *sB, bool a_fm, bool b_fm, int K, int K4,
                                         double *red, double (&acc)[2][2][2]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int quad = warp & 3, khalf = warp >> 2;
    const int wm = (quad >> 1) * 16, wn = (quad & 1) * 16;
    const int a_sf = a_fm ? panel_ls(K) : 1, a_sk = a_fm ? 1 : SPAD;    // strides of the free / k index in smem
    const int b_sf = b_fm ? panel_ls(K) : 1, b_sk = b_fm ? 1 : SPAD;
    const int ksteps = K4 / 4, kmid = ((ksteps + 1) / 2) * 4;
    const int kbeg = khalf ? kmid : 0, kend = khalf ? K4 : kmid;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 4
    for (int kk = kbeg; kk < kend; kk += 4) {
        double a[2], bf[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            a[i] = sA[(wm + i * 8 + g) * a_sf + (kk + t) * a_sk];
            bf[i] = sB[(wn + i * 8 + g) * b_sf + (kk + t) * b_sk];
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bf[j]);
    }
    __syncthreads();                                              // panels are dead: reuse as the reduction buffer
    if (khalf) {                                                  // [4 quadrants][8 values][32 lanes]
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int c = 0; c < 2; ++c) red[(quad * 8 + i * 4 + j * 2 + c) * 32 + lane] = acc[i][j][c];
    }
    __syncthreads();
    if (khalf) return false;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c) acc[i][j][c] += red[(quad * 8 + i * 4 + j * 2 + c) * 32 + lane];
    return true;
}

// dynamic shared memory of one tile: both panels, or the reduction buffer if that is larger
inline size_t smem_bytes(int K, int transA, int transB) {
    const size_t K4 = (size_t)((K + 3) & ~3);
    const size_t a = !transA ? (size_t)TS * panel_ls(K) : K4 * SPAD;
    const size_t b = transB ? (size_t)TS * panel_ls(K) : K4 * SPAD;
    const size_t red = 4 * 8 * 32;
    return ((a + b) > red ? (a + b) : red) * sizeof(double);
}

}  // namespace panel
}  // namespace oo
