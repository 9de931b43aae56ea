// General small batched DGEMM with fused epilogue on FP64 tensor cores (DMMA.8x8x4):
//   D[b] = alpha * op(A[b]) op(B[b]) + beta * E[b] + gamma * I
// Used for the N x N products of the hot path: Pade/Newton-Schulz/squaring steps
// of expm(-kappa) (reference oo_energy.py:226-230), C' = X C_oao U (:173-176, :201)
// and C^T h C (:44-46).  Operands may be transposed and arbitrarily strided, so
// tiles are staged through padded shared memory with plain loads, register-prefetched
// one k-block ahead (these products are latency-bound: 2 N^3 flop on <= 256^2
// outputs); the large contractions use the TMA kernel in dgemm_tn.cu instead.
#include <stdlib.h>

#include "common.cuh"
#include "panel.cuh"

namespace oo {

namespace {
constexpr int TS = panel::TS;        // CTA tile (TS x TS), 4 warps of 16 x 16
constexpr int TK = 16;               // k-block
constexpr int SPAD = panel::SPAD;    // smem row stride (doubles): (t*36 + g) mod 16 distinct per half-warp

struct SmallArgs {
    const double *A, *B, *E;
    double *D;
    int M, N, K;
    int lda, ldb, lde, ldd;
    int64_t sA, sB, sE, sD;
    double alpha, beta, gamma;
    int transA, transB;
    int eye_n;   // gamma * I is added on rows < eye_n only (keeps zero padding zero)
};

// One k-block of op(A) / op(B) per thread: 4 + 4 elements, fetched into registers first (all eight loads in
// flight at once) and stored to shared memory afterwards; the next block's loads are issued before the DMMAs of
// the current one (two smem stages, one barrier per k-block).
struct Stage {
    double a[TS * TK / 128], b[TS * TK / 128];
};

__device__ __forceinline__ void fetch_block(const SmallArgs &p, const double *__restrict__ A,
                                            const double *__restrict__ B, int m0, int n0, int k0, Stage &r) {
#pragma unroll
    for (int i = 0; i < TS * TK / 128; ++i) {
        const int idx = threadIdx.x + i * 128;
        int k, f;
        if (p.transA) { k = idx / TS; f = idx % TS; } else { f = idx / TK; k = idx % TK; }
        const int gm = m0 + f, gk = k0 + k;
        double v = 0.0;
        if (gm < p.M && gk < p.K)
            v = p.transA ? __ldg(A + (int64_t)gk * p.lda + gm) : __ldg(A + (int64_t)gm * p.lda + gk);
        r.a[i] = v;
    }
#pragma unroll
    for (int i = 0; i < TS * TK / 128; ++i) {
        const int idx = threadIdx.x + i * 128;
        int k, f;
        if (!p.transB) { k = idx / TS; f = idx % TS; } else { f = idx / TK; k = idx % TK; }
        const int gn = n0 + f, gk = k0 + k;
        double v = 0.0;
        if (gn < p.N && gk < p.K)
            v = p.transB ? __ldg(B + (int64_t)gn * p.ldb + gk) : __ldg(B + (int64_t)gk * p.ldb + gn);
        r.b[i] = v;
    }
}

__device__ __forceinline__ void store_block(const SmallArgs &p, const Stage &r, double (*sA)[SPAD],
                                            double (*sB)[SPAD]) {
#pragma unroll
    for (int i = 0; i < TS * TK / 128; ++i) {
        const int idx = threadIdx.x + i * 128;
        int k, f;
        if (p.transA) { k = idx / TS; f = idx % TS; } else { f = idx / TK; k = idx % TK; }
        sA[k][f] = r.a[i];
        if (!p.transB) { k = idx / TS; f = idx % TS; } else { f = idx / TK; k = idx % TK; }
        sB[k][f] = r.b[i];
    }
}

__global__ void __launch_bounds__(128) dgemm_small_kernel(const SmallArgs p) {
    __shared__ double sA[2][TK][SPAD];   // sA[stage][k][m]
    __shared__ double sB[2][TK][SPAD];   // sB[stage][k][n]
    const int b = blockIdx.z;
    const int m0 = blockIdx.y * TS, n0 = blockIdx.x * TS;
    const double *A = p.A + (int64_t)b * p.sA;
    const double *B = p.B + (int64_t)b * p.sB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 1) * 16, wn = (warp & 1) * 16;

    double acc[2][2][2] = {};
    Stage r;
    fetch_block(p, A, B, m0, n0, 0, r);
    store_block(p, r, sA[0], sB[0]);
    __syncthreads();
    int cur = 0;
    for (int k0 = 0; k0 < p.K; k0 += TK) {
        const bool more = k0 + TK < p.K;
        if (more) fetch_block(p, A, B, m0, n0, k0 + TK, r);
#pragma unroll
        for (int kk = 0; kk < TK; kk += 4) {
            double a[2], bf[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                a[i] = sA[cur][kk + t][wm + i * 8 + g];
                bf[i] = sB[cur][kk + t][wn + i * 8 + g];
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bf[j]);
        }
        if (more) store_block(p, r, sA[cur ^ 1], sB[cur ^ 1]);
        __syncthreads();
        cur ^= 1;
    }

    double *D = p.D + (int64_t)b * p.sD;
    const double *E = p.E ? p.E + (int64_t)b * p.sE : nullptr;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int row = m0 + wm + i * 8 + g;
        if (row >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int col = n0 + wn + j * 8 + 2 * t + c;
                if (col >= p.N) continue;
                double v = p.alpha * acc[i][j][c];
                if (E) v += p.beta * E[(int64_t)row * p.lde + col];
                if (row == col && row < p.eye_n) v += p.gamma;
                D[(int64_t)row * p.ldd + col] = v;
            }
    }
}

// ---- panel variant: K <= 256, 16-byte aligned operands (staging and DMMA loop: panel.cuh) ------------------
using panel::PK_MAX;
using panel::PTHREADS;
using panel::panel_ls;
using panel::stage_panel;

__global__ void __launch_bounds__(PTHREADS) dgemm_panel_kernel(const SmallArgs p) {
    extern __shared__ __align__(16) double psm[];
    const int K = p.K, K4 = (K + 3) & ~3;
    const bool a_fm = !p.transA, b_fm = p.transB != 0;            // free-major staging (global rows run along k)
    double *sA = psm;
    double *sB = psm + (a_fm ? TS * panel_ls(K) : K4 * SPAD);
    const int b = blockIdx.z;
    const int m0 = blockIdx.y * TS, n0 = blockIdx.x * TS;
    stage_panel(sA, p.A + (int64_t)b * p.sA, p.lda, a_fm, m0, p.M, K, K4);
    stage_panel(sB, p.B + (int64_t)b * p.sB, p.ldb, b_fm, n0, p.N, K, K4);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    double acc[2][2][2];
    if (!panel::tile_mma(sA, sB, a_fm, b_fm, K, K4, psm, acc)) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = ((warp & 3) >> 1) * 16, wn = (warp & 1) * 16;
    double *D = p.D + (int64_t)b * p.sD;
    const double *E = p.E ? p.E + (int64_t)b * p.sE : nullptr;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int row = m0 + wm + i * 8 + g;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int col = n0 + wn + j * 8 + 2 * t + c;
                if (row >= p.M || col >= p.N) continue;
                double v = p.alpha * acc[i][j][c];
                if (E) v += p.beta * E[(int64_t)row * p.lde + col];
                if (row == col && row < p.eye_n) v += p.gamma;
                D[(int64_t)row * p.ldd + col] = v;
            }
    }
}

bool panel_eligible(const SmallArgs &p) {
    static const bool off = getenv("OO_OPT_NO_PANEL_GEMM") != nullptr;
    if (off || p.K > PK_MAX || p.K < 8 || (p.K & 1)) return false;
    if ((p.lda | p.ldb) & 1) return false;
    if (((p.sA | p.sB) & 1) != 0) return false;
    if ((reinterpret_cast<uintptr_t>(p.A) | reinterpret_cast<uintptr_t>(p.B)) & 15) return false;
    // rows along the free index are read in pairs (f, f+1): the pair must stay inside the row (even extent or
    // a padded leading dimension) -- guaranteed when the free extent is even
    if (p.transA && (p.M & 1)) return false;
    if (!p.transB && (p.N & 1)) return false;
    return true;
}
}  // namespace

int dgemm_small(int transA, int transB, int M, int N, int K, double alpha, const double *A, int lda,
                int64_t strideA, const double *B, int ldb, int64_t strideB, double beta,
                const double *E, int lde, int64_t strideE, double gamma, double *D, int ldd,
                int64_t strideD, int batch, cudaStream_t stream, int eye_n) {
    OO_REQUIRE(A && B && D);
    OO_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0);
    if (batch > 65535) return OO_ERR_UNSUPPORTED;
    SmallArgs p;
    p.A = A; p.B = B; p.E = E; p.D = D;
    p.M = M; p.N = N; p.K = K;
    p.lda = lda; p.ldb = ldb; p.lde = lde; p.ldd = ldd;
    p.sA = strideA; p.sB = strideB; p.sE = strideE; p.sD = strideD;
    p.alpha = alpha; p.beta = beta; p.gamma = gamma;
    p.transA = transA; p.transB = transB;
    p.eye_n = eye_n < 0 ? (M < N ? M : N) : eye_n;
    dim3 grid((unsigned)ceil_div(N, TS), (unsigned)ceil_div(M, TS), (unsigned)batch);
    if (panel_eligible(p)) {
        static unsigned long long configured = 0;
        if (once_per_device(configured)) {
            OO_CUDA_CHECK(cudaFuncSetAttribute(dgemm_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)panel::smem_bytes(PK_MAX, 1, 0)));     // both panels k-major: the largest
        }
        dgemm_panel_kernel<<<grid, PTHREADS, panel::smem_bytes(K, transA, transB), stream>>>(p);
        OO_LAUNCH_CHECK();
        return OO_OK;
    }
    dgemm_small_kernel<<<grid, 128, 0, stream>>>(p);
    OO_LAUNCH_CHECK();
    return OO_OK;
}

}  // namespace oo

extern "C" int oo_dgemm_small_f64(int transA, int transB, int M, int N, int K, double alpha,
                                  const double *A, int lda, int64_t strideA, const double *B, int ldb,
                                  int64_t strideB, double beta, const double *E, int lde,
                                  int64_t strideE, double gamma, double *D, int ldd, int64_t strideD,
                                  int batch, void *stream) {
    return oo::dgemm_small(transA, transB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, E,
                           lde, strideE, gamma, D, ldd, strideD, batch, (cudaStream_t)stream, -1);
}
