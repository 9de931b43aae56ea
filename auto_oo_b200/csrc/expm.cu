// K1: kappa -> skew matrix -> U = expm(-K) (reference oo_energy.py:213-230, :63-87),
// plus the N x N products around it: C' = X C_oao U (:173-176, :201, :235) and
// h' = C^T h C (:44-46).
//
// expm: scaling and squaring with the diagonal Pade-[7/7] approximant
// (Higham 2005, theta_7 = 0.95).  With A = -K / 2^s:
//   A2 = A A, A4 = A2 A2, A6 = A4 A2
//   W = A (b7 A6 + b5 A4 + b3 A2 + b1 I),  V = b6 A6 + b4 A4 + b2 A2 + b0 I
//   r(A) = (V - W)^{-1} (V + W),   U = r(A)^(2^s)
// For skew A, V is symmetric and W skew, so V - W = (V + W)^T is normal with
// eigenvalues ~ b0 exp(-i lambda / 2): perfectly conditioned and within
// 2 sin(0.95/4) = 0.47 of b0 I.  The inverse is therefore formed by the
// quadratically convergent Newton-Schulz iteration X <- X (2I - Q X) from
// X = I (residuals 0.47 -> 0.22 -> .049 -> 2.4e-3 -> 5.7e-6 -> 3.3e-11 -> 1e-21),
// which keeps the whole expm on batched DMMA GEMMs (dgemm_small.cu) with no
// pivoting and no host synchronisation.  Bases of up to 64 orbitals run the whole
// chain in ONE launch out of shared memory (expm_fused_kernel below).
#include <stdlib.h>

#include "common.cuh"

namespace oo {

int dgemm_small(int transA, int transB, int M, int N, int K, double alpha, const double *A, int lda,
                int64_t strideA, const double *B, int ldb, int64_t strideB, double beta,
                const double *E, int lde, int64_t strideE, double gamma, double *D, int ldd,
                int64_t strideD, int batch, cudaStream_t stream, int eye_n);

namespace {

// A[b] = scale * K(kappa[b]):  K[l,r] = +kappa_j, K[r,l] = -kappa_j
__global__ void skew_scatter_kernel(const double *__restrict__ kappa, const int32_t *__restrict__ pl,
                                    const int32_t *__restrict__ pr, int nk, int ld, double scale,
                                    double *__restrict__ A) {
    const int b = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nk) return;
    const double v = scale * kappa[(int64_t)b * nk + j];
    double *Ab = A + (int64_t)b * ld * ld;
    const int l = pl[j], r = pr[j];
    Ab[(int64_t)l * ld + r] = v;
    Ab[(int64_t)r * ld + l] = -v;
}

// out = c1 X1 + c2 X2 + c3 X3 + cI I   (batched ld x ld, contiguous)
__global__ void lincomb_kernel(double *__restrict__ out, double c1, const double *__restrict__ X1,
                               double c2, const double *__restrict__ X2, double c3,
                               const double *__restrict__ X3, double cI, int N, int ld, int64_t total) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t e = i % ((int64_t)ld * ld);
        const int row = (int)(e / ld), col = (int)(e % ld);
        double v = 0.0;
        if (X1) v += c1 * X1[i];
        if (X2) v += c2 * X2[i];
        if (X3) v += c3 * X3[i];
        if (row == col && row < N) v += cI;
        out[i] = v;
    }
}

int lincomb(double *out, double c1, const double *X1, double c2, const double *X2, double c3,
            const double *X3, double cI, int N, int ld, int batch, cudaStream_t stream) {
    const int64_t total = (int64_t)batch * ld * ld;
    int blocks = (int)ceil_div(total, 256);
    if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
    lincomb_kernel<<<blocks, 256, 0, stream>>>(out, c1, X1, c2, X2, c3, X3, cI, N, ld, total);
    OO_LAUNCH_CHECK();
    return OO_OK;
}

constexpr int kExpmSlots = 10;
constexpr int kNewtonSchulzIters = 5;

// expm of A (already scaled by 2^-s), in place workspace; result in U
int expm_scaled(const double *A, int N, int ld, int batch, int squarings, double *U, double *ws,
                cudaStream_t stream) {
    const int64_t mat = (int64_t)ld * ld;
    const int64_t sl = mat * batch;
    double *A2 = ws + 0 * sl, *A4 = ws + 1 * sl, *A6 = ws + 2 * sl, *W = ws + 3 * sl;
    double *V = ws + 4 * sl, *P = ws + 5 * sl, *Q = ws + 6 * sl, *X = ws + 7 * sl;
    double *T = ws + 8 * sl, *Y = ws + 9 * sl;
    // Pade-[7/7] coefficients normalised by b0 = 17297280
    const double b0 = 17297280.0;
    const double c1 = 8648640.0 / b0, c2 = 1995840.0 / b0, c3 = 277200.0 / b0, c4 = 25200.0 / b0,
                 c5 = 1512.0 / b0, c6 = 56.0 / b0, c7 = 1.0 / b0;
    int rc;
#define GEMM(a, b, alpha, e, beta, gam, d)                                                        \
    if ((rc = dgemm_small(0, 0, ld, ld, ld, (alpha), (a), ld, mat, (b), ld, mat, (beta), (e), ld,  \
                          mat, (gam), (d), ld, mat, batch, stream, N)))                           \
    return rc
    GEMM(A, A, 1.0, nullptr, 0.0, 0.0, A2);
    GEMM(A2, A2, 1.0, nullptr, 0.0, 0.0, A4);
    GEMM(A4, A2, 1.0, nullptr, 0.0, 0.0, A6);
    if ((rc = lincomb(W, c7, A6, c5, A4, c3, A2, c1, N, ld, batch, stream))) return rc;
    if ((rc = lincomb(V, c6, A6, c4, A4, c2, A2, 1.0, N, ld, batch, stream))) return rc;
    GEMM(A, W, 1.0, V, 1.0, 0.0, P);                                    // P = V + A W
    if ((rc = lincomb(Q, 2.0, V, -1.0, P, 0.0, nullptr, 0.0, N, ld, batch, stream))) return rc;  // Q = V - A W
    // Newton-Schulz: X1 = 2I - Q, then X <- X (2I - Q X)
    if ((rc = lincomb(X, -1.0, Q, 0.0, nullptr, 0.0, nullptr, 2.0, N, ld, batch, stream))) return rc;
    double *Xc = X, *Xn = Y;
    for (int it = 0; it < kNewtonSchulzIters; ++it) {
        GEMM(Q, Xc, -1.0, nullptr, 0.0, 2.0, T);                       // T = 2I - Q X
        GEMM(Xc, T, 1.0, nullptr, 0.0, 0.0, Xn);
        double *tmp = Xc; Xc = Xn; Xn = tmp;
    }
    // r = X P, then square; ping-pong so the last product lands in U
    double *Rc = (squarings % 2 == 0) ? U : T;
    double *Rn = (squarings % 2 == 0) ? T : U;
    GEMM(Xc, P, 1.0, nullptr, 0.0, 0.0, Rc);
    for (int s = 0; s < squarings; ++s) {
        GEMM(Rc, Rc, 1.0, nullptr, 0.0, 0.0, Rn);
        double *tmp = Rc; Rc = Rn; Rn = tmp;
    }
#undef GEMM
    return OO_OK;
}


// ---- fused expm for ld <= 64: one CTA per matrix, everything resident in shared memory -------------
// The unfused route above costs 21 + s launches whose kernels are a few microseconds each; for the
// small bases of the reference's molecules (7 ... 43 orbitals) the whole Pade / Newton-Schulz / squaring
// chain fits five n8 x n8 shared-memory slots (n8 = N rounded up to 8), so one launch does all of it:
// 16 warps, DMMA.8x8x4 straight from shared memory (row stride = 4 mod 16 doubles: conflict-free A and B
// fragment loads), same arithmetic in the same order as expm_scaled().
constexpr int kFusedMaxN8 = 64;
constexpr int kFusedThreads = 512;

__host__ __device__ inline int fused_row_stride(int n8) { return ((n8 + 11) / 16) * 16 + 4; }   // >= n8, = 4 mod 16

size_t fused_smem_bytes(int n8) { return (size_t)5 * n8 * fused_row_stride(n8) * sizeof(double); }

// D = alpha X Y + beta E + gamma I(rows < eyeN); D must not alias X, Y or E
template <int GR, int GC>
__device__ __forceinline__ void smem_gemm(double *__restrict__ D, const double *X, const double *Y, int n8,
                                          int LS, double alpha, const double *E, double beta, double gamma,
                                          int eyeN) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int nt = n8 >> 3;
    const int ngr = (nt + GR - 1) / GR, ngc = (nt + GC - 1) / GC;
    for (int grp = warp; grp < ngr * ngc; grp += nwarp) {
        const int r0 = (grp / ngc) * GR, c0 = (grp % ngc) * GC;
        double acc[GR][GC][2];
#pragma unroll
        for (int i = 0; i < GR; ++i)
#pragma unroll
            for (int j = 0; j < GC; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int k = 0; k < n8; k += 4) {
            double a[GR], b[GC];
#pragma unroll
            for (int i = 0; i < GR; ++i) a[i] = (r0 + i < nt) ? X[(8 * (r0 + i) + g) * LS + k + t] : 0.0;
#pragma unroll
            for (int j = 0; j < GC; ++j) b[j] = (c0 + j < nt) ? Y[(k + t) * LS + 8 * (c0 + j) + g] : 0.0;
#pragma unroll
            for (int i = 0; i < GR; ++i)
#pragma unroll
                for (int j = 0; j < GC; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
#pragma unroll
        for (int i = 0; i < GR; ++i) {
            if (r0 + i >= nt) continue;
            const int row = 8 * (r0 + i) + g;
#pragma unroll
            for (int j = 0; j < GC; ++j) {
                if (c0 + j >= nt) continue;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int col = 8 * (c0 + j) + 2 * t + c;
                    double v = alpha * acc[i][j][c];
                    if (E) v += beta * E[row * LS + col];
                    if (row == col && row < eyeN) v += gamma;
                    D[row * LS + col] = v;
                }
            }
        }
    }
    __syncthreads();
}

// out = c1 X1 + c2 X2 + c3 X3 + cI I(rows < N)   (elementwise; out may alias an input)
__device__ __forceinline__ void smem_lincomb(double *out, double c1, const double *X1, double c2, const double *X2,
                                             double c3, const double *X3, double cI, int N, int n8, int LS) {
    for (int e = threadIdx.x; e < n8 * n8; e += blockDim.x) {
        const int row = e / n8, col = e - row * n8, o = row * LS + col;
        double v = 0.0;
        if (X1) v += c1 * X1[o];
        if (X2) v += c2 * X2[o];
        if (X3) v += c3 * X3[o];
        if (row == col && row < N) v += cI;
        out[o] = v;
    }
    __syncthreads();
}

struct FusedExpmArgs {
    const double *kappa;          // (batch, nk) packed rotation parameters, or nullptr
    const int32_t *pl, *pr;
    int nk;
    const double *Ain;            // (batch, ld, ld) dense input when kappa == nullptr
    double scale;                 // A = scale * K(kappa)   resp.   A = scale * Ain
    int N, ld, n8, squarings;    // squarings < 0: chosen on the device from ||A||_1, per matrix
    double *U;                    // (batch, ld, ld)
};

template <int GR, int GC>
__global__ void __launch_bounds__(kFusedThreads) expm_fused_kernel(const FusedExpmArgs p) {
    extern __shared__ __align__(16) double sm[];
    const int n8 = p.n8, LS = fused_row_stride(n8), N = p.N;
    const int slot = n8 * LS;
    double *S0 = sm, *S1 = sm + slot, *S2 = sm + 2 * slot, *S3 = sm + 3 * slot, *S4 = sm + 4 * slot;
    const int b = blockIdx.x;
    // ---- A = scale * input into S0 (zero padded to n8)
    for (int e = threadIdx.x; e < slot; e += blockDim.x) S0[e] = 0.0;
    __syncthreads();
    if (p.kappa) {
        const double *kap = p.kappa + (int64_t)b * p.nk;
        for (int j = threadIdx.x; j < p.nk; j += blockDim.x) {
            const double v = p.scale * kap[j];
            const int l = p.pl[j], r = p.pr[j];
            S0[l * LS + r] = v;
            S0[r * LS + l] = -v;
        }
    } else {
        const double *Ab = p.Ain + (int64_t)b * p.ld * p.ld;
        for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
            const int row = e / N, col = e - row * N;
            S0[row * LS + col] = p.scale * Ab[(int64_t)row * p.ld + col];
        }
    }
    __syncthreads();
    int squarings = p.squarings;
    if (squarings < 0) {
        // device-side choice, per matrix: s = max(0, ceil(log2(||A||_1 / 0.95))), A <- A / 2^s (exact scaling);
        // the same rule the host applies (engine.squarings_for) without its device->host round trip
        double *colsum = S1;
        for (int col = threadIdx.x; col < n8; col += blockDim.x) {
            double acc = 0.0;
            for (int row = 0; row < n8; ++row) acc += fabs(S0[row * LS + col]);
            colsum[col] = acc;
        }
        __syncthreads();
        double norm1 = 0.0;
        for (int col = 0; col < n8; ++col) norm1 = fmax(norm1, colsum[col]);     // every thread: same order, same value
        squarings = 0;
        if (norm1 > 0.95 && norm1 < 1e300) squarings = min(64, (int)ceil(log2(norm1 / 0.95)));
        __syncthreads();
        if (squarings > 0) {
            const double sc = ldexp(1.0, -squarings);
            for (int e = threadIdx.x; e < slot; e += blockDim.x) S0[e] *= sc;
            __syncthreads();
        }
    }
    const double b0 = 17297280.0;
    const double c1 = 8648640.0 / b0, c2 = 1995840.0 / b0, c3 = 277200.0 / b0, c4 = 25200.0 / b0,
                 c5 = 1512.0 / b0, c6 = 56.0 / b0, c7 = 1.0 / b0;
    double *A = S0, *A2 = S1, *A4 = S2, *A6 = S3, *W = S4;
    smem_gemm<GR, GC>(A2, A, A, n8, LS, 1.0, nullptr, 0.0, 0.0, 0);
    smem_gemm<GR, GC>(A4, A2, A2, n8, LS, 1.0, nullptr, 0.0, 0.0, 0);
    smem_gemm<GR, GC>(A6, A4, A2, n8, LS, 1.0, nullptr, 0.0, 0.0, 0);
    smem_lincomb(W, c7, A6, c5, A4, c3, A2, c1, N, n8, LS);
    double *V = S3;
    smem_lincomb(V, c6, A6, c4, A4, c2, A2, 1.0, N, n8, LS);                 // in place over A6
    double *P = S1;
    smem_gemm<GR, GC>(P, A, W, n8, LS, 1.0, V, 1.0, 0.0, 0);                  // P = V + A W   (A2 dead)
    double *Q = S2;
    smem_lincomb(Q, 2.0, V, -1.0, P, 0.0, nullptr, 0.0, N, n8, LS);          // Q = V - A W   (A4 dead)
    double *Xc = S4, *Xn = S3, *T = S0;
    smem_lincomb(Xc, -1.0, Q, 0.0, nullptr, 0.0, nullptr, 2.0, N, n8, LS);   // X1 = 2I - Q    (W dead)
    for (int it = 0; it < kNewtonSchulzIters; ++it) {
        smem_gemm<GR, GC>(T, Q, Xc, n8, LS, -1.0, nullptr, 0.0, 2.0, N);     // T = 2I - Q X   (A, V dead)
        smem_gemm<GR, GC>(Xn, Xc, T, n8, LS, 1.0, nullptr, 0.0, 0.0, 0);
        double *tmp = Xc; Xc = Xn; Xn = tmp;
    }
    double *Rc = S0, *Rn = S2;                                              // T and Q are dead now
    smem_gemm<GR, GC>(Rc, Xc, P, n8, LS, 1.0, nullptr, 0.0, 0.0, 0);
    for (int s = 0; s < squarings; ++s) {
        smem_gemm<GR, GC>(Rn, Rc, Rc, n8, LS, 1.0, nullptr, 0.0, 0.0, 0);
        double *tmp = Rc; Rc = Rn; Rn = tmp;
    }
    double *Ub = p.U + (int64_t)b * p.ld * p.ld;
    for (int e = threadIdx.x; e < p.ld * p.ld; e += blockDim.x) {
        const int row = e / p.ld, col = e - row * p.ld;
        Ub[e] = Rc[row * LS + col];
    }
}

bool expm_fused_enabled() {
    static const bool on = getenv("OO_OPT_EXPM_UNFUSED") == nullptr;
    return on;
}

template <int GR, int GC>
int launch_expm_fused(const FusedExpmArgs &p, int batch, cudaStream_t stream) {
    static unsigned long long configured = 0;
    const size_t smem = fused_smem_bytes(p.n8);
    if (once_per_device(configured)) {
        OO_CUDA_CHECK(cudaFuncSetAttribute(expm_fused_kernel<GR, GC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)fused_smem_bytes(kFusedMaxN8)));
    }
    expm_fused_kernel<GR, GC><<<batch, kFusedThreads, smem, stream>>>(p);
    OO_LAUNCH_CHECK();
    return OO_OK;
}

int expm_fused(FusedExpmArgs p, int batch, cudaStream_t stream) {
    p.n8 = (p.N + 7) / 8 * 8;
    const int nt = p.n8 / 8;
    if (nt <= 4) return launch_expm_fused<1, 1>(p, batch, stream);
    if (nt == 5) return launch_expm_fused<1, 2>(p, batch, stream);
    if (nt == 6) return launch_expm_fused<1, 3>(p, batch, stream);
    return launch_expm_fused<2, 2>(p, batch, stream);
}

}  // namespace

size_t rotation_ws_bytes(int ld, int batch) {
    return (size_t)(kExpmSlots + 1) * batch * ld * ld * sizeof(double);
}

int kappa_rotation(const double *kappa, const int32_t *pl, const int32_t *pr, int nk, int N, int ld,
                   int batch, int squarings, double *U, void *ws, size_t ws_bytes,
                   cudaStream_t stream) {
    OO_REQUIRE(kappa && U && ws && (nk == 0 || (pl && pr)));
    OO_REQUIRE(N > 0 && ld >= N && (ld % 2) == 0 && batch > 0 && squarings >= -1 && squarings <= 64);
    if (ws_bytes < rotation_ws_bytes(ld, batch)) return OO_ERR_WORKSPACE;
    if (batch > 65535) return OO_ERR_UNSUPPORTED;
    if (N <= kFusedMaxN8 && expm_fused_enabled()) {
        FusedExpmArgs p{kappa, pl, pr, nk, nullptr, squarings < 0 ? -1.0 : -ldexp(1.0, -squarings), N, ld, 0,
                        squarings, U};
        return expm_fused(p, batch, stream);
    }
    if (squarings < 0) return OO_ERR_UNSUPPORTED;      // the multi-launch route needs the host's choice
    double *w = reinterpret_cast<double *>(ws);
    const int64_t sl = (int64_t)batch * ld * ld;
    double *A = w + (int64_t)kExpmSlots * sl;
    OO_CUDA_CHECK(cudaMemsetAsync(A, 0, sl * sizeof(double), stream));
    if (nk > 0) {
        dim3 grid((unsigned)ceil_div(nk, 256), (unsigned)batch);
        // A = -K / 2^s
        skew_scatter_kernel<<<grid, 256, 0, stream>>>(kappa, pl, pr, nk, ld, -ldexp(1.0, -squarings), A);
        OO_LAUNCH_CHECK();
    }
    return expm_scaled(A, N, ld, batch, squarings, U, w, stream);
}

int expm_general(const double *Ain, double sign, int N, int ld, int batch, int squarings, double *U,
                 void *ws, size_t ws_bytes, cudaStream_t stream) {
    OO_REQUIRE(Ain && U && ws);
    OO_REQUIRE(N > 0 && ld >= N && (ld % 2) == 0 && batch > 0 && squarings >= 0 && squarings <= 64);
    if (ws_bytes < rotation_ws_bytes(ld, batch)) return OO_ERR_WORKSPACE;
    if (N <= kFusedMaxN8 && batch <= 65535 && expm_fused_enabled()) {
        FusedExpmArgs p{nullptr, nullptr, nullptr, 0, Ain, sign * ldexp(1.0, -squarings), N, ld, 0, squarings, U};
        return expm_fused(p, batch, stream);
    }
    double *w = reinterpret_cast<double *>(ws);
    const int64_t sl = (int64_t)batch * ld * ld;
    double *A = w + (int64_t)kExpmSlots * sl;
    int rc = lincomb(A, sign * ldexp(1.0, -squarings), Ain, 0.0, nullptr, 0.0, nullptr, 0.0, N, ld,
                     batch, stream);
    if (rc) return rc;
    return expm_scaled(A, N, ld, batch, squarings, U, w, stream);
}

size_t int1e_ws_bytes(int ld, int batch) { return (size_t)batch * ld * ld * sizeof(double); }

int mo_coeff(const double *X, int64_t strideX, const double *Coao, int64_t strideCoao, const double *U,
             int64_t strideU, int N, int ld, int batch, double *Cout, void *ws, size_t ws_bytes,
             cudaStream_t stream) {
    OO_REQUIRE(X && Coao && Cout);
    OO_REQUIRE(N > 0 && ld >= N && batch > 0);
    const int64_t mat = (int64_t)ld * ld;
    if (!U)
        return dgemm_small(0, 0, ld, ld, ld, 1.0, X, ld, strideX, Coao, ld, strideCoao, 0.0, nullptr, ld, 0, 0.0,
                           Cout, ld, mat, batch, stream, 0);
    OO_REQUIRE(ws);
    if (ws_bytes < int1e_ws_bytes(ld, batch)) return OO_ERR_WORKSPACE;
    double *T = reinterpret_cast<double *>(ws);
    int rc = dgemm_small(0, 0, ld, ld, ld, 1.0, Coao, ld, strideCoao, U, ld, strideU, 0.0, nullptr, ld, 0,
                         0.0, T, ld, mat, batch, stream, 0);
    if (rc) return rc;
    return dgemm_small(0, 0, ld, ld, ld, 1.0, X, ld, strideX, T, ld, mat, 0.0, nullptr, ld, 0, 0.0, Cout, ld,
                       mat, batch, stream, 0);
}

int int1e_transform(const double *h, int64_t stride_h, const double *C, int64_t strideC, int N, int ld,
                    int batch, double *hmo, void *ws, size_t ws_bytes, cudaStream_t stream) {
    OO_REQUIRE(h && C && hmo && ws);
    OO_REQUIRE(N > 0 && ld >= N && batch > 0);
    if (ws_bytes < int1e_ws_bytes(ld, batch)) return OO_ERR_WORKSPACE;
    const int64_t mat = (int64_t)ld * ld;
    double *T = reinterpret_cast<double *>(ws);
    // T = h C ; h' = C^T T
    int rc = dgemm_small(0, 0, ld, ld, ld, 1.0, h, ld, stride_h, C, ld, strideC, 0.0, nullptr, ld, 0, 0.0, T, ld,
                         mat, batch, stream, 0);
    if (rc) return rc;
    return dgemm_small(1, 0, ld, ld, ld, 1.0, C, ld, strideC, T, ld, mat, 0.0, nullptr, ld, 0, 0.0, hmo,
                       ld, mat, batch, stream, 0);
}

}  // namespace oo

extern "C" {

int oo_expm_device_squarings_max_n(void) { return oo::expm_fused_enabled() ? oo::kFusedMaxN8 : 0; }

int oo_kappa_rotation_f64(const double *kappa, const int32_t *pair_l, const int32_t *pair_r, int nk,
                          int N, int ld, int batch, int squarings, double *U, void *ws,
                          size_t ws_bytes, void *stream) {
    return oo::kappa_rotation(kappa, pair_l, pair_r, nk, N, ld, batch, squarings, U, ws, ws_bytes,
                              (cudaStream_t)stream);
}

int oo_expm_f64(const double *A, double sign, int N, int ld, int batch, int squarings, double *U,
                void *ws, size_t ws_bytes, void *stream) {
    return oo::expm_general(A, sign, N, ld, batch, squarings, U, ws, ws_bytes, (cudaStream_t)stream);
}

int oo_mo_coeff_f64(const double *X, int64_t strideX, const double *Coao, int64_t strideCoao, const double *U,
                    int64_t strideU, int N, int ld, int batch, double *Cout, void *ws, size_t ws_bytes,
                    void *stream) {
    return oo::mo_coeff(X, strideX, Coao, strideCoao, U, strideU, N, ld, batch, Cout, ws, ws_bytes,
                        (cudaStream_t)stream);
}

int oo_int1e_transform_f64(const double *h_ao, int64_t stride_h, const double *C, int64_t strideC, int N,
                           int ld, int batch, double *h_mo, void *ws, size_t ws_bytes, void *stream) {
    return oo::int1e_transform(h_ao, stride_h, C, strideC, N, ld, batch, h_mo, ws, ws_bytes,
                               (cudaStream_t)stream);
}
}
