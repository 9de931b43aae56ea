// K1: kappa -> skew matrix -> U = expm(-K) (reference oo_energy.py:213-230, :63-87),
// plus the N x N products around it: C' = X C_oao U (:173-176, :201, :235) and
// h' = C^T h C (:44-46).
//
// expm: scaling and squaring around a degree-18 Taylor polynomial.  With A = -K / 2^s, ||A||_1 <= 0.95, the
// truncation error 0.95^19 / 19! = 3e-18 is below the unit round-off, and for skew A every term is bounded by
// ||A||^k / k! (exp(A) is orthogonal: no cancellation to lose digits to).  The polynomial is evaluated by
// Paterson-Stockmeyer in blocks of four,
//   A2 = A A, A3 = A2 A, A4 = A2 A2                                            (3 products)
//   B_j = c_4j I + c_4j+1 A + c_4j+2 A2 + c_4j+3 A3   (j = 0..3),  B_4 = c_16 I + c_17 A + c_18 A2,   c_k = 1/k!
//   R = B_4;  R <- R A4 + B_j  for j = 3, 2, 1, 0                              (4 products, B_j added in the epilogue)
//   U = R^(2^s)                                                                (s products)
// i.e. 7 + s DMMA GEMMs with no inverse, no pivoting and no host synchronisation (the first version used the
// Pade-[7/7] approximant with a Newton-Schulz inverse: 17 + s products for the same accuracy).  Bases of up to 64
// orbitals run the whole chain in ONE launch out of shared memory (expm_fused_kernel below).
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
constexpr int kTaylorDegree = 18;

// c_k = 1 / k!
__host__ __device__ inline double inv_factorial(int k) {
    double f = 1.0;
    for (int i = 2; i <= k; ++i) f *= (double)i;
    return 1.0 / f;
}

// B_j = c_4j I + c_4j+1 A + c_4j+2 A2 + c_4j+3 A3 for j = 0..4 (terms past the degree dropped), all in one pass
__global__ void taylor_blocks_kernel(const double *__restrict__ A, const double *__restrict__ A2,
                                     const double *__restrict__ A3, double *__restrict__ B, int64_t slot, int N, int ld,
                                     int64_t total) {
    double c[kTaylorDegree + 1];
#pragma unroll
    for (int k = 0; k <= kTaylorDegree; ++k) c[k] = inv_factorial(k);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t e = i % ((int64_t)ld * ld);
        const int row = (int)(e / ld), col = (int)(e % ld);
        const double a1 = A[i], a2 = A2[i], a3 = A3[i];
        const double eye = (row == col && row < N) ? 1.0 : 0.0;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            double v = c[4 * j] * eye + c[4 * j + 1] * a1 + c[4 * j + 2] * a2;
            if (4 * j + 3 <= kTaylorDegree) v += c[4 * j + 3] * a3;
            B[j * slot + i] = v;
        }
    }
}

// expm of A (already scaled by 2^-s), in place workspace; result in U
int expm_scaled(const double *A, int N, int ld, int batch, int squarings, double *U, double *ws,
                cudaStream_t stream) {
    const int64_t mat = (int64_t)ld * ld;
    const int64_t sl = mat * batch;
    double *A2 = ws + 0 * sl, *A3 = ws + 1 * sl, *A4 = ws + 2 * sl, *B = ws + 3 * sl;      // B_0 .. B_4: slots 3-7
    double *T0 = ws + 8 * sl, *T1 = ws + 9 * sl;
    int rc;
#define GEMM(a, b, e, beta, d)                                                                          \
    if ((rc = dgemm_small(0, 0, ld, ld, ld, 1.0, (a), ld, mat, (b), ld, mat, (beta), (e), ld, mat, 0.0, \
                          (d), ld, mat, batch, stream, N)))                                             \
    return rc
    GEMM(A, A, nullptr, 0.0, A2);
    GEMM(A2, A, nullptr, 0.0, A3);
    GEMM(A2, A2, nullptr, 0.0, A4);
    {
        int blocks = (int)ceil_div(sl, 256);
        if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
        taylor_blocks_kernel<<<blocks, 256, 0, stream>>>(A, A2, A3, B, sl, N, ld, sl);
        OO_LAUNCH_CHECK();
    }
    // Horner in A4; the chain and the squarings ping-pong so that the last product lands in U
    const int products = 4 + squarings;
    double *Rc = B + 4 * sl;
    for (int i = 0; i < products; ++i) {
        double *Rn = (i == products - 1) ? U : ((i % 2 == 0) ? T0 : T1);
        if (i < 4) {
            GEMM(Rc, A4, B + (3 - i) * sl, 1.0, Rn);                    // R <- R A4 + B_(3-i)
        } else {
            GEMM(Rc, Rc, nullptr, 0.0, Rn);
        }
        Rc = Rn;
    }
#undef GEMM
    return OO_OK;
}


// ---- fused expm for ld <= 64: one CTA per matrix, everything resident in shared memory -------------
// The unfused route above costs 21 + s launches whose kernels are a few microseconds each; for the
// small bases of the reference's molecules (7 ... 43 orbitals) the whole Taylor / squaring
// chain fits six n8 x n8 shared-memory slots (n8 = N rounded up to 8; 209 KB at n8 = 64), so one launch does all of it:
// 16 warps, DMMA.8x8x4 straight from shared memory (row stride = 4 mod 16 doubles: conflict-free A and B
// fragment loads), same arithmetic in the same order as expm_scaled().
constexpr int kFusedMaxN8 = 64;
constexpr int kFusedThreads = 512;

__host__ __device__ inline int fused_row_stride(int n8) { return ((n8 + 11) / 16) * 16 + 4; }   // >= n8, = 4 mod 16

size_t fused_smem_bytes(int n8) { return (size_t)6 * n8 * fused_row_stride(n8) * sizeof(double); }

// D = alpha X Y + beta E + gamma I(rows < eyeN); D must not alias X, Y or E
// P1, P2, P3 (optional): the epilogue also adds p1 P1 + p2 P2 + p3 P3 -- a Taylor block B_j formed on the fly
template <int GR, int GC>
__device__ __forceinline__ void smem_gemm(double *__restrict__ D, const double *X, const double *Y, int n8,
                                          int LS, double alpha, const double *E, double beta, double gamma,
                                          int eyeN, const double *P1 = nullptr, double p1 = 0.0,
                                          const double *P2 = nullptr, double p2 = 0.0, const double *P3 = nullptr,
                                          double p3 = 0.0) {
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
                    if (P1) v += p1 * P1[row * LS + col];
                    if (P2) v += p2 * P2[row * LS + col];
                    if (P3) v += p3 * P3[row * LS + col];
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
    double *S5 = sm + 5 * slot;
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
    // degree-18 Taylor polynomial by Paterson-Stockmeyer (same arithmetic, same order as expm_scaled())
    double c[kTaylorDegree + 1];
#pragma unroll
    for (int k = 0; k <= kTaylorDegree; ++k) c[k] = inv_factorial(k);
    double *A = S0, *A2 = S1, *A3 = S2, *A4 = S3;
    smem_gemm<GR, GC>(A2, A, A, n8, LS, 1.0, nullptr, 0.0, 0.0, 0);
    smem_gemm<GR, GC>(A3, A2, A, n8, LS, 1.0, nullptr, 0.0, 0.0, 0);
    smem_gemm<GR, GC>(A4, A2, A2, n8, LS, 1.0, nullptr, 0.0, 0.0, 0);
    double *Rc = S4, *Rn = S5;
    smem_lincomb(Rc, c[17], A, c[18], A2, 0.0, nullptr, c[16], N, n8, LS);               // B_4
#pragma unroll
    for (int j = 3; j >= 0; --j) {                                                        // R <- R A4 + B_j
        smem_gemm<GR, GC>(Rn, Rc, A4, n8, LS, 1.0, nullptr, 0.0, c[4 * j], N, A, c[4 * j + 1], A2, c[4 * j + 2], A3,
                          c[4 * j + 3]);
        double *tmp = Rc; Rc = Rn; Rn = tmp;
    }
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
