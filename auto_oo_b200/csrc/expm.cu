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
// pivoting and no host synchronisation.
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

}  // namespace

size_t rotation_ws_bytes(int ld, int batch) {
    return (size_t)(kExpmSlots + 1) * batch * ld * ld * sizeof(double);
}

int kappa_rotation(const double *kappa, const int32_t *pl, const int32_t *pr, int nk, int N, int ld,
                   int batch, int squarings, double *U, void *ws, size_t ws_bytes,
                   cudaStream_t stream) {
    OO_REQUIRE(kappa && U && ws && (nk == 0 || (pl && pr)));
    OO_REQUIRE(N > 0 && ld >= N && (ld % 2) == 0 && batch > 0 && squarings >= 0 && squarings <= 64);
    if (ws_bytes < rotation_ws_bytes(ld, batch)) return OO_ERR_WORKSPACE;
    if (batch > 65535) return OO_ERR_UNSUPPORTED;
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
