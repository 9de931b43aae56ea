// Partial ("class") integral transformation: only the two integral classes the energy, the
// orbital gradient and the I-space Hessian read,
//   J[m,n,a,b] = g'[a,b,m,n] = (ab|mn)      K[n,m,a,b] = g'[a,m,n,b] = (am|nb)
// with m, n in I = occ + act (padded to nIp) and a, b general, instead of the complete
// g'[i,j,k,l] of general_4index_transform (reference oo_energy.py:21-30).  The numbers every
// consumer reads are the same sums in the same association as the full transform followed by
// slicing; the cost drops from 8 N^5 to 2 N^4 nI + 12 N^3 nI^2 flop.
//
// Every step is the rotating TN-GEMM of dgemm_tn.cu (the contracted LEADING index leaves, its
// image is appended at the back), starting from the pair-transposed AO tensor
// gp[r,s,p,q] = g[p,q,r,s] (built once per problem by oo_transpose_f64; no symmetry of g is assumed):
//   Q1            T1[s,p,q,m]  = sum_r gp[r,(s p q)] C[r,m]        2 N^4 nI flop   (m < nIp)
//   J:  Q2        X[p,q,m,n]   = sum_s T1[s,(p q m)] C[s,n]
//       Q3        X'[q,m,n,a]  = sum_p X[p,(q m n)]  C[p,a]
//       Q4        J[m,n,a,b]   = sum_q X'[q,(m n a)] C[q,b]
//   K:  swap      T1t[q,p,s,m] = T1[s,p,q,m]                       (second store in Q1's epilogue)
//       K2        X[p,s,n,m]   = sum_q T1t[q,(p s n)] C[q,m]
//       K3        X'[s,n,m,a]  = sum_p X[p,(s n m)]   C[p,a]
//       K4        K[n,m,a,b]   = sum_s X'[s,(n m a)]  C[s,b]
// The results land in ONE buffer  cls = [K rows (nIp^2) ; J rows (nIp^2) ; h' row], each row ld^2,
// which is exactly the B operand of the Hessian's T-matrix GEMM (hessian.cu) -- no gather pass.
#include "common.cuh"

namespace oo {

int dgemm_tn(const double *At, const double *B, double *C, int64_t M, int64_t N, int64_t K,
             int64_t lda, int64_t ldb, int64_t ldc, int batch, int64_t strideA, int64_t strideB,
             int64_t strideC, cudaStream_t stream);
int dgemm_tn_swap02(const double *At, const double *B, double *C, double *C2, int d0, int d1, int d2,
                    int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int batch, int64_t strideA,
                    int64_t strideB, int64_t strideC, cudaStream_t stream);

namespace {

// dst[c][r] = src[r][c]  (rows x cols doubles), 32x32 tiles through padded shared memory
__global__ void __launch_bounds__(256) transpose_kernel(const double *__restrict__ src,
                                                        double *__restrict__ dst, int64_t rows,
                                                        int64_t cols) {
    __shared__ double tile[32][33];
    const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int64_t r = r0 + j, c = c0 + tx;
        if (r < rows && c < cols) tile[j][tx] = src[r * cols + c];
    }
    __syncthreads();
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int64_t c = c0 + j, r = r0 + tx;
        if (r < rows && c < cols) dst[c * rows + r] = tile[tx][j];
    }
}

}  // namespace

int transpose(const double *src, double *dst, int64_t rows, int64_t cols, cudaStream_t stream) {
    OO_REQUIRE(src && dst && rows > 0 && cols > 0 && src != dst);
    const int64_t gx = ceil_div(cols, 32), gy = ceil_div(rows, 32);
    if (gx > 0x7fffffffll || gy > 65535) return OO_ERR_UNSUPPORTED;
    transpose_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, stream>>>(src, dst, rows, cols);
    OO_LAUNCH_CHECK();
    return OO_OK;
}

size_t class_transform_ws_bytes(int ld, int nIp, int batch) {
    const size_t ld2 = (size_t)ld * ld;
    return (size_t)batch * (2 * ld2 * ld * nIp + 2 * ld2 * nIp * nIp) * sizeof(double);
}

size_t class_buffer_bytes(int ld, int nIp) {
    return (size_t)(2 * (size_t)nIp * nIp + 1) * ld * ld * sizeof(double);
}

// batch evaluations at once: C[b] (strideC, 0 = shared), gp[b] (strideG, 0 = shared AO integrals),
// cls[b] contiguous class buffers.  Every step is one batched TN-GEMM launch.
int class_transform(const double *gp, int64_t strideG, const double *C, int64_t strideC, int N, int ld,
                    int nIp, int batch, double *cls, void *ws, size_t ws_bytes, cudaStream_t stream) {
    OO_REQUIRE(gp && C && cls && ws);
    OO_REQUIRE(N > 0 && ld >= N && (ld % 2) == 0 && nIp > 0 && (nIp % 2) == 0 && nIp <= ld && batch > 0);
    if (ws_bytes < class_transform_ws_bytes(ld, nIp, batch)) return OO_ERR_WORKSPACE;
    if (ld > 65535) return OO_ERR_UNSUPPORTED;
    const int64_t ld2 = (int64_t)ld * ld, ld3 = ld2 * ld, nI2 = (int64_t)nIp * nIp;
    const int64_t sT1 = ld3 * nIp, sX = ld2 * nI2, sCls = (2 * nI2 + 1) * ld2;
    double *T1 = reinterpret_cast<double *>(ws);
    double *T1t = T1 + (int64_t)batch * sT1;
    double *X = T1t + (int64_t)batch * sT1;
    double *Xp = X + (int64_t)batch * sX;
    double *Kout = cls, *Jout = cls + nI2 * ld2;
    int rc;
#define Q(in, sIn, out, sOut, M, Ncols)                                                                  \
    if ((rc = dgemm_tn((in), C, (out), (M), (Ncols), ld, (M), ld, (Ncols), batch, (sIn), strideC, (sOut), \
                       stream)))                                                                         \
    return rc
    // Q1 writes T1[s,p,q,m] and its (s <-> q) swapped copy T1t[q,p,s,m] from the same accumulators
    if ((rc = dgemm_tn_swap02(gp, C, T1, T1t, ld, ld, ld, nIp, ld, ld3, ld, nIp, batch, strideG, strideC, sT1,
                              stream)))
        return rc;
    Q(T1, sT1, X, sX, ld2 * nIp, nIp);         // [p,q,m,n]
    Q(X, sX, Xp, sX, ld * nI2, ld);            // [q,m,n,a]
    Q(Xp, sX, Jout, sCls, nI2 * ld, ld);       // [m,n,a,b]
    Q(T1t, sT1, X, sX, ld2 * nIp, nIp);        // [p,s,n,m]
    Q(X, sX, Xp, sX, ld * nI2, ld);            // [s,n,m,a]
    Q(Xp, sX, Kout, sCls, nI2 * ld, ld);       // [n,m,a,b]
#undef Q
    return OO_OK;
}

}  // namespace oo

extern "C" {

int oo_transpose_f64(const double *src, double *dst, int64_t rows, int64_t cols, void *stream) {
    return oo::transpose(src, dst, rows, cols, (cudaStream_t)stream);
}

int oo_class_transform_f64(const double *g_pairT, int64_t strideG, const double *C, int64_t strideC, int N,
                           int ld, int nIp, int batch, double *cls, void *ws, size_t ws_bytes, void *stream) {
    return oo::class_transform(g_pairT, strideG, C, strideC, N, ld, nIp, batch, cls, ws, ws_bytes,
                               (cudaStream_t)stream);
}
}
