// K3 / K4-gradient: the RDM contractions of the hot path.  All are HBM/L2-bound
// gather-reduce kernels over slices of the transformed integrals g' (ld^4):
//   active_hamiltonian : c0, c1, c2      (reference utils/active_space.py:111-212)
//   energy             : E = c0 + <c1,gamma> + <c2,Gamma>      (oo_energy.py:194-197)
//   fock_core_active   : F^I, F^A                              (oo_energy.py:272-298)
//   fock_general       : generalized Fock                      (oo_energy.py:238-270)
//   gradient           : G = 2 (F - F^T) and its packed vector (oo_energy.py:300-309, :221-224)
//   fock_gradient_vjp  : adjoint w.r.t. the RDMs (what autograd provides in oo_pqc.py:113-123)
// Index patterns follow the reference exactly (no use of the 8-fold symmetry of g),
// so results agree for arbitrary input tensors.  Reductions are warp-shuffle trees
// in a fixed order (deterministic).
#include "common.cuh"

namespace oo {
namespace {

// ---------------------------------------------------------------- active Hamiltonian
// grid (1 + na*na + ceil(na^4/256), batch).  block 0: c0; blocks 1..na^2: c1[t,u]; rest: c2.
template <class GV>
__global__ void __launch_bounds__(256)
active_hamiltonian_kernel(const double *__restrict__ h, int64_t h_stride, const GV gv, int no, int na,
                          int ld, double e_nuc, const double *__restrict__ e_nuc_b, double *__restrict__ c0,
                          double *__restrict__ c1,
                          double *__restrict__ c2) {
    __shared__ double scratch[32];
    const int b = blockIdx.y;
    const double *hb = h + (int64_t)b * h_stride;
    const GV gb = gv.at(b);
    const int na2 = na * na;
    const int64_t na4 = (int64_t)na2 * na2;
    const int blk = blockIdx.x;
    if (blk == 0) {
        // c0 = e_nuc + 2 sum_i h_ii + sum_ij (2 g_iijj - g_ijji)
        double s = 0.0;
        for (int e = threadIdx.x; e < no * no; e += blockDim.x) {
            const int i = e / no, j = e % no;
            s += 2.0 * gb.coul(i, i, j, j) - gb.exch(i, j, j, i);
            if (j == 0) s += 2.0 * hb[(int64_t)i * ld + i];
        }
        s = block_sum(s, scratch);
        if (threadIdx.x == 0) c0[b] = s + (e_nuc_b ? e_nuc_b[b] : e_nuc);
    } else if (blk <= na2) {
        const int t = (blk - 1) / na, u = (blk - 1) % na;
        const int T = no + t, U = no + u;
        double s = 0.0;
        for (int i = threadIdx.x; i < no; i += blockDim.x)
            s += 2.0 * gb.coul(T, U, i, i) - gb.exch(T, i, i, U);
        s = block_sum(s, scratch);
        if (threadIdx.x == 0) c1[(int64_t)b * na2 + t * na + u] = s + hb[(int64_t)T * ld + U];
    } else {
        const int64_t e = (int64_t)(blk - 1 - na2) * blockDim.x + threadIdx.x;
        if (e < na4) {
            const int w = (int)(e % na), v = (int)((e / na) % na), u = (int)((e / na2) % na),
                      t = (int)(e / ((int64_t)na2 * na));
            c2[(int64_t)b * na4 + e] = 0.5 * gb.coul(no + t, no + u, no + v, no + w);
        }
    }
}

// ---------------------------------------------------------------- energy
__global__ void __launch_bounds__(1024)
energy_kernel(const double *__restrict__ c0, const double *__restrict__ c1, const double *__restrict__ c2,
              const double *__restrict__ d1, int64_t sd1, const double *__restrict__ d2, int64_t sd2,
              int na, double *__restrict__ E) {
    __shared__ double scratch[32];
    const int b = blockIdx.x;
    const int na2 = na * na;
    const int64_t na4 = (int64_t)na2 * na2;
    const double *c1b = c1 + (int64_t)b * na2, *c2b = c2 + (int64_t)b * na4;
    const double *d1b = d1 + (int64_t)b * sd1, *d2b = d2 + (int64_t)b * sd2;
    double s1 = 0.0, s2 = 0.0;
    for (int e = threadIdx.x; e < na2; e += blockDim.x) s1 += c1b[e] * d1b[e];
    for (int64_t e = threadIdx.x; e < na4; e += blockDim.x) s2 += c2b[e] * d2b[e];
    s1 = block_sum(s1, scratch);
    s2 = block_sum(s2, scratch);
    if (threadIdx.x == 0) E[b] = (c0[b] + s1) + s2;   // same association as sum((c0, e1, e2))
}

// ---------------------------------------------------------------- F^I and F^A
// one warp per (m, n); lanes split the i / (v,w) sums.  grid (ceil(ld*ld/8), batch), 256 threads.
template <class GV>
__global__ void __launch_bounds__(256)
fock_core_active_kernel(const double *__restrict__ h, int64_t h_stride, const GV gv,
                        const double *__restrict__ d1, int64_t sd1, int no, int na, int N, int ld,
                        double *__restrict__ FI, double *__restrict__ FA) {
    extern __shared__ double s_d1[];   // gamma (na*na)
    const int b = blockIdx.y;
    const double *d1b = d1 + (int64_t)b * sd1;
    for (int e = threadIdx.x; e < na * na; e += blockDim.x) s_d1[e] = d1b[e];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pair >= ld * ld) return;
    const int m = pair / ld, n = pair % ld;
    const int64_t mat = (int64_t)ld * ld;
    double fi = 0.0, fa = 0.0;
    if (m < N && n < N) {
        const GV gb = gv.at(b);
        for (int i = lane; i < no; i += 32) fi += 2.0 * gb.coul(m, n, i, i) - gb.exch(m, i, i, n);
        for (int e = lane; e < na * na; e += 32) {
            const int v = e / na, w = e % na;
            fa += s_d1[e] * (gb.coul(m, n, no + v, no + w) - 0.5 * gb.exch(m, no + w, no + v, n));
        }
        fi = warp_sum(fi);
        fa = warp_sum(fa);
        fi += h[(int64_t)b * h_stride + (int64_t)m * ld + n];
    }
    if (lane == 0) {
        FI[(int64_t)b * mat + pair] = fi;
        if (FA) FA[(int64_t)b * mat + pair] = fa;
    }
}

// Class-buffer form of the same two matrices: every J / K row of the class buffer is an (m, n) plane, so a thread
// owns two consecutive (m, n) elements and streams down the 2 no + 2 na^2 rows it needs (coalesced 16-byte loads,
// four rows in flight) instead of one strided 8-byte load per lane:
//   F^I = h + sum_i (2 J[(i i)] - K[(i i)]),   F^A = sum_vw gamma_vw (J[(v w)] - K[(v w)] / 2)
constexpr int kFockCols = 64;      // element pairs per CTA; 256 / kFockCols row groups share the row loop

__global__ void __launch_bounds__(256)
fock_core_active_class_kernel(const double *__restrict__ h, int64_t h_stride, const ClassView gv,
                              const double *__restrict__ d1, int64_t sd1, int no, int na, int N, int ld,
                              double *__restrict__ FI, double *__restrict__ FA) {
    extern __shared__ double s_d1[];   // gamma (na*na), then the partial sums of the row groups
    constexpr int kGroups = 256 / kFockCols;
    double2 *red = reinterpret_cast<double2 *>(s_d1 + ((na * na + 1) & ~1));      // [2][kGroups][kFockCols]
    const int b = blockIdx.y;
    const double *d1b = d1 + (int64_t)b * sd1;
    for (int e = threadIdx.x; e < na * na; e += blockDim.x) s_d1[e] = d1b[e];
    __syncthreads();
    const int64_t mat = (int64_t)ld * ld;
    const int tx = threadIdx.x % kFockCols, ty = threadIdx.x / kFockCols;
    const int64_t e0 = ((int64_t)blockIdx.x * kFockCols + tx) * 2;
    const bool live = e0 < mat;
    const ClassView gb = gv.at(b);
    const int nIp = gb.nIp;
    const double *J = gb.J + e0, *K = gb.K + e0;
    double2 fi = make_double2(0.0, 0.0), fa = make_double2(0.0, 0.0);
    if (live) {
#pragma unroll 2
        for (int r = ty; r < no + na * na; r += kGroups) {       // fixed assignment of rows to groups: deterministic
            const bool core = r < no;
            const int e = r - no;
            const int a = core ? r : no + e / na, c = core ? r : no + e % na;
            const int64_t row = ((int64_t)a * nIp + c) * mat;
            const double2 j = __ldg(reinterpret_cast<const double2 *>(J + row));
            const double2 k = __ldg(reinterpret_cast<const double2 *>(K + row));
            if (core) {
                fi.x += 2.0 * j.x - k.x;
                fi.y += 2.0 * j.y - k.y;
            } else {
                const double gm = s_d1[e];
                fa.x += gm * (j.x - 0.5 * k.x);
                fa.y += gm * (j.y - 0.5 * k.y);
            }
        }
    }
    red[ty * kFockCols + tx] = fi;
    red[(kGroups + ty) * kFockCols + tx] = fa;
    __syncthreads();
    if (ty != 0 || !live) return;
    fi = *reinterpret_cast<const double2 *>(h + (int64_t)b * h_stride + e0);
    fa = make_double2(0.0, 0.0);
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
        const double2 pi = red[g * kFockCols + tx], pa = red[(kGroups + g) * kFockCols + tx];
        fi.x += pi.x; fi.y += pi.y;
        fa.x += pa.x; fa.y += pa.y;
    }
    // rows / columns beyond N are zero padding
    const int m = (int)(e0 / ld), n = (int)(e0 % ld);
    if (m >= N) fi = fa = make_double2(0.0, 0.0);
    if (n >= N) fi.x = fa.x = 0.0;
    if (n + 1 >= N) fi.y = fa.y = 0.0;
    *reinterpret_cast<double2 *>(FI + (int64_t)b * mat + e0) = fi;
    if (FA) *reinterpret_cast<double2 *>(FA + (int64_t)b * mat + e0) = fa;
}

template <class GV>
static void launch_fock_core_active(const double *h, int64_t h_stride, const GV &gv, const double *d1, int64_t sd1,
                                    int no, int na, int N, int ld, int batch, double *FI, double *FA, size_t sm1,
                                    cudaStream_t stream) {
    dim3 grid((unsigned)ceil_div((int64_t)ld * ld, 8), (unsigned)batch);
    fock_core_active_kernel<GV><<<grid, 256, sm1, stream>>>(h, h_stride, gv, d1, sd1, no, na, N, ld, FI, FA);
}

static void launch_fock_core_active(const double *h, int64_t h_stride, const ClassView &gv, const double *d1,
                                    int64_t sd1, int no, int na, int N, int ld, int batch, double *FI, double *FA,
                                    size_t sm1, cudaStream_t stream) {
    dim3 grid((unsigned)ceil_div((int64_t)ld * ld / 2, kFockCols), (unsigned)batch);
    const size_t smem = (size_t)((na * na + 1) & ~1) * sizeof(double) + 2 * 256 * sizeof(double2);
    fock_core_active_class_kernel<<<grid, 256, smem, stream>>>(h, h_stride, gv, d1, sd1, no, na, N, ld, FI, FA);
    (void)sm1;
}

// ---------------------------------------------------------------- generalized Fock
// grid (ld, batch): block n computes column n of F (rows = first index).
//   F[i,n] = 2 (FI[n,i] + FA[n,i])                       i in occ
//   F[v,n] = sum_w FI[n,w] d1[v,w] + sum_wxy d2[v,w,x,y] g[n,w,x,y]    v in act
//   F[a,n] = 0                                           a in virt / padding
template <class GV>
__global__ void __launch_bounds__(256)
fock_general_kernel(const GV gv, const double *__restrict__ FI,
                    const double *__restrict__ FA, const double *__restrict__ d1, int64_t sd1,
                    const double *__restrict__ d2, int64_t sd2, int no, int na, int N, int ld,
                    double *__restrict__ F) {
    extern __shared__ double s_g[];   // g[n, act, act, act]  (na^3)
    const int b = blockIdx.y, n = blockIdx.x;
    const int64_t mat = (int64_t)ld * ld;
    const double *FIb = FI + (int64_t)b * mat, *FAb = FA + (int64_t)b * mat;
    double *Fb = F + (int64_t)b * mat;
    const int na2 = na * na, na3 = na2 * na;
    if (n >= N) {
        for (int r = threadIdx.x; r < ld; r += blockDim.x) Fb[(int64_t)r * ld + n] = 0.0;
        return;
    }
    const GV gb = gv.at(b);
    for (int e = threadIdx.x; e < na3; e += blockDim.x) {
        const int w = e / na2, x = (e / na) % na, y = e % na;
        s_g[e] = gb.coul(n, no + w, no + x, no + y);
    }
    __syncthreads();
    for (int r = threadIdx.x; r < ld; r += blockDim.x) {
        if (r < no) Fb[(int64_t)r * ld + n] = 2.0 * (FIb[(int64_t)n * ld + r] + FAb[(int64_t)n * ld + r]);
        else if (r >= no + na) Fb[(int64_t)r * ld + n] = 0.0;
    }
    const double *d1b = d1 + (int64_t)b * sd1;
    const double *d2b = d2 + (int64_t)b * sd2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int v = warp; v < na; v += nwarp) {
        double s = 0.0;
        for (int w = lane; w < na; w += 32) s += FIb[(int64_t)n * ld + no + w] * d1b[v * na + w];
        const double *d2v = d2b + (int64_t)v * na3;
        double s2 = 0.0;
        for (int e = lane; e < na3; e += 32) s2 += d2v[e] * s_g[e];
        s = warp_sum(s);
        s2 = warp_sum(s2);
        if (lane == 0) Fb[(int64_t)(no + v) * ld + n] = s + s2;
    }
}

// ---------------------------------------------------------------- gradient
__global__ void gradient_matrix_kernel(const double *__restrict__ F, int ld, double *__restrict__ G) {
    const int b = blockIdx.z;
    const int p = blockIdx.y * blockDim.y + threadIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= ld || q >= ld) return;
    const int64_t mat = (int64_t)ld * ld;
    const double *Fb = F + (int64_t)b * mat;
    G[(int64_t)b * mat + (int64_t)p * ld + q] = 2.0 * (Fb[(int64_t)p * ld + q] - Fb[(int64_t)q * ld + p]);
}

__global__ void gradient_pack_kernel(const double *__restrict__ F, const int32_t *__restrict__ pl,
                                     const int32_t *__restrict__ pr, int nk, int ld,
                                     double *__restrict__ gvec) {
    const int b = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nk) return;
    const double *Fb = F + (int64_t)b * ld * ld;
    const int l = pl[j], r = pr[j];
    gvec[(int64_t)b * nk + j] = 2.0 * (Fb[(int64_t)l * ld + r] - Fb[(int64_t)r * ld + l]);
}

// ---------------------------------------------------------------- VJP w.r.t. the RDMs
// Fbar = 2 (Gbar - Gbar^T).
//  blocks [0, na^2):  gbar1[v,w] = sum_{i,n} Fbar[i,n] 2 (g[n,i,v,w] - g[n,w,v,i]/2) + sum_n Fbar[v,n] FI[n,w]
//  blocks [na^2, na^2 + na^4): gbar2[v,w,x,y] = sum_n Fbar[v,n] g[n,w,x,y]
template <class GV>
__global__ void __launch_bounds__(128)
fock_gradient_vjp_kernel(const GV g, const double *__restrict__ FI,
                         const double *__restrict__ Gbar, int no, int na, int N, int ld,
                         double *__restrict__ gbar1, double *__restrict__ gbar2) {
    __shared__ double scratch[32];
    const int na2 = na * na;
    auto fbar = [&](int p, int q) {
        return 2.0 * (Gbar[(int64_t)p * ld + q] - Gbar[(int64_t)q * ld + p]);
    };
    const int64_t blk = blockIdx.x;
    if (blk < na2) {
        const int v = (int)(blk / na), w = (int)(blk % na);
        const int V = no + v, W = no + w;
        double s = 0.0;
        for (int e = threadIdx.x; e < no * N; e += blockDim.x) {
            const int i = e / N, n = e % N;
            s += fbar(i, n) * 2.0 * (g.coul(n, i, V, W) - 0.5 * g.exch(n, W, V, i));
        }
        for (int n = threadIdx.x; n < N; n += blockDim.x) s += fbar(V, n) * FI[(int64_t)n * ld + W];
        s = block_sum(s, scratch);
        if (threadIdx.x == 0) gbar1[v * na + w] = s;
    } else {
        const int64_t e = blk - na2;
        const int y = (int)(e % na), x = (int)((e / na) % na), w = (int)((e / na2) % na),
                  v = (int)(e / ((int64_t)na2 * na));
        double s = 0.0;
        for (int n = threadIdx.x; n < N; n += blockDim.x)
            s += fbar(no + v, n) * g.coul(n, no + w, no + x, no + y);
        s = block_sum(s, scratch);
        if (threadIdx.x == 0) gbar2[e] = s;
    }
}

// ---------------------------------------------------------------- padded <-> dense copies
__global__ void pad_copy_kernel(const double *__restrict__ src, double *__restrict__ dst, int N, int ld,
                                int rank, int to_padded, int64_t total) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t per_pad = rank == 2 ? (int64_t)ld * ld : (int64_t)ld * ld * ld * ld;
    const int64_t per_dense = rank == 2 ? (int64_t)N * N : (int64_t)N * N * N * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        if (to_padded) {
            const int64_t b = i / per_pad;
            int64_t e = i % per_pad;
            int idx[4] = {0, 0, 0, 0};
            bool inside = true;
            for (int d = rank - 1; d >= 0; --d) {
                idx[d] = (int)(e % ld);
                e /= ld;
                inside = inside && idx[d] < N;
            }
            double v = 0.0;
            if (inside) {
                int64_t s = 0;
                for (int d = 0; d < rank; ++d) s = s * N + idx[d];
                v = src[b * per_dense + s];
            }
            dst[i] = v;
        } else {
            const int64_t b = i / per_dense;
            int64_t e = i % per_dense;
            int idx[4] = {0, 0, 0, 0};
            for (int d = rank - 1; d >= 0; --d) {
                idx[d] = (int)(e % N);
                e /= N;
            }
            int64_t s = 0;
            for (int d = 0; d < rank; ++d) s = s * ld + idx[d];
            dst[i] = src[b * per_pad + s];
        }
    }
}

}  // namespace

template <class GV>
static int active_hamiltonian_t(const double *h, int64_t h_stride, const GV &gv, int no, int na, int N,
                                int ld, int batch, double e_nuc, const double *e_nuc_b, double *c0, double *c1,
                                double *c2,
                                cudaStream_t stream) {
    OO_REQUIRE(h && c0 && c1 && c2);
    OO_REQUIRE(no >= 0 && na > 0 && no + na <= N && ld >= N && batch > 0);
    if (batch > 65535) return OO_ERR_UNSUPPORTED;
    const int64_t na4 = (int64_t)na * na * na * na;
    dim3 grid((unsigned)(1 + na * na + ceil_div(na4, 256)), (unsigned)batch);
    active_hamiltonian_kernel<GV><<<grid, 256, 0, stream>>>(h, h_stride, gv, no, na, ld, e_nuc, e_nuc_b, c0, c1, c2);
    OO_LAUNCH_CHECK();
    return OO_OK;
}

int energy(const double *c0, const double *c1, const double *c2, const double *d1, int64_t sd1,
           const double *d2, int64_t sd2, int na, int batch, double *E, cudaStream_t stream) {
    OO_REQUIRE(c0 && c1 && c2 && d1 && d2 && E && na > 0 && batch > 0);
    energy_kernel<<<batch, 1024, 0, stream>>>(c0, c1, c2, d1, sd1, d2, sd2, na, E);
    OO_LAUNCH_CHECK();
    return OO_OK;
}

template <class GV>
static int fock_gradient_t(const double *h, int64_t h_stride, const GV &gv, const double *d1, int64_t sd1,
                           const double *d2, int64_t sd2, int no, int na, int N, int ld, int batch,
                           const int32_t *pl, const int32_t *pr, int nk, double *FI, double *FA, double *F,
                           double *Gmat, double *gvec, cudaStream_t stream) {
    OO_REQUIRE(h && d1 && d2 && FI && FA && F);
    OO_REQUIRE(no >= 0 && na > 0 && no + na <= N && ld >= N && batch > 0);
    OO_REQUIRE(!gvec || nk == 0 || (pl && pr));
    if (batch > 65535) return OO_ERR_UNSUPPORTED;
    const size_t sm1 = (size_t)na * na * sizeof(double);
    const size_t sm3 = (size_t)na * na * na * sizeof(double);
    if (sm1 > 48 * 1024 || sm3 > 48 * 1024) return OO_ERR_UNSUPPORTED;   // na <= 18
    launch_fock_core_active(h, h_stride, gv, d1, sd1, no, na, N, ld, batch, FI, FA, sm1, stream);
    OO_LAUNCH_CHECK();
    {
        dim3 grid((unsigned)ld, (unsigned)batch);
        fock_general_kernel<GV><<<grid, 256, sm3, stream>>>(gv, FI, FA, d1, sd1, d2, sd2, no, na, N, ld, F);
        OO_LAUNCH_CHECK();
    }
    if (Gmat) {
        dim3 block(32, 8);
        dim3 grid((unsigned)ceil_div(ld, 32), (unsigned)ceil_div(ld, 8), (unsigned)batch);
        gradient_matrix_kernel<<<grid, block, 0, stream>>>(F, ld, Gmat);
        OO_LAUNCH_CHECK();
    }
    if (gvec && nk > 0) {
        dim3 grid((unsigned)ceil_div(nk, 256), (unsigned)batch);
        gradient_pack_kernel<<<grid, 256, 0, stream>>>(F, pl, pr, nk, ld, gvec);
        OO_LAUNCH_CHECK();
    }
    return OO_OK;
}

template <class GV>
static int fock_gradient_vjp_t(const GV &gv, const double *FI, const double *Gbar, int no, int na, int N,
                               int ld, double *gbar1, double *gbar2, cudaStream_t stream) {
    OO_REQUIRE(FI && Gbar && gbar1 && gbar2);
    OO_REQUIRE(no >= 0 && na > 0 && no + na <= N && ld >= N);
    const int64_t blocks = (int64_t)na * na + (int64_t)na * na * na * na;
    if (blocks > 0x7fffffffll) return OO_ERR_UNSUPPORTED;
    fock_gradient_vjp_kernel<GV><<<(unsigned)blocks, 128, 0, stream>>>(gv, FI, Gbar, no, na, N, ld, gbar1, gbar2);
    OO_LAUNCH_CHECK();
    return OO_OK;
}

static inline FullView full_view(const double *g, int ld) {
    return FullView{g, ld, (int64_t)ld * ld * ld * ld};
}

// class tensors live in one buffer per evaluation: [K rows nIp^2][J rows nIp^2][h' row], rows of ld^2
static inline int64_t class_rows(int nIp) { return 2 * (int64_t)nIp * nIp + 1; }
static inline ClassView class_view(const double *cls, int ld, int nIp) {
    const int64_t mat = (int64_t)ld * ld;
    return ClassView{cls, cls + (int64_t)nIp * nIp * mat, ld, nIp, class_rows(nIp) * mat};
}
static inline const double *class_h(const double *cls, int ld, int nIp) {
    return cls + 2 * (int64_t)nIp * nIp * ld * ld;
}

int active_hamiltonian(const double *h, const double *g, int no, int na, int N, int ld, int batch,
                       double e_nuc, const double *e_nuc_b, double *c0, double *c1, double *c2,
                       cudaStream_t stream) {
    OO_REQUIRE(g);
    return active_hamiltonian_t(h, (int64_t)ld * ld, full_view(g, ld), no, na, N, ld, batch, e_nuc, e_nuc_b, c0,
                                c1, c2, stream);
}

int fock_gradient(const double *h, const double *g, const double *d1, int64_t sd1, const double *d2,
                  int64_t sd2, int no, int na, int N, int ld, int batch, const int32_t *pl,
                  const int32_t *pr, int nk, double *FI, double *FA, double *F, double *Gmat,
                  double *gvec, cudaStream_t stream) {
    OO_REQUIRE(g);
    return fock_gradient_t(h, (int64_t)ld * ld, full_view(g, ld), d1, sd1, d2, sd2, no, na, N, ld, batch, pl, pr,
                           nk, FI, FA, F, Gmat, gvec, stream);
}

int fock_gradient_vjp(const double *g, const double *FI, const double *Gbar, int no, int na, int N,
                      int ld, double *gbar1, double *gbar2, cudaStream_t stream) {
    OO_REQUIRE(g);
    return fock_gradient_vjp_t(full_view(g, ld), FI, Gbar, no, na, N, ld, gbar1, gbar2, stream);
}

// ---- the same contractions on the class tensors of the partial transform (classes.cu)
int class_active_hamiltonian(const double *cls, int no, int na, int N, int ld, int nIp, int batch,
                             double e_nuc, const double *e_nuc_b, double *c0, double *c1, double *c2,
                             cudaStream_t stream) {
    OO_REQUIRE(cls && nIp >= no + na && (nIp % 2) == 0 && nIp <= ld);
    const ClassView v = class_view(cls, ld, nIp);
    return active_hamiltonian_t(class_h(cls, ld, nIp), v.batch_stride, v, no, na, N, ld, batch, e_nuc, e_nuc_b,
                                c0, c1, c2, stream);
}

int class_fock_gradient(const double *cls, const double *d1, int64_t sd1, const double *d2, int64_t sd2,
                        int no, int na, int N, int ld, int nIp, int batch, const int32_t *pl,
                        const int32_t *pr, int nk, double *FI, double *FA, double *F, double *Gmat,
                        double *gvec, cudaStream_t stream) {
    OO_REQUIRE(cls && nIp >= no + na && (nIp % 2) == 0 && nIp <= ld);
    const ClassView v = class_view(cls, ld, nIp);
    return fock_gradient_t(class_h(cls, ld, nIp), v.batch_stride, v, d1, sd1, d2, sd2, no, na, N, ld, batch, pl,
                           pr, nk, FI, FA, F, Gmat, gvec, stream);
}

int class_fock_gradient_vjp(const double *cls, const double *FI, const double *Gbar, int no, int na, int N,
                            int ld, int nIp, double *gbar1, double *gbar2, cudaStream_t stream) {
    OO_REQUIRE(cls && nIp >= no + na && (nIp % 2) == 0 && nIp <= ld);
    return fock_gradient_vjp_t(class_view(cls, ld, nIp), FI, Gbar, no, na, N, ld, gbar1, gbar2, stream);
}

int pad_copy(const double *src, double *dst, int N, int ld, int rank, int batch, int to_padded,
             cudaStream_t stream) {
    OO_REQUIRE(src && dst && N > 0 && ld >= N && batch > 0 && (rank == 2 || rank == 4));
    const int64_t per = to_padded ? (rank == 2 ? (int64_t)ld * ld : (int64_t)ld * ld * ld * ld)
                                  : (rank == 2 ? (int64_t)N * N : (int64_t)N * N * N * N);
    const int64_t total = per * batch;
    int64_t blocks = ceil_div(total, 256);
    if (blocks > 16 * sm_count()) blocks = 16 * sm_count();
    pad_copy_kernel<<<(unsigned)blocks, 256, 0, stream>>>(src, dst, N, ld, rank, to_padded, total);
    OO_LAUNCH_CHECK();
    return OO_OK;
}

}  // namespace oo

extern "C" {

int oo_active_hamiltonian_f64(const double *h_mo, const double *g_mo, int no, int na, int N, int ld,
                              int batch, double e_nuc, const double *e_nuc_batch, double *c0, double *c1,
                              double *c2, void *stream) {
    return oo::active_hamiltonian(h_mo, g_mo, no, na, N, ld, batch, e_nuc, e_nuc_batch, c0, c1, c2,
                                  (cudaStream_t)stream);
}

int oo_energy_f64(const double *c0, const double *c1, const double *c2, const double *gamma,
                  int64_t stride_rdm1, const double *Gamma, int64_t stride_rdm2, int na, int batch,
                  double *E, void *stream) {
    return oo::energy(c0, c1, c2, gamma, stride_rdm1, Gamma, stride_rdm2, na, batch, E,
                      (cudaStream_t)stream);
}

int oo_fock_gradient_f64(const double *h_mo, const double *g_mo, const double *gamma,
                         int64_t stride_rdm1, const double *Gamma, int64_t stride_rdm2, int no, int na,
                         int N, int ld, int batch, const int32_t *pair_l, const int32_t *pair_r, int nk,
                         double *FI, double *FA, double *F, double *Gmat, double *gvec, void *stream) {
    return oo::fock_gradient(h_mo, g_mo, gamma, stride_rdm1, Gamma, stride_rdm2, no, na, N, ld, batch,
                             pair_l, pair_r, nk, FI, FA, F, Gmat, gvec, (cudaStream_t)stream);
}

int oo_fock_gradient_vjp_f64(const double *g_mo, const double *FI, const double *Gbar, int no, int na,
                             int N, int ld, double *gbar1, double *gbar2, void *stream) {
    return oo::fock_gradient_vjp(g_mo, FI, Gbar, no, na, N, ld, gbar1, gbar2, (cudaStream_t)stream);
}

int oo_pad_copy_f64(const double *src, double *dst, int N, int ld, int rank, int batch, int to_padded,
                    void *stream) {
    return oo::pad_copy(src, dst, N, ld, rank, batch, to_padded, (cudaStream_t)stream);
}
}

extern "C" {

int oo_class_active_hamiltonian_f64(const double *cls, int no, int na, int N, int ld, int nIp, int batch,
                                    double e_nuc, const double *e_nuc_batch, double *c0, double *c1,
                                    double *c2, void *stream) {
    return oo::class_active_hamiltonian(cls, no, na, N, ld, nIp, batch, e_nuc, e_nuc_batch, c0, c1, c2,
                                        (cudaStream_t)stream);
}

int oo_class_fock_gradient_f64(const double *cls, const double *gamma, int64_t stride_rdm1,
                               const double *Gamma, int64_t stride_rdm2, int no, int na, int N, int ld,
                               int nIp, int batch, const int32_t *pair_l, const int32_t *pair_r, int nk,
                               double *FI, double *FA, double *F, double *Gmat, double *gvec, void *stream) {
    return oo::class_fock_gradient(cls, gamma, stride_rdm1, Gamma, stride_rdm2, no, na, N, ld, nIp, batch,
                                   pair_l, pair_r, nk, FI, FA, F, Gmat, gvec, (cudaStream_t)stream);
}

int oo_class_fock_gradient_vjp_f64(const double *cls, const double *FI, const double *Gbar, int no, int na,
                                   int N, int ld, int nIp, double *gbar1, double *gbar2, void *stream) {
    return oo::class_fock_gradient_vjp(cls, FI, Gbar, no, na, N, ld, nIp, gbar1, gbar2, (cudaStream_t)stream);
}
}
