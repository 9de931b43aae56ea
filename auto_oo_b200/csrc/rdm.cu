// Spin-summed reduced density matrices from a state vector.
//
// Reference: Parameterized_circuit.get_rdms_from_state (pqc.py:192-218) evaluates
//   gamma_pq = Re <psi| E_pq |psi>,   Gamma_pqrs = Re <psi| e_pqrs |psi>,
//   E_pq = sum_sigma a+_{p sigma} a_{q sigma},   e_pqrs = E_pq E_rs - delta_qr E_ps
// (utils/active_space.py:29-83) one operator at a time: na^2 + na^4 sparse mat-vecs on a
// 2^(2 na) vector, with openfermion's Jordan-Wigner matrices (qubit 0 = most significant bit,
// spin orbitals 2p / 2p+1, or p / p+na with up_then_down).
//
// Here: with phi_rs = E_rs |v> (na^2 vectors, generated on the fly by bit manipulation: one
// XOR mask and one popcount parity per spin term, no operator matrices),
//   Re <u| E_pq E_rs |v> = Re <E_qp u | E_rs v> = sum_x Re( conj(phi^u_qp[x]) phi^v_rs[x] )
// is ONE TN-DGEMM (dgemm_tn.cu) over x with the real and imaginary parts stacked along k:
//   C[m, n] = sum_k PhiL[k, m] PhiR[k, n],   columns m, n = (r s) plus one identity column (the state
// itself, which yields gamma), rows k = [Re x-chunk ; Im x-chunk].  2 * 2^(2na) * (na^2+1)^2 flop on the
// FP64 tensor pipe instead of na^4 sparse mat-vecs; the gather that builds Phi is L2/HBM-bound.
// The same machinery gives the bilinear (transition) form T(u, v) and the operator application
//   A(g1, g2) v = sum_pq g1_pq E_pq v + sum_pqrs g2_pqrs e_pqrs v
//               = sum_pq E_pq [ Phi_v Mext ][:, pq]
// which are each other's adjoints, so the Python layer differentiates the RDMs to any order in the state.
#include "common.cuh"

namespace oo {

namespace {

struct SpinMap {
    int nq, na, up_then_down;
    // bit position (in the basis-state index) of spin orbital (orb, spin)
    __device__ __forceinline__ int bit(int orb, int spin) const {
        const int q = up_then_down ? orb + spin * na : 2 * orb + spin;
        return nq - 1 - q;
    }
};

// (a+_i a_j v)[x] = sign * v[xsrc] if the term exists; bi, bj = bit positions of i, j
__device__ __forceinline__ bool excite_source(uint32_t x, int bi, int bj, uint32_t &xsrc, double &sign) {
    if (bi == bj) {
        if (!((x >> bi) & 1u)) return false;
        xsrc = x;
        sign = 1.0;
        return true;
    }
    if (!((x >> bi) & 1u) || ((x >> bj) & 1u)) return false;
    xsrc = x ^ (1u << bi) ^ (1u << bj);
    const int lo = bi < bj ? bi : bj, hi = bi < bj ? bj : bi;
    const uint32_t between = ((1u << hi) - 1u) & ~((1u << (lo + 1)) - 1u);
    sign = (__popc(x & between) & 1) ? -1.0 : 1.0;
    return true;
}

// Phi[k][c] (transposed == 0, row length ncolp) or Phi[c][k] (transposed == 1, row length nrows)
//   k = x - x0 (real parts) and nx + x - x0 (imaginary parts, complex input only)
//   c < na^2: (E_rs v)[x], c = r*na + s;  c == na^2: v[x];  c > na^2: zero padding
// xlist != null: rows are the COMPACT basis states x = xlist[x0 + xl] (x0 + xl >= nlist: zero row) -- the
// particle-number sectors the state lives in (rdm_sector_* below); E_rs never leaves a sector.
__global__ void __launch_bounds__(256) rdm_excite_kernel(const double *__restrict__ psi, int is_complex,
                                                         SpinMap sm, int64_t x0, int64_t nx, int ncolp,
                                                         int transposed, double *__restrict__ Phi,
                                                         const int32_t *__restrict__ xlist, int64_t nlist) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nx * ncolp) return;
    int64_t xl;
    int c;
    if (transposed) {
        c = (int)(idx / nx);
        xl = idx % nx;
    } else {
        xl = idx / ncolp;
        c = (int)(idx % ncolp);
    }
    uint32_t x = (uint32_t)(x0 + xl);
    bool live = true;
    if (xlist) {
        live = x0 + xl < nlist;
        x = live ? (uint32_t)xlist[x0 + xl] : 0u;
    }
    const int ncol = sm.na * sm.na;
    double re = 0.0, im = 0.0;
    if (!live) {
    } else if (c < ncol) {
        const int r = c / sm.na, s = c % sm.na;
#pragma unroll
        for (int spin = 0; spin < 2; ++spin) {
            uint32_t xs;
            double sg;
            if (excite_source(x, sm.bit(r, spin), sm.bit(s, spin), xs, sg)) {
                if (is_complex) {
                    const double2 v = reinterpret_cast<const double2 *>(psi)[xs];
                    re += sg * v.x;
                    im += sg * v.y;
                } else {
                    re += sg * psi[xs];
                }
            }
        }
    } else if (c == ncol) {
        if (is_complex) {
            const double2 v = reinterpret_cast<const double2 *>(psi)[x];
            re = v.x;
            im = v.y;
        } else {
            re = psi[x];
        }
    }
    const int64_t nrows = is_complex ? 2 * nx : nx;
    if (transposed) {
        Phi[(int64_t)c * nrows + xl] = re;
        if (is_complex) Phi[(int64_t)c * nrows + nx + xl] = im;
    } else {
        Phi[xl * ncolp + c] = re;
        if (is_complex) Phi[(nx + xl) * ncolp + c] = im;
    }
}

// acc[i] += sum_b parts[b][i]   (fixed order: deterministic)
__global__ void rdm_accumulate_kernel(const double *__restrict__ parts, int nparts, int64_t n,
                                      double *__restrict__ acc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = acc[i];
    for (int b = 0; b < nparts; ++b) s += parts[(int64_t)b * n + i];
    acc[i] = s;
}

// gamma_pq = C[identity, (p q)];  Gamma_pqrs = C[(q p), (r s)] - delta_qr gamma_ps
__global__ void rdm_assemble_kernel(const double *__restrict__ C, int na, int ncolp, double *__restrict__ one,
                                    double *__restrict__ two) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n2 = na * na;
    if (idx >= n2 * n2) return;
    const int s = idx % na, r = (idx / na) % na, q = (idx / n2) % na, p = idx / (n2 * na);
    const double *gam = C + (int64_t)n2 * ncolp;
    double v = C[(int64_t)(q * na + p) * ncolp + (r * na + s)];
    if (q == r) v -= gam[p * na + s];
    two[idx] = v;
    if (idx < n2) one[idx] = gam[idx];
}

// Mext[c][(p q)] (row length ncol2): c = (r s): g2[p,q,r,s];  c = na^2: g1[p,q] - sum_t g2[p,t,t,q];  else 0
__global__ void rdm_operator_matrix_kernel(const double *__restrict__ g1, const double *__restrict__ g2, int na,
                                           int ncolp, int ncol2, double *__restrict__ M) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ncolp * ncol2) return;
    const int c = idx / ncol2, pq = idx % ncol2, n2 = na * na;
    double v = 0.0;
    if (pq < n2) {
        if (c < n2) {
            v = g2[(int64_t)pq * n2 + c];
        } else if (c == n2) {
            const int p = pq / na, q = pq % na;
            v = g1[pq];
            for (int t = 0; t < na; ++t) v -= g2[(((int64_t)p * na + t) * na + t) * na + q];
        }
    }
    M[idx] = v;
}

// w[x] = sum_pq sum_sigma sign * Wt[pq][xsrc]   (E_pq applied to column pq, summed).
// Compact form (xlist/pos != null): k runs over the R compact basis states, x = xlist[k], and column xsrc of
// Wt sits at pos[xsrc]; w (full length, zero-initialised by the caller) is written at x only.
__global__ void __launch_bounds__(256) rdm_apply_gather_kernel(const double *__restrict__ Wre,
                                                               const double *__restrict__ Wim, SpinMap sm,
                                                               int64_t R, int64_t ldW,
                                                               const int32_t *__restrict__ xlist,
                                                               const int32_t *__restrict__ pos,
                                                               double *__restrict__ w) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= R) return;
    const uint32_t x = xlist ? (uint32_t)xlist[k] : (uint32_t)k;
    double re = 0.0, im = 0.0;
    for (int p = 0; p < sm.na; ++p)
        for (int q = 0; q < sm.na; ++q) {
            const int64_t row = (int64_t)(p * sm.na + q) * ldW;
#pragma unroll
            for (int spin = 0; spin < 2; ++spin) {
                uint32_t xs;
                double sg;
                if (excite_source(x, sm.bit(p, spin), sm.bit(q, spin), xs, sg)) {
                    const int64_t ks = pos ? (int64_t)pos[xs] : (int64_t)xs;
                    re += sg * Wre[row + ks];
                    if (Wim) im += sg * Wim[row + ks];
                }
            }
        }
    if (Wim) {
        reinterpret_cast<double2 *>(w)[x] = make_double2(re, im);
    } else {
        w[x] = re;
    }
}

// ---- particle-number sectors ---------------------------------------------------------------------
// sector(x) = n_up(x) * (na + 1) + n_down(x).  Every E_rs conserves both counts, so Phi has non-zero rows only
// on the sectors psi occupies: for the number-conserving ansaetze of the reference (one sector) that is
// C(na, n_up) C(na, n_down) of the 4^na basis states (5 % at CAS(12,12)).
struct SectorMasks {
    uint32_t up, down;
    int na;
    __device__ __forceinline__ int id(uint32_t x) const { return __popc(x & up) * (na + 1) + __popc(x & down); }
};

__global__ void rdm_sector_flags_kernel(const double *__restrict__ psi, int is_complex, SectorMasks m, int64_t D,
                                        int32_t *__restrict__ flags) {
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= D) return;
    const bool nz = is_complex ? (psi[2 * x] != 0.0 || psi[2 * x + 1] != 0.0) : (psi[x] != 0.0);
    if (nz) flags[m.id((uint32_t)x)] = 1;          // benign race: every writer stores 1
}

__global__ void rdm_sector_mask_kernel(const int32_t *__restrict__ flags, SectorMasks m, int64_t D,
                                       int32_t *__restrict__ mask) {
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x < D) mask[x] = flags[m.id((uint32_t)x)] ? 1 : 0;
}

SectorMasks sector_masks(int ncas, int up_then_down) {
    SectorMasks m{0u, 0u, ncas};
    const int nq = 2 * ncas;
    for (int orb = 0; orb < ncas; ++orb)
        for (int spin = 0; spin < 2; ++spin) {
            const int q = up_then_down ? orb + spin * ncas : 2 * orb + spin;
            (spin ? m.down : m.up) |= 1u << (nq - 1 - q);
        }
    return m;
}

inline int rdm_ncolp(int na) {
    const int n = na * na + 1;
    return n + (n & 1);
}

}  // namespace

int rdm_excitations(const double *psi, int is_complex, int ncas, int up_then_down, int64_t x0, int64_t nx,
                    int transposed, double *Phi, const int32_t *xlist, int64_t nlist, cudaStream_t stream) {
    OO_REQUIRE(psi && Phi && ncas > 0 && ncas <= 15 && x0 >= 0 && nx > 0);
    OO_REQUIRE(xlist ? (nlist >= 0 && nlist <= (1ll << (2 * ncas))) : (x0 + nx <= (1ll << (2 * ncas))));
    const SpinMap sm{2 * ncas, ncas, up_then_down ? 1 : 0};
    const int ncolp = rdm_ncolp(ncas);
    const int64_t total = nx * ncolp;
    const int64_t grid = ceil_div(total, 256);
    if (grid > 0x7fffffffll) return OO_ERR_UNSUPPORTED;
    rdm_excite_kernel<<<(unsigned)grid, 256, 0, stream>>>(psi, is_complex ? 1 : 0, sm, x0, nx, ncolp,
                                                          transposed ? 1 : 0, Phi, xlist, nlist);
    OO_LAUNCH_CHECK();
    return OO_OK;
}

}  // namespace oo

extern "C" {

int64_t oo_rdm_columns(int ncas) { return ncas > 0 ? oo::rdm_ncolp(ncas) : 0; }

int oo_rdm_excitations_f64(const double *psi, int is_complex, int ncas, int up_then_down, int64_t x0, int64_t nx,
                           int transposed, double *Phi, const int32_t *xlist, int64_t nlist, void *stream) {
    return oo::rdm_excitations(psi, is_complex, ncas, up_then_down, x0, nx, transposed, Phi, xlist, nlist,
                               (cudaStream_t)stream);
}

int oo_rdm_sector_flags_f64(const double *psi, int is_complex, int ncas, int up_then_down, int32_t *flags,
                            void *stream) {
    OO_REQUIRE(psi && flags && ncas > 0 && ncas <= 15);
    const int64_t D = 1ll << (2 * ncas);
    oo::rdm_sector_flags_kernel<<<(unsigned)oo::ceil_div(D, 256), 256, 0, (cudaStream_t)stream>>>(
        psi, is_complex ? 1 : 0, oo::sector_masks(ncas, up_then_down), D, flags);
    OO_LAUNCH_CHECK();
    return OO_OK;
}

int oo_rdm_sector_mask(const int32_t *flags, int ncas, int up_then_down, int32_t *mask, void *stream) {
    OO_REQUIRE(flags && mask && ncas > 0 && ncas <= 15);
    const int64_t D = 1ll << (2 * ncas);
    oo::rdm_sector_mask_kernel<<<(unsigned)oo::ceil_div(D, 256), 256, 0, (cudaStream_t)stream>>>(
        flags, oo::sector_masks(ncas, up_then_down), D, mask);
    OO_LAUNCH_CHECK();
    return OO_OK;
}

int oo_rdm_accumulate_f64(const double *parts, int nparts, int64_t n, double *acc, void *stream) {
    OO_REQUIRE(parts && acc && nparts > 0 && n > 0);
    oo::rdm_accumulate_kernel<<<(unsigned)oo::ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(parts, nparts, n,
                                                                                                acc);
    OO_LAUNCH_CHECK();
    return OO_OK;
}

int oo_rdm_assemble_f64(const double *C, int ncas, double *one_rdm, double *two_rdm, void *stream) {
    OO_REQUIRE(C && one_rdm && two_rdm && ncas > 0 && ncas <= 15);
    const int n4 = ncas * ncas * ncas * ncas;
    oo::rdm_assemble_kernel<<<(unsigned)oo::ceil_div(n4, 256), 256, 0, (cudaStream_t)stream>>>(
        C, ncas, oo::rdm_ncolp(ncas), one_rdm, two_rdm);
    OO_LAUNCH_CHECK();
    return OO_OK;
}

int oo_rdm_operator_matrix_f64(const double *g1, const double *g2, int ncas, double *Mext, void *stream) {
    OO_REQUIRE(g1 && g2 && Mext && ncas > 0 && ncas <= 15);
    const int ncolp = oo::rdm_ncolp(ncas), n2 = ncas * ncas, ncol2 = n2 + (n2 & 1);
    oo::rdm_operator_matrix_kernel<<<(unsigned)oo::ceil_div((int64_t)ncolp * ncol2, 256), 256, 0,
                                     (cudaStream_t)stream>>>(g1, g2, ncas, ncolp, ncol2, Mext);
    OO_LAUNCH_CHECK();
    return OO_OK;
}

int oo_rdm_apply_gather_f64(const double *Wt_re, const double *Wt_im, int ncas, int up_then_down, int64_t R,
                            int64_t ldW, const int32_t *xlist, const int32_t *pos, double *w, void *stream) {
    OO_REQUIRE(Wt_re && w && ncas > 0 && ncas <= 15 && R >= 0 && ldW >= R);
    OO_REQUIRE((xlist == nullptr) == (pos == nullptr));
    OO_REQUIRE(R <= (1ll << (2 * ncas)) && (xlist || R == (1ll << (2 * ncas))));
    if (R == 0) return OO_OK;
    const oo::SpinMap sm{2 * ncas, ncas, up_then_down ? 1 : 0};
    oo::rdm_apply_gather_kernel<<<(unsigned)oo::ceil_div(R, 256), 256, 0, (cudaStream_t)stream>>>(
        Wt_re, Wt_im, sm, R, ldW, xlist, pos, w);
    OO_LAUNCH_CHECK();
    return OO_OK;
}
}
