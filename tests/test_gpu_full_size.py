"""GPU: the BASELINE.json configurations at their FULL sizes (cfg 4: 114 AOs CAS(6,6); cfg 5: 256 AOs
CAS(12,12); cfg 2: 64 geometries of 34 AOs CAS(4,4)).

Parity with the reference: ``tests/golden/n114_cas66.npz`` holds E, the packed gradient and a digest of the
Hessian (diagonal, products with 8 probe vectors, 20 000 sampled elements, Frobenius norm) computed by the
VERBATIM reference on the seeded inputs this module regenerates; ``n256_cas1212.npz`` the same from
``oracle/class_oracle.py`` (the reference cannot run at N = 256; that oracle is pinned on the verbatim
reference up to N = 114) -- ``oracle/make_golden_large.py``.  Both integral routes are compared with them at
1e-10 Ha / 1e-8 per element.  On top of that, properties that do not depend on the size:

* the rotation by the identity returns the AO integrals bit for bit;
* an orthogonal rotation preserves the Frobenius norm and the two pair traces of the ERI tensor,
  and rotating back returns the input;
* the two independent routes (complete four-index transform / partial J-K class transform, both
  symmetric-packed and general) give the same E, gradient and Hessian;
* E is invariant under occupied-occupied and virtual-virtual rotations (the redundant parameters the
  reference removes, oo_energy.py:97-118);
* the analytic gradient and Hessian are the first and second directional derivatives of
  ``energy_from_kappa`` (finite differences of the energies of a batched kappa sweep);
* the Hessian is symmetric; a batch equals its members evaluated alone.

Tolerances: energy 1e-10 Ha, gradient / Hessian elements 1e-8 (BASELINE.json north_star), scaled only
where a finite-difference truncation error enters (stated at the assertion)."""
import os

import numpy as np
import pytest
import torch

from auto_oo_b200 import _lib
from helpers import GOLDEN, TOL_E, TOL_GH

pytestmark = pytest.mark.gpu
F64 = torch.float64

FULL = ["c6h6_ccpvdz_cas66", "synthetic_n256_cas1212"]
GOLDEN_OF = {"c6h6_ccpvdz_cas66": "n114_cas66", "synthetic_n256_cas1212": "n256_cas1212"}
SEED = 11                      # oracle/make_golden_large.py draws its inputs from the same seed, on the CPU


class Problem:
    def __init__(self, workload):
        from auto_oo_b200 import OO_energy
        from auto_oo_b200.synthetic import CONFIG_SHAPES, SyntheticMol, random_rdms, random_kappa
        self.workload = workload
        self.dev = dev = torch.device("cuda", 0)
        self.nao, self.nelec, self.ncas, self.nelecas = CONFIG_SHAPES[workload]
        # random numbers drawn on the CPU (as the golden generator does), B B^T formed on the device
        self.mol = SyntheticMol(self.nao, self.nelec, seed=SEED, device=dev, rng_device="cpu")
        m = self.mol
        self.input_checksum = np.array([float(m._int1e.sum()), float(m._B.sum()), float(m._oao.sum()),
                                        float(m._oao_mo.sum()), float((m._B ** 2).sum())])
        self.g_sum = float(m._int2e.sum())
        self.oo = OO_energy(self.mol, self.ncas, self.nelecas, oao_mo_coeff=self.mol.random_oao_mo_coeff,
                            device=dev)
        self.mol._B = None
        self.eng = self.oo.engine
        one, two = random_rdms(self.ncas, self.nelecas, seed=SEED)
        self.one, self.two = one.to(dev), two.to(dev)
        self.kappa = random_kappa(self.oo.n_kappa, seed=SEED, batch=2).to(dev)
        self.Coao = self.eng.to_padded(self.oo.oao_mo_coeff, 2)


@pytest.fixture(scope="module", params=FULL)
def prob(request):
    p = Problem(request.param)
    yield p
    p.eng.release_workspaces()
    del p
    torch.cuda.empty_cache()


def _compare_with_golden(d, E, G, H, what):
    """E (1,), G (1, nk), H (nk, nk) device tensors against the fixture's reference values."""
    assert abs(E[0].item() - float(d["E"])) < TOL_E, what
    assert np.abs(G[0].cpu().numpy() - d["G"]).max() < TOL_GH, what
    assert np.abs(H.diagonal().cpu().numpy() - d["H_diag"]).max() < TOL_GH, what
    si = torch.as_tensor(d["H_sample_i"], device=H.device)
    sj = torch.as_tensor(d["H_sample_j"], device=H.device)
    assert np.abs(H[si, sj].cpu().numpy() - d["H_samples"]).max() < TOL_GH, what
    V = torch.as_tensor(d["H_probes"], device=H.device)                       # unit 2-norm probes
    assert np.abs((V @ H.T).cpu().numpy() - d["H_times_probes"]).max() < TOL_GH, what
    assert abs(torch.linalg.matrix_norm(H).item() - float(d["H_fro"])) < TOL_GH * max(1.0, float(d["H_fro"])), what


def test_energy_gradient_hessian_match_the_reference_golden(prob):
    """cfg 4 against the verbatim reference, cfg 5 against the class-path oracle, both integral routes
    (oo_energy.py:199-202, :404-424)."""
    d = np.load(os.path.join(GOLDEN, GOLDEN_OF[prob.workload] + ".npz"))
    assert np.allclose(prob.input_checksum, d["checksum"], rtol=1e-12, atol=0), \
        "seeded inputs no longer reproduce the fixture's inputs"
    assert abs(prob.g_sum - float(d["g_sum"])) <= 1e-10 * abs(float(d["g_sum"]))
    assert np.array_equal(prob.kappa[0].cpu().numpy(), d["kappa"])
    eng, nk = prob.eng, prob.oo.n_kappa
    H = torch.empty(1, nk, nk, dtype=F64, device=prob.dev)
    for path in ("class", "full"):
        E, G, _ = eng.evaluate(prob.Coao, prob.one, prob.two, kappa=prob.kappa[:1], H_out=H, path=path)
        _compare_with_golden(d, E, G, H[0], path)
    del H
    # the active-space Hamiltonian of the same rotation through the public API (active_space.py:147-174)
    U = eng.rotation(prob.kappa[:1])
    Cp = eng.from_padded(eng.mo_coeff(prob.Coao, U), 2)[0]
    c0, c1, c2 = prob.oo.get_active_integrals(Cp)
    assert abs(c0.item() - float(d["c0"])) < TOL_E
    assert np.abs(c1.cpu().numpy() - d["c1"]).max() < TOL_E and np.abs(c2.cpu().numpy() - d["c2"]).max() < TOL_E
    # and the host-tensor face of the fused call
    Eh, Gh, Hh = prob.oo.energy_gradient_hessian(prob.kappa[:1].cpu(), prob.one.cpu(), prob.two.cpu())
    _compare_with_golden(d, Eh, Gh, Hh[0], "host face")


def test_identity_is_exact_and_orthogonal_rotation_preserves_invariants(prob):
    eng, N, ld = prob.eng, prob.nao, prob.eng.ld
    # three N^4 buffers are live below (103 GB at N = 256): drop what the class path built in earlier tests
    eng.release_workspaces()
    eng.g_packed = eng.g_pairT = None
    torch.cuda.empty_cache()
    g = eng.g_ao
    eye = torch.eye(ld, dtype=F64, device=prob.dev)[None]
    out = eng.int2e_transform(eye)
    assert torch.equal(out[0], g)                                    # products with 1.0 and sums of zeros
    # an orthogonal U: Frobenius norm, sum_ij g_iijj and sum_ij g_ijji are invariants
    U = eng.rotation(prob.kappa[:1] * 4.0)
    assert (U[0, :N, :N].T @ U[0, :N, :N] - torch.eye(N, dtype=F64, device=prob.dev)).abs().max() < 2e-12
    out = eng.int2e_transform(U, out=out)
    g4, o4 = g.reshape(ld, ld, ld, ld), out[0].reshape(ld, ld, ld, ld)
    scale = g.abs().max().item()
    n0, n1 = torch.linalg.vector_norm(g).item(), torch.linalg.vector_norm(out).item()
    assert abs(n0 - n1) < 1e-12 * n0
    j0, j1 = torch.einsum('iijj->', g4).item(), torch.einsum('iijj->', o4).item()
    k0, k1 = torch.einsum('ijji->', g4).item(), torch.einsum('ijji->', o4).item()
    assert abs(j0 - j1) < 1e-11 * max(abs(j0), scale) and abs(k0 - k1) < 1e-11 * max(abs(k0), scale)
    # the 8-fold symmetry of the input survives the four one-sided quarter transforms
    assert (o4 - o4.permute(1, 0, 2, 3)).abs().max().item() < 1e-11 * scale
    assert (o4[:32] - o4.permute(2, 3, 0, 1)[:32]).abs().max().item() < 1e-11 * scale
    # and rotating back with U^T returns the AO integrals (N = 114 only: a third N^4 buffer at N = 256
    # would not leave room for the class path's copies)
    if N <= 128:
        back = eng.int2e_transform(U.transpose(1, 2).contiguous(), g_ao=out[0])
        assert (back[0] - g).abs().max().item() < 1e-11 * scale
    del out


def test_routes_agree_on_energy_gradient_hessian(prob):
    eng, nk = prob.eng, prob.oo.n_kappa
    Hf = torch.empty(1, nk, nk, dtype=F64, device=prob.dev)
    Hc = torch.empty(1, nk, nk, dtype=F64, device=prob.dev)
    Ef, Gf, _ = eng.evaluate(prob.Coao, prob.one, prob.two, kappa=prob.kappa[:1], H_out=Hf, path="full")
    assert eng.eri_is_symmetric()
    Ec, Gc, _ = eng.evaluate(prob.Coao, prob.one, prob.two, kappa=prob.kappa[:1], H_out=Hc, path="class")
    assert (Ef - Ec).abs().max().item() < TOL_E
    assert (Gf - Gc).abs().max().item() < TOL_GH
    assert (Hf - Hc).abs().max().item() < TOL_GH
    assert (Hc[0] - Hc[0].T).abs().max().item() < TOL_GH
    # Hessian assembly: the row-tiled kernel (taken at this size) against one thread per element
    try:
        eng.flags = _lib.OO_FLAG_HESSIAN_ASSEMBLE_PER_ELEMENT
        _, _, _ = eng.evaluate(prob.Coao, prob.one, prob.two, kappa=prob.kappa[:1], H_out=Hf, path="class")
    finally:
        eng.flags = 0
    assert (Hf - Hc).abs().max().item() < 1e-11 * max(1.0, Hc.abs().max().item())
    # assembly of the rows outside occ+act: bulk-async streamed kernel (default) against the load-per-thread kernel
    try:
        eng.flags = _lib.OO_FLAG_HESSIAN_ASSEMBLE_UNSTREAMED
        _, _, _ = eng.evaluate(prob.Coao, prob.one, prob.two, kappa=prob.kappa[:1], H_out=Hf, path="class")
    finally:
        eng.flags = 0
    assert torch.equal(Hf, Hc)
    # G blocks of the Hessian T-matrix: bulk-async streamed kernel (taken at this size) against per-thread loads
    try:
        eng.flags = _lib.OO_FLAG_HESSIAN_GROUP_UNSTREAMED
        _, _, _ = eng.evaluate(prob.Coao, prob.one, prob.two, kappa=prob.kappa[:1], H_out=Hf, path="class")
    finally:
        eng.flags = 0
    assert (Hf - Hc).abs().max().item() < 1e-11 * max(1.0, Hc.abs().max().item())     # DMMA vs FMA summation order
    # class-pair packing: fused into the quarter-2 GEMM epilogues (default) against the separate pass
    try:
        eng.flags = _lib.OO_FLAG_CLASS_UNFUSED_PACK
        Eu, Gu, _ = eng.evaluate(prob.Coao, prob.one, prob.two, kappa=prob.kappa[:1], H_out=Hf, path="class")
    finally:
        eng.flags = 0
    assert (Eu - Ec).abs().max().item() < 1e-11 and (Gu - Gc).abs().max().item() < 1e-11
    assert (Hf - Hc).abs().max().item() < 1e-11 * max(1.0, Hc.abs().max().item())
    # quarter 2: triangular kernel (only class pairs n <= m are computed; default) against the rectangular GEMM
    try:
        eng.flags = _lib.OO_FLAG_CLASS_Q2_RECTANGULAR
        Er, Gr, _ = eng.evaluate(prob.Coao, prob.one, prob.two, kappa=prob.kappa[:1], H_out=Hf, path="class")
    finally:
        eng.flags = 0
    assert torch.equal(Er, Ec) and torch.equal(Gr, Gc) and torch.equal(Hf, Hc)     # same k order per element
    # occ-occ off-diagonal columns of the T-matrix in pairs sharing their rows (default) against column by column
    try:
        eng.flags = _lib.OO_FLAG_HESSIAN_SPMM_UNPAIRED
        _, _, _ = eng.evaluate(prob.Coao, prob.one, prob.two, kappa=prob.kappa[:1], H_out=Hf, path="class")
    finally:
        eng.flags = 0
    assert (Hf - Hc).abs().max().item() < 1e-11 * max(1.0, Hc.abs().max().item())     # fma order within a column
    # every GEMM of the transform: tiles staged in shared memory and shipped by bulk copies (default) against direct stores
    try:
        eng.flags = _lib.OO_FLAG_CLASS_DIRECT_STORES
        Ed, Gd, _ = eng.evaluate(prob.Coao, prob.one, prob.two, kappa=prob.kappa[:1], H_out=Hf, path="class")
    finally:
        eng.flags = 0
    # (the staged Coulomb quarter 4 computes J[mn][a][b] for a >= b only and mirrors it: round-off apart)
    assert (Ed - Ec).abs().max().item() < 1e-11 and (Gd - Gc).abs().max().item() < 1e-11
    assert (Hf - Hc).abs().max().item() < 1e-11 * max(1.0, Hc.abs().max().item())
    # AO integrals with one pair packed (N^4 / 2, TMA-tiled quarter 1) against the default 8-fold packed tensor
    # (N^4 / 8, quarter-1 rows gathered by bulk copies); both read the same g up to its own symmetry defect
    eng.g_packed, eng.eri_packing = None, "pair"
    try:
        Ep, Gp, _ = eng.evaluate(prob.Coao, prob.one, prob.two, kappa=prob.kappa[:1], H_out=Hf, path="class")
    finally:
        eng.g_packed, eng.eri_packing = None, "8fold"
        torch.cuda.empty_cache()
    assert (Ep - Ec).abs().max().item() < TOL_E and (Gp - Gc).abs().max().item() < TOL_GH
    assert (Hf - Hc).abs().max().item() < TOL_GH
    # general (no symmetry assumed) class route; the complete transform's N^4 workspace makes room first
    eng._eri_symmetric = False
    eng._ws.pop("i2e", None)
    eng._ws.pop("cls", None)
    eng._icache.clear()
    try:
        Hg = Hf
        Eg, Gg, _ = eng.evaluate(prob.Coao, prob.one, prob.two, kappa=prob.kappa[:1], H_out=Hg, path="class")
    finally:
        eng._eri_symmetric = True
        eng.g_pairT = None
        eng._ws.pop("cls", None)
        eng._icache.clear()
        torch.cuda.empty_cache()
    assert (Eg - Ec).abs().max().item() < TOL_E
    assert (Gg - Gc).abs().max().item() < TOL_GH
    assert (Hg - Hc).abs().max().item() < TOL_GH


def test_batch_equals_members_and_redundant_rotations_leave_energy_unchanged(prob):
    eng, N = prob.eng, prob.nao
    no, na = eng.no, eng.na
    E2, G2, _ = eng.evaluate(prob.Coao, prob.one, prob.two, kappa=prob.kappa, want_hessian=False)
    for b in range(2):
        E1, G1, _ = eng.evaluate(prob.Coao, prob.one, prob.two, kappa=prob.kappa[b:b + 1], want_hessian=False)
        assert (E1[0] - E2[b]).abs().item() < TOL_E and (G1[0] - G2[b]).abs().max().item() < TOL_GH
    # C_oao -> C_oao . blockdiag(R_occ, 1, R_virt): same energy (oo_energy.py:97-118 drops these rotations)
    gen = torch.Generator().manual_seed(3)
    R = torch.eye(N, dtype=F64)
    if no > 1:
        R[:no, :no] = torch.linalg.qr(torch.randn(no, no, dtype=F64, generator=gen))[0]
    nv = N - no - na
    if nv > 1:
        R[no + na:, no + na:] = torch.linalg.qr(torch.randn(nv, nv, dtype=F64, generator=gen))[0]
    Crot = eng.to_padded(prob.oo.oao_mo_coeff.cpu() @ R, 2)
    Ea, _, _ = eng.evaluate(prob.Coao, prob.one, prob.two, want_hessian=False)
    Eb, _, _ = eng.evaluate(Crot, prob.one, prob.two, want_hessian=False)
    assert abs(Ea.item() - Eb.item()) < TOL_E * max(1.0, abs(Ea.item()) * 1e-2)


def test_paired_quarter_one_equals_one_gemm_per_evaluation(prob):
    """Class ranges of at most 24 orbitals (114 orbitals CAS(6,6)): two evaluations of a batch share ONE quarter-1 GEMM
    over the packed integrals (48 columns [C_2j | C_2j+1]) and the quarter-2 launches read their halves of its rows;
    an odd last evaluation runs alone.  Same k order per element: bit-identical to one GEMM per evaluation."""
    eng, nk = prob.eng, prob.oo.n_kappa
    kap = torch.cat([prob.kappa, 0.5 * prob.kappa[:1]])                    # a pair and an odd one
    H3 = torch.empty(3, nk, nk, dtype=F64, device=prob.dev) if prob.nao <= 128 else None
    E3, G3, _ = eng.evaluate(prob.Coao, prob.one, prob.two, kappa=kap, want_hessian=H3 is not None, H_out=H3)
    H1 = None if H3 is None else torch.empty_like(H3)
    try:
        eng.flags = _lib.OO_FLAG_CLASS_Q1_UNPAIRED
        E1, G1, _ = eng.evaluate(prob.Coao, prob.one, prob.two, kappa=kap, want_hessian=H1 is not None, H_out=H1)
    finally:
        eng.flags = 0
    assert torch.equal(E3, E1) and torch.equal(G3, G1)
    assert H3 is None or torch.equal(H3, H1)
    for b in range(3):                  # and each member on its own (its own squaring count in expm: round-off apart)
        Eb, Gb, _ = eng.evaluate(prob.Coao, prob.one, prob.two, kappa=kap[b:b + 1], want_hessian=False)
        assert (Eb[0] - E3[b]).abs().item() < TOL_E and (Gb[0] - G3[b]).abs().max().item() < TOL_GH


def test_gradient_and_hessian_are_derivatives_of_the_energy(prob):
    eng, nk = prob.eng, prob.oo.n_kappa
    H = torch.empty(1, nk, nk, dtype=F64, device=prob.dev)
    E0, G, _ = eng.evaluate(prob.Coao, prob.one, prob.two, H_out=H)
    gen = torch.Generator(device=prob.dev).manual_seed(5)
    for trial in range(2):
        d = torch.randn(nk, dtype=F64, device=prob.dev, generator=gen)
        d /= torch.linalg.vector_norm(d)
        t = 1e-2
        steps = torch.tensor([-2.0, -1.0, 1.0, 2.0], dtype=F64, device=prob.dev) * t
        E, _, _ = eng.evaluate(prob.Coao, prob.one, prob.two, kappa=steps[:, None] * d[None, :], want_hessian=False)
        em2, em1, ep1, ep2 = (x.item() for x in E)
        e0 = E0.item()
        d1 = (em2 - 8 * em1 + 8 * ep1 - ep2) / (12 * t)                       # 4th-order stencils
        d2 = (-em2 + 16 * em1 - 30 * e0 + 16 * ep1 - ep2) / (12 * t * t)
        g_d = torch.dot(G[0], d).item()
        h_dd = torch.dot(d, H[0] @ d).item()
        curv = max(abs(h_dd), 1.0)
        # truncation ~ t^4 E^(5,6)/30..90 and round-off ~ 30 eps |E| / (12 t^2): both < 1e-5 of the curvature scale
        assert abs(d1 - g_d) < 1e-6 * max(abs(g_d), curv), (d1, g_d)
        assert abs(d2 - h_dd) < 1e-5 * curv, (d2, h_dd)


def test_berry_loop_shape_64_geometries_against_the_oracle():
    """cfg 2 at full size: 64 geometries of 34 AOs, CAS(4,4), one batched launch per stage; three of them are
    re-evaluated by the CPU oracle."""
    from auto_oo_b200 import OO_energy_geometries
    from auto_oo_b200.synthetic import CONFIG_SHAPES, SyntheticMol, random_rdms, random_kappa
    from oracle import oo_oracle as orc
    nao, nelec, ncas, nelecas = CONFIG_SHAPES["ch2nh_631gs_cas44"]
    G = 64
    mols = [SyntheticMol(nao, nelec, seed=100 + g) for g in range(G)]
    C = torch.stack([torch.as_tensor(m.random_oao_mo_coeff) for m in mols])
    oog = OO_energy_geometries(mols, ncas, nelecas, C, device="cuda:0")
    one, two = random_rdms(ncas, nelecas, seed=2)
    kappa = random_kappa(oog.n_kappa, seed=2, batch=G)
    E, Gv, H = oog.energy_gradient_hessian(kappa, one, two)
    assert E.shape == (G,) and Gv.shape == (G, oog.n_kappa) and H.shape == (G, oog.n_kappa, oog.n_kappa)
    for g in (0, 31, 63):
        m = mols[g]
        p = orc.OracleProblem(m.int1e_ao, m.int2e_ao, m.oao_coeff, m.random_oao_mo_coeff, m.nuc, nelec, ncas,
                              nelecas, False)
        Eo, Go, Ho = p.evaluate(one, two, kappa[g])
        assert abs(E[g].item() - float(Eo)) < TOL_E
        assert np.abs(Gv[g].numpy() - np.asarray(Go)).max() < TOL_GH
        assert np.abs(H[g].numpy() - np.asarray(Ho)).max() < TOL_GH
