"""RDM extraction from a state vector: the CPU restatement of the reference's
``get_rdms_from_state`` (oracle/rdm_oracle.py) against the reference's own golden numbers, and the CUDA
path (auto_oo_b200/rdm.py, csrc/rdm.cu) against that oracle -- values, adjoint pair, first and second
derivatives through autograd."""
import itertools

import numpy as np
import pytest
import torch

F64 = torch.float64

# reference test/test_pqc.py:273-292 (test_rdms, first case: CAS(2,2) UCCD, theta = 2.71668364); the state
# is cos(theta/2)|1100> - sin(theta/2)|0011> (read off test_state's golden vector, test/test_pqc.py:33-60)
REF_THETA = 2.71668364
REF_ONE = np.array([[8.89237535e-02, 0.0], [0.0, 1.91107625e+00]])
REF_TWO_NONZERO = {(0, 0, 0, 0): 8.89237535e-02, (0, 1, 0, 1): -4.12237884e-01, (1, 0, 1, 0): -4.12237884e-01,
                   (1, 1, 1, 1): 1.91107625e+00}


def ref_state():
    psi = np.zeros(16)
    psi[0b1100] = np.cos(REF_THETA / 2)
    psi[0b0011] = -np.sin(REF_THETA / 2)
    return psi


def random_state(ncas, seed, cplx=True, nelec=None):
    rng = np.random.default_rng(seed)
    D = 4 ** ncas
    psi = rng.standard_normal(D) + (1j * rng.standard_normal(D) if cplx else 0.0)
    if nelec is not None:                                   # project on a particle-number sector
        pop = np.array([bin(x).count("1") for x in range(D)])
        psi = np.where(pop == nelec, psi, 0.0)
    return psi / np.linalg.norm(psi)


# ------------------------------------------------------------------------------------------ CPU: oracle
def test_oracle_reproduces_reference_golden_rdms():
    from oracle import rdm_oracle as ro
    one, two = ro.rdms_from_state(ref_state(), 2)
    assert np.abs(one - REF_ONE).max() < 1e-8
    ref_two = np.zeros((2,) * 4)
    for k, v in REF_TWO_NONZERO.items():
        ref_two[k] = v
    assert np.abs(two - ref_two).max() < 1e-8


@pytest.mark.parametrize("utd", [False, True])
def test_oracle_equals_determinant_space_rdms(utd):
    """independent construction: string-space excitation tables of auto_oo_b200.synthetic.CIVectorCircuit"""
    from oracle import rdm_oracle as ro
    from auto_oo_b200.synthetic import CIVectorCircuit
    c = CIVectorCircuit(3, 4, n_theta=3, seed=1)
    ci = c.state(torch.tensor([0.3, -0.2, 0.5]))
    o1, o2 = c.get_rdms_from_state(ci)
    psi = ro.embed_ci_vector(ci.numpy(), 3, c.nelecas, up_then_down=utd)
    a1, a2 = ro.rdms_from_state(psi, 3, up_then_down=utd)
    assert np.abs(a1 - o1.numpy()).max() < 1e-13 and np.abs(a2 - o2.numpy()).max() < 1e-13


def test_oracle_sum_rules():
    from oracle import rdm_oracle as ro
    psi = random_state(2, 5, cplx=True, nelec=3)
    one, two = ro.rdms_from_state(psi, 2)
    assert abs(np.trace(one) - 3) < 1e-12
    assert abs(np.einsum('pprr->', two) - 3 * 2) < 1e-12
    assert np.abs(two - two.transpose(2, 3, 0, 1)).max() < 1e-12          # Gamma_pqrs = Gamma_rspq


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_cuda_rdms_reference_golden():
    from auto_oo_b200 import StatevectorRDM
    one, two = StatevectorRDM(2).get_rdms_from_state(torch.as_tensor(ref_state()))
    assert one.device.type == "cpu"
    assert np.abs(one.numpy() - REF_ONE).max() < 1e-8
    for idx in itertools.product(range(2), repeat=4):
        assert abs(two[idx].item() - REF_TWO_NONZERO.get(idx, 0.0)) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("ncas,cplx,utd,chunk", [(1, True, False, 4), (2, True, False, 1 << 18), (2, False, True, 4),
                                                 (3, True, True, 16), (3, False, False, 1 << 18),
                                                 (4, True, False, 64), (5, False, False, 256)])
def test_cuda_rdms_match_oracle(ncas, cplx, utd, chunk):
    from oracle import rdm_oracle as ro
    from auto_oo_b200 import StatevectorRDM
    psi = random_state(ncas, 10 + ncas, cplx=cplx)
    rdm = StatevectorRDM(ncas, up_then_down=utd, chunk=chunk)
    one, two = rdm.get_rdms_from_state(torch.as_tensor(psi).cuda())
    assert one.is_cuda and two.shape == (ncas,) * 4
    r1, r2 = ro.rdms_from_state(psi, ncas, up_then_down=utd)
    assert np.abs(one.cpu().numpy() - r1).max() < 1e-13
    assert np.abs(two.cpu().numpy() - r2).max() < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("cplx_u,cplx_v", [(True, True), (False, True), (True, False), (False, False)])
def test_cuda_transition_and_apply_are_adjoint_and_match_oracle(cplx_u, cplx_v):
    from oracle import rdm_oracle as ro
    from auto_oo_b200 import StatevectorRDM
    ncas = 3
    u, v = random_state(ncas, 1, cplx_u), random_state(ncas, 2, cplx_v)
    rng = np.random.default_rng(3)
    g1, g2 = rng.standard_normal((ncas,) * 2), rng.standard_normal((ncas,) * 4)
    rdm = StatevectorRDM(ncas, chunk=16)
    t1, t2 = rdm.transition_rdms(torch.as_tensor(u), torch.as_tensor(v))
    r1, r2 = ro.transition_rdms(u, v, ncas)
    assert np.abs(t1.numpy() - r1).max() < 1e-13 and np.abs(t2.numpy() - r2).max() < 1e-13
    w = rdm.apply_operator(g1, g2, torch.as_tensor(v))
    wr = ro.apply_operator(g1, g2, v, ncas)
    assert np.abs(w.numpy() - wr).max() < 1e-13
    # <u| A v> = sum g T(u, v)
    lhs = np.real(np.vdot(u, w.numpy()))
    rhs = float((g1 * t1.numpy()).sum() + (g2 * t2.numpy()).sum())
    assert abs(lhs - rhs) < 1e-12


@pytest.mark.gpu
def test_cuda_rdms_of_a_ci_vector_cas44():
    from oracle import rdm_oracle as ro
    from auto_oo_b200 import StatevectorRDM
    from auto_oo_b200.synthetic import CIVectorCircuit
    c = CIVectorCircuit(4, 4, n_theta=3, seed=2)
    ci = c.state(torch.tensor([0.4, 0.1, -0.3]))
    o1, o2 = c.get_rdms_from_state(ci)
    psi = ro.embed_ci_vector(ci.numpy(), 4, c.nelecas)
    one, two = StatevectorRDM(4).get_rdms_from_state(torch.as_tensor(psi))
    assert (one - o1).abs().max().item() < 1e-13 and (two - o2).abs().max().item() < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("ncas,nelec", [(6, 6), (8, 8)])
def test_cuda_rdms_sum_rules_at_size(ncas, nelec):
    """sizes the oracle cannot reach in seconds: trace and symmetry properties of N-electron states,
    several x-chunks and k-slices"""
    from auto_oo_b200 import StatevectorRDM
    psi = torch.as_tensor(random_state(ncas, 7, cplx=False, nelec=nelec)).cuda()
    one, two = StatevectorRDM(ncas, chunk=1 << 12).get_rdms_from_state(psi)
    assert abs(one.trace().item() - nelec) < 1e-11
    assert abs(torch.einsum('pprr->', two).item() - nelec * (nelec - 1)) < 1e-10
    assert (one - one.T).abs().max().item() < 1e-13
    assert (two - two.permute(2, 3, 0, 1)).abs().max().item() < 1e-13
    assert (two - two.permute(1, 0, 3, 2)).abs().max().item() < 1e-13       # real state
    # contraction to the 1-RDM: sum_r Gamma_pqrr = (N - 1) gamma_pq
    assert (torch.einsum('pqrr->pq', two) - (nelec - 1) * one).abs().max().item() < 1e-11
    # chunking does not change the numbers beyond summation order
    one2, two2 = StatevectorRDM(ncas, chunk=1 << 18).get_rdms_from_state(psi)
    assert (two - two2).abs().max().item() < 1e-13


def sector_state(ncas, seed, sectors, cplx, utd=False):
    """random state supported on the given (n_up, n_down) particle-number sectors"""
    rng = np.random.default_rng(seed)
    D, nq = 4 ** ncas, 2 * ncas
    up = [nq - 1 - (p if utd else 2 * p) for p in range(ncas)]
    dn = [nq - 1 - (p + ncas if utd else 2 * p + 1) for p in range(ncas)]
    x = np.arange(D)
    nup = sum(((x >> b) & 1) for b in up)
    ndn = sum(((x >> b) & 1) for b in dn)
    keep = np.zeros(D, dtype=bool)
    for a, b in sectors:
        keep |= (nup == a) & (ndn == b)
    psi = rng.standard_normal(D) + (1j * rng.standard_normal(D) if cplx else 0.0)
    psi = np.where(keep, psi, 0.0)
    return psi / np.linalg.norm(psi), int(keep.sum())


@pytest.mark.gpu
@pytest.mark.parametrize("ncas,sectors,cplx,utd,chunk", [
    (3, [(2, 1)], True, False, 4), (3, [(1, 1)], False, True, 1 << 18), (4, [(2, 2)], True, False, 16),
    (4, [(2, 2), (1, 3)], False, False, 8), (3, [(0, 0)], False, False, 4), (5, [(3, 2)], True, True, 1 << 18)])
def test_cuda_rdms_on_the_occupied_sectors_only(ncas, sectors, cplx, utd, chunk):
    """Number-conserving states: every pass runs on the compact list of basis states of the occupied sectors
    (found on the device) and gives the numbers of the full-space pass; values, the transition form with a
    partner in other sectors, the operator application and first derivatives."""
    from oracle import rdm_oracle as ro
    from auto_oo_b200 import StatevectorRDM
    psi, count = sector_state(ncas, 3, sectors, cplx, utd)
    comp = StatevectorRDM(ncas, up_then_down=utd, chunk=chunk)
    full = StatevectorRDM(ncas, up_then_down=utd, chunk=chunk, compact=False)
    t = torch.as_tensor(psi).cuda().requires_grad_(True)
    one, two = comp.get_rdms_from_state(t)
    assert comp._k.last_rows == count and count < 4 ** ncas // 2
    t2 = torch.as_tensor(psi).cuda().requires_grad_(True)
    one_f, two_f = full.get_rdms_from_state(t2)
    assert full._k.last_rows == 4 ** ncas
    assert (one - one_f).abs().max().item() < 1e-13 and (two - two_f).abs().max().item() < 1e-13
    if ncas <= 3:
        r1, r2 = ro.rdms_from_state(psi, ncas, up_then_down=utd)
        assert np.abs(one.detach().cpu().numpy() - r1).max() < 1e-13
        assert np.abs(two.detach().cpu().numpy() - r2).max() < 1e-13
    # first derivatives through the compact adjoint
    rng = np.random.default_rng(8)
    w1 = torch.as_tensor(rng.standard_normal((ncas,) * 2)).cuda()
    w2 = torch.as_tensor(rng.standard_normal((ncas,) * 4)).cuda()
    ((one * w1).sum() + (two * w2).sum()).backward()
    ((one_f * w1).sum() + (two_f * w2).sum()).backward()
    assert (t.grad - t2.grad).abs().max().item() < 1e-12
    # transition form: the partner also lives in a sector the state does not occupy -> only the overlap counts
    other, _ = sector_state(ncas, 4, sectors + [(min(ncas, sectors[0][0] + 1), sectors[0][1])], cplx, utd)
    a1, a2 = comp.transition_rdms(torch.as_tensor(other).cuda(), torch.as_tensor(psi).cuda())
    b1, b2 = full.transition_rdms(torch.as_tensor(other).cuda(), torch.as_tensor(psi).cuda())
    assert comp._k.last_rows == count
    assert (a1 - b1).abs().max().item() < 1e-13 and (a2 - b2).abs().max().item() < 1e-13
    wa = comp.apply_operator(w1, w2, torch.as_tensor(psi).cuda())
    wb = full.apply_operator(w1, w2, torch.as_tensor(psi).cuda())
    assert (wa - wb).abs().max().item() < 1e-12


def _dense_operators(ncas):
    from oracle import rdm_oracle as ro
    a, ad = ro.ladder_operators(2 * ncas)
    E = np.stack([np.stack([ro.e_pq_matrix(p, q, ncas, a, ad).toarray() for q in range(ncas)]) for p in range(ncas)])
    return torch.as_tensor(E)


def _dense_rdms(E, psi):
    """torch-CPU dense evaluation (differentiable to any order) of the same RDMs"""
    n = E.shape[0]
    Ec = E.to(psi.dtype)
    w = torch.einsum('pqij,j->pqi', Ec, psi)
    one = torch.einsum('i,pqi->pq', psi.conj(), w).real
    two = torch.einsum('qpi,rsi->pqrs', w.conj(), w).real
    return one, two - torch.einsum('qr,ps->pqrs', torch.eye(n, dtype=F64), one)


@pytest.mark.gpu
@pytest.mark.parametrize("cplx", [False, True])
def test_cuda_rdms_first_and_second_derivatives(cplx):
    """d/dtheta and d2/dtheta2 of a scalar function of the RDMs of psi(theta), CUDA autograd Functions
    against torch-CPU dense algebra (what the circuit-circuit Hessian of OO_pqc needs, oo_pqc.py:103-107)."""
    from auto_oo_b200 import StatevectorRDM
    ncas = 2
    D = 4 ** ncas
    E = _dense_operators(ncas)
    gen = torch.Generator().manual_seed(5)
    dt = torch.complex128 if cplx else F64
    basis = torch.randn(3, D, dtype=dt, generator=gen)
    psi0 = torch.randn(D, dtype=dt, generator=gen)
    c1 = torch.randn(ncas, ncas, dtype=F64, generator=gen)
    c2 = torch.randn(ncas, ncas, ncas, ncas, dtype=F64, generator=gen)
    rdm = StatevectorRDM(ncas)

    def state(theta):
        v = psi0 + torch.einsum('k,kx->x', theta.to(dt), basis) + 0.3 * torch.einsum('k,kx->x', (theta ** 2).to(dt),
                                                                                    basis.flip(0))
        return v / torch.linalg.vector_norm(v)

    def f_cuda(theta):
        one, two = rdm.get_rdms_from_state(state(theta))
        return (c1 * one).sum() + (c2 * two).sum() + 0.1 * (one * one).sum()

    def f_dense(theta):
        one, two = _dense_rdms(E, state(theta))
        return (c1 * one).sum() + (c2 * two).sum() + 0.1 * (one * one).sum()

    th = torch.tensor([0.3, -0.5, 0.2], dtype=F64)
    assert abs(f_cuda(th).item() - f_dense(th).item()) < 1e-12
    g_c = torch.autograd.functional.jacobian(f_cuda, th)
    g_d = torch.autograd.functional.jacobian(f_dense, th)
    assert (g_c - g_d).abs().max().item() < 1e-11
    h_c = torch.autograd.functional.hessian(f_cuda, th)
    h_d = torch.autograd.functional.hessian(f_dense, th)
    assert (h_c - h_d).abs().max().item() < 1e-10
    assert (h_c - h_c.T).abs().max().item() < 1e-10


@pytest.mark.gpu
def test_oo_pqc_with_statevector_circuit():
    """OO_pqc driven by a state-vector circuit whose RDMs come from the CUDA extraction: energy, composite
    gradient and the three Hessian blocks against the same circuit with torch-CPU dense RDMs."""
    import auto_oo_b200
    from auto_oo_b200 import StatevectorCircuit
    from helpers import load_case
    c = load_case("n7_cas44")
    ncas, nelecas = c.ncas, c.nelecas
    D = 4 ** ncas
    E = _dense_operators(ncas)
    pop = torch.tensor([bin(x).count("1") for x in range(D)])
    gen = torch.Generator().manual_seed(11)
    basis = torch.randn(3, D, dtype=F64, generator=gen) * (pop == nelecas)
    hf = torch.zeros(D, dtype=F64)
    hf[int("1" * nelecas + "0" * (2 * ncas - nelecas), 2)] = 1.0

    def state(theta):
        v = hf + torch.einsum('k,kx->x', torch.sin(theta), basis)
        return v / torch.linalg.vector_norm(v)

    class DenseCircuit:
        theta_shape = 3

        def get_rdms(self, theta, restricted=True):
            return _dense_rdms(E, state(theta))

    theta = torch.tensor([0.2, -0.1, 0.3], dtype=F64)
    results = []
    for pqc in (StatevectorCircuit(ncas, nelecas, state, 3), DenseCircuit()):
        oo = auto_oo_b200.OO_pqc(pqc, c.mol(), ncas, nelecas, oao_mo_coeff=c.oao_mo_coeff, freeze_active=True)
        kappa = torch.zeros(oo.n_kappa, dtype=F64)
        e = oo.energy_from_parameters(theta, kappa)
        gc = oo.circuit_gradient(theta)
        go = oo.orbital_gradient(theta)
        hcc = oo.circuit_circuit_hessian(theta)
        hco = oo.orbital_circuit_hessian(theta)
        results.append((e, gc, go, hcc, hco))
    for a, b in zip(*results):
        assert (torch.as_tensor(a) - torch.as_tensor(b)).abs().max().item() < 1e-9
