"""TEST INFRASTRUCTURE (never imported by ``auto_oo_b200``): CPU restatement of the reference's
RDM extraction from a state vector, ``Parameterized_circuit.get_rdms_from_state``
(``src/auto_oo/pqc.py:192-218``) with the operators of ``utils/active_space.py:29-83``.

The reference builds every ``E_pq`` / ``e_pqrs`` as a sparse matrix with
``openfermion.get_sparse_operator`` (third-party, not installed here; any 1.x release: the
reference's ``pyproject.toml`` does not pin it) and evaluates ``<psi| O |psi>`` one operator at a
time.  openfermion's published convention, restated here: Jordan-Wigner ladder operators

    a_j = Z^{(x) j}  (x)  |0><1|  (x)  1^{(x) (n-j-1)},        |1> = occupied,

with qubit 0 the LEFTMOST tensor factor, i.e. the most significant bit of the basis-state index
(the same ordering PennyLane uses for ``qml.state()``).  Spin orbitals of spatial orbital ``p`` are
qubits ``2p`` (up) and ``2p+1`` (down), or ``p`` and ``p + ncas`` with ``up_then_down``
(``active_space.py:44-50``).

Pinned by ``tests/test_rdm.py`` against the one analytically known case of the reference's own
``test_rdms`` (``test/test_pqc.py:273-617``, CAS(2,2) UCCD: the state is ``cos(t/2)|1100> -
sin(t/2)|0011>``, read off ``test_state``'s golden vector) and against the independent determinant-space
construction of ``auto_oo_b200.synthetic.CIVectorCircuit``; the other ``test_rdms`` cases need PennyLane
circuit simulation and stay unpinned.
"""
from __future__ import annotations

import itertools

import numpy as np
import scipy.sparse as sp


def _annihilator(j, n):
    """JW a_j on n qubits as a sparse matrix (qubit 0 = most significant bit)."""
    Z = sp.csr_matrix(np.diag([1.0, -1.0]))
    low = sp.csr_matrix(np.array([[0.0, 1.0], [0.0, 0.0]]))      # |0><1|
    op = sp.identity(1, format="csr")
    for k in range(n):
        op = sp.kron(op, Z if k < j else (low if k == j else sp.identity(2, format="csr")), format="csr")
    return op


def ladder_operators(n_qubits):
    a = [_annihilator(j, n_qubits) for j in range(n_qubits)]
    return a, [x.T.tocsr() for x in a]


def e_pq_matrix(p, q, ncas, a, ad, restricted=True, up_then_down=False):
    """``e_pq`` of ``active_space.py:29-54``."""
    if not restricted:
        return ad[p] @ a[q]
    if up_then_down:
        return ad[p] @ a[q] + ad[p + ncas] @ a[q + ncas]
    return ad[2 * p] @ a[2 * q] + ad[2 * p + 1] @ a[2 * q + 1]


def e_pqrs_matrix(p, q, r, s, ncas, a, ad, restricted=True, up_then_down=False):
    """``e_pqrs`` of ``active_space.py:57-83``: restricted ``E_pq E_rs - delta_qr E_ps``,
    unrestricted ``a+_p a+_q a_r a_s``."""
    if not restricted:
        return ad[p] @ ad[q] @ a[r] @ a[s]
    op = e_pq_matrix(p, q, ncas, a, ad, True, up_then_down) @ e_pq_matrix(r, s, ncas, a, ad, True, up_then_down)
    if q == r:
        op = op - e_pq_matrix(p, s, ncas, a, ad, True, up_then_down)
    return op


def rdms_from_state(state, ncas, restricted=True, up_then_down=False):
    """``get_rdms_from_state`` (``pqc.py:192-218``): one operator at a time, real part of
    ``<psi| O |psi>``."""
    state = np.asarray(state)
    nq = 2 * ncas
    assert state.shape == (2 ** nq,)
    a, ad = ladder_operators(nq)
    size = ncas if restricted else 2 * ncas
    one = np.zeros((size, size))
    two = np.zeros((size,) * 4)
    bra = state.conj()
    for p, q in itertools.product(range(size), repeat=2):
        one[p, q] = (bra @ (e_pq_matrix(p, q, ncas, a, ad, restricted, up_then_down) @ state)).real
        for r, s in itertools.product(range(size), repeat=2):
            two[p, q, r, s] = (bra @ (e_pqrs_matrix(p, q, r, s, ncas, a, ad, restricted, up_then_down)
                                      @ state)).real
    return one, two


def transition_rdms(u, v, ncas, up_then_down=False):
    """Re <u| E_pq |v>, Re <u| e_pqrs |v> (restricted): the bilinear form whose diagonal is the RDM pair."""
    nq = 2 * ncas
    a, ad = ladder_operators(nq)
    one = np.zeros((ncas, ncas))
    two = np.zeros((ncas,) * 4)
    bra = np.asarray(u).conj()
    v = np.asarray(v)
    for p, q in itertools.product(range(ncas), repeat=2):
        one[p, q] = (bra @ (e_pq_matrix(p, q, ncas, a, ad, True, up_then_down) @ v)).real
        for r, s in itertools.product(range(ncas), repeat=2):
            two[p, q, r, s] = (bra @ (e_pqrs_matrix(p, q, r, s, ncas, a, ad, True, up_then_down) @ v)).real
    return one, two


def apply_operator(g1, g2, v, ncas, up_then_down=False):
    """(sum_pq g1_pq E_pq + sum_pqrs g2_pqrs e_pqrs) |v>  (restricted)."""
    nq = 2 * ncas
    a, ad = ladder_operators(nq)
    v = np.asarray(v)
    out = np.zeros(v.shape, dtype=np.result_type(v.dtype, np.float64))
    for p, q in itertools.product(range(ncas), repeat=2):
        if g1[p, q] != 0.0:
            out = out + g1[p, q] * (e_pq_matrix(p, q, ncas, a, ad, True, up_then_down) @ v)
        for r, s in itertools.product(range(ncas), repeat=2):
            if g2[p, q, r, s] != 0.0:
                out = out + g2[p, q, r, s] * (e_pqrs_matrix(p, q, r, s, ncas, a, ad, True, up_then_down) @ v)
    return out


def embed_ci_vector(ci, ncas, nelec_ab, up_then_down=False):
    """Determinant-space CI vector (alpha string major, strings = combinations of orbitals in lexical
    order of their bit patterns as ``auto_oo_b200.synthetic._strings`` lists them) -> 2^(2 ncas) qubit state.
    Operator order inside a determinant: all alpha creators (ascending orbital) left of all beta creators;
    the sign to the JW order of the interleaved layout is the parity of the alpha/beta interleaving."""
    na_, nb_ = nelec_ab
    astr = [sum(1 << i for i in c) for c in itertools.combinations(range(ncas), na_)]
    bstr = [sum(1 << i for i in c) for c in itertools.combinations(range(ncas), nb_)]
    nq = 2 * ncas
    out = np.zeros(2 ** nq, dtype=np.asarray(ci).dtype)
    ci = np.asarray(ci).reshape(len(astr), len(bstr))
    for ia, sa in enumerate(astr):
        for ib, sb in enumerate(bstr):
            occ_a = [i for i in range(ncas) if (sa >> i) & 1]
            occ_b = [i for i in range(ncas) if (sb >> i) & 1]
            if up_then_down:
                qubits_a, qubits_b = occ_a, [i + ncas for i in occ_b]
            else:
                qubits_a, qubits_b = [2 * i for i in occ_a], [2 * i + 1 for i in occ_b]
            # |det> = prod_a a+_a prod_b a+_b |0>; reorder into ascending qubit order
            seq = qubits_a + qubits_b
            inv = sum(1 for i in range(len(seq)) for j in range(i + 1, len(seq)) if seq[i] > seq[j])
            idx = sum(1 << (nq - 1 - q) for q in seq)
            out[idx] = ((-1) ** inv) * ci[ia, ib]
    return out
