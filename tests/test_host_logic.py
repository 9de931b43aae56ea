"""CPU: host-side logic of the product that needs no kernel -- cache keys of the transformed integrals, lazy-consumer
holds, the packed Hessian layout, pending results.  (Anything that computes goes through the CUDA library and is
tested with ``-m gpu``.)"""
import gc

import numpy as np
import torch

from auto_oo_b200.engine import MOIntegrals, _IntegralsKey
from auto_oo_b200.oo_energy import PendingEvaluation, unpack_hessian


def test_integrals_key_hits_and_misses():
    C = torch.randn(5, 5, dtype=torch.float64)
    key = _IntegralsKey("mo", C)
    assert key.matches("mo", C)                                   # same object, same version: no comparison needed
    assert key.matches("mo", C.clone())                           # another object with equal values (host compare)
    assert not key.matches("oao", C)                              # same matrix, different meaning
    assert not key.matches("mo", C + 1e-12)
    assert not key.matches("mo", C[:4, :4].contiguous())
    C.mul_(2.0)                                                   # in-place write: version counter moves, values differ
    assert not key.matches("mo", C)
    kap = torch.randn(7, dtype=torch.float64)
    kkey = _IntegralsKey("oao", C, kap)
    assert kkey.matches("oao", C, kap.clone()) and not kkey.matches("oao", C) and not kkey.matches("oao", C, 2 * kap)
    assert not key.matches("mo", C, kap)
    # the key keeps its source alive, so the identity test can never be fooled by a recycled address
    D = torch.randn(3, 3, dtype=torch.float64)
    dkey = _IntegralsKey("mo", D)
    ptr = D.data_ptr()
    del D
    gc.collect()
    assert dkey.src.data_ptr() == ptr


def test_lazy_consumers_hold_their_integrals():
    ints = MOIntegrals(None, "class", cls=torch.zeros(1))
    assert not ints.held
    tok = ints.hold()
    tok2 = ints.hold()
    assert ints.held
    del tok
    gc.collect()
    assert ints.held
    del tok2
    gc.collect()
    assert not ints.held


def test_unpack_hessian_layout():
    n = 6
    rng = np.random.default_rng(0)
    H = rng.standard_normal((2, n, n))
    H = H + H.transpose(0, 2, 1)
    r, c = np.tril_indices(n)
    packed = H[:, r, c]                                           # np.tril_indices order, as oo_pack_lower_f64 writes it
    assert packed.shape == (2, n * (n + 1) // 2)
    assert np.array_equal(unpack_hessian(packed), H)
    assert torch.equal(unpack_hessian(torch.as_tensor(packed)), torch.as_tensor(H))
    assert np.array_equal(unpack_hessian(packed[0]), H[0])


def test_pending_evaluation_copies_or_aliases():
    E, G, H = torch.ones(2, dtype=torch.float64), torch.ones(2, 3, dtype=torch.float64), None
    fresh = PendingEvaluation((E, G, H), None, copy=True).wait()
    assert fresh[2] is None and torch.equal(fresh[0], E) and fresh[0].data_ptr() != E.data_ptr()
    same = PendingEvaluation((E, G, H), None, copy=False)
    assert same.wait()[1].data_ptr() == G.data_ptr() and same.wait()[1].data_ptr() == G.data_ptr()


def test_quarter_one_pairing_rule_matches_the_library():
    """engine.pairs_quarter_one mirrors the condition in csrc/classes.cu: 8-fold packed integrals shared by the batch,
    triangular quarter 2 (16 < nIp <= 48), 2 nIp columns that fit a tile configuration (<= 48 or 81 .. 96), no A/B flag."""
    from types import SimpleNamespace
    from auto_oo_b200 import _lib
    from auto_oo_b200.engine import HotPathEngine

    def eng(nIp, **kw):
        d = dict(nIp=nIp, eri_packing="8fold", pair_shard=None, n_geom=0, flags=0, eri_is_symmetric=lambda: True)
        d.update(kw)
        return SimpleNamespace(**d)
    rule = HotPathEngine.pairs_quarter_one
    assert rule(eng(24)) and rule(eng(18)) and rule(eng(44)) and rule(eng(48)) and rule(eng(42))
    assert not rule(eng(16)) and not rule(eng(26)) and not rule(eng(40)) and not rule(eng(50))
    assert not rule(eng(24, eri_packing="pair")) and not rule(eng(24, n_geom=4)) and not rule(eng(24, pair_shard=object()))
    assert not rule(eng(24, eri_is_symmetric=lambda: False))
    for fl in (_lib.OO_FLAG_CLASS_Q1_UNPAIRED, _lib.OO_FLAG_CLASS_UNFUSED_PACK, _lib.OO_FLAG_CLASS_Q2_RECTANGULAR):
        assert not rule(eng(44, flags=fl))
    assert rule(eng(44, flags=_lib.OO_FLAG_CLASS_DIRECT_STORES))
