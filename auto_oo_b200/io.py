"""Interchange format for molecular data produced OUTSIDE this package (e.g. by the reference's
``Moldata_pyscf`` on a machine that has PySCF): one ``.npz`` holding what ``OO_energy.__init__`` reads
(reference ``oo_energy.py:143-165``) plus optional orbitals and optimisation state.

    save_problem("ch2nh.npz", mol, oao_mo_coeff=C, theta=theta)      # where PySCF exists
    mol, extras = load_problem("ch2nh.npz")                          # anywhere
    oo = OO_energy(mol, ncas, nelecas, oao_mo_coeff=extras["oao_mo_coeff"])

Real-orbital two-electron integrals are 8-fold symmetric; ``save_problem`` then stores only the
``P(P+1)/2`` unique elements (``P = N(N+1)/2``, PySCF's ``s8`` order: lower triangle of the pair-by-pair
matrix of lower-triangular pairs), an eighth of the dense tensor, and ``load_problem`` expands them -- or, with
``eri="packed"``, hands them over in the 8-fold packed DEVICE layout ``g8[RS][PQ]`` of the symmetric class transform
(``oo_pack_eri_8fold_f64``): the N^4 tensor is then formed neither on the host nor on the device, quarter 1 of the
class transform unpacks the pairs in its producer (``OO_energy`` sees ``mol.int2e_packed8``).
``save_trajectory`` / ``load_trajectory`` checkpoint the orbitals (and circuit parameters, energies) of an
optimisation or of the geometries of a Berry-phase loop.
"""
from __future__ import annotations

import numpy as np

FORMAT_VERSION = 1
_REQUIRED = ("int1e_ao", "int2e_ao", "overlap", "oao_coeff", "nuc", "nelectron")


class ArrayMol:
    """Duck-typed ``Moldata_pyscf`` built from arrays (``moldata_pyscf.py:19-56``): attributes
    ``int1e_ao, int2e_ao, overlap, oao_coeff, nuc, nao, nelectron`` and ``get_active_space_idx``."""

    def __init__(self, int1e_ao, int2e_ao, overlap, oao_coeff, nuc, nelectron, int2e_packed8=None):
        self.int1e_ao = np.asarray(int1e_ao, dtype=np.float64)
        self.int2e_ao = None if int2e_ao is None else np.asarray(int2e_ao, dtype=np.float64)
        self.int2e_packed8 = int2e_packed8               # (ld(ld+1)/2, pair_ld) 8-fold packed, or None
        self.overlap = np.asarray(overlap, dtype=np.float64)
        self.oao_coeff = np.asarray(oao_coeff, dtype=np.float64)
        self.nuc = float(nuc)
        self.nelectron = int(nelectron)
        self.nao = self.int1e_ao.shape[0]
        if self.overlap.shape != (self.nao,) * 2 or (self.int2e_ao is None) == (int2e_packed8 is None) or \
                (self.int2e_ao is not None and self.int2e_ao.shape != (self.nao,) * 4):
            raise ValueError("inconsistent array shapes for an AO basis of size %d" % self.nao)

    def get_active_space_idx(self, ncas, nelecas):
        """Same rule (and error) as ``moldata_pyscf.py:42-56``."""
        nelecore = self.nelectron - nelecas
        if nelecore % 2 == 1:
            raise ValueError('odd number of core electrons')
        occ_idx = np.arange(nelecore // 2)
        act_idx = (occ_idx[-1] + 1 + np.arange(ncas)) if len(occ_idx) > 0 else np.arange(ncas)
        virt_idx = np.arange(act_idx[-1] + 1, self.nao)
        return occ_idx, act_idx, virt_idx


def pack_eri_s8(g):
    """Unique elements of an 8-fold symmetric ``(N,N,N,N)`` tensor: ``out[tri(PQ, RS)]`` with ``PQ = tri(p, q)``,
    ``p >= q``, ``PQ >= RS`` (``tri(a, b) = a(a+1)/2 + b``).  Raises if ``g`` is not 8-fold symmetric."""
    g = np.asarray(g, dtype=np.float64)
    n = g.shape[0]
    scale = max(np.abs(g).max(), 1e-300)
    if (np.abs(g - g.transpose(1, 0, 2, 3)).max() > 1e-12 * scale
            or np.abs(g - g.transpose(2, 3, 0, 1)).max() > 1e-12 * scale):
        raise ValueError("two-electron integrals are not 8-fold symmetric")
    p, q = np.tril_indices(n)
    pairs = g[p, q][:, p, q]                                 # (P, P) pair-by-pair matrix
    a, b = np.tril_indices(len(p))
    return np.ascontiguousarray(pairs[a, b])


def unpack_eri_s8(packed, nao):
    """Inverse of :func:`pack_eri_s8`: the dense ``(N,N,N,N)`` tensor."""
    p, q = np.tril_indices(nao)
    P = len(p)
    a, b = np.tril_indices(P)
    if packed.shape != (len(a),):
        raise ValueError("packed integrals do not match a basis of size %d" % nao)
    pairs = np.empty((P, P))
    pairs[a, b] = packed
    pairs[b, a] = packed
    half = np.empty((P, nao, nao))
    half[:, p, q] = pairs
    half[:, q, p] = pairs
    g = np.empty((nao, nao, nao, nao))
    g[p, q] = half
    g[q, p] = half
    return g


def s8_to_packed8(packed, nao):
    """The ``s8`` vector as the 8-fold packed device layout ``g8[RS][PQ]`` of the symmetric class transform: pairs
    ``tri(p, q)`` over the orbitals padded to an even count (the padding pairs come last: zero rows / columns), row
    length rounded up to even.  ``(ld (ld+1)/2, pair_ld)`` doubles -- a quarter of the dense tensor's size, never N^4."""
    P = nao * (nao + 1) // 2
    a, b = np.tril_indices(P)
    if packed.shape != (len(a),):
        raise ValueError("packed integrals do not match a basis of size %d" % nao)
    ld = nao + (nao & 1)
    rows = ld * (ld + 1) // 2
    out = np.zeros((rows, rows + (rows & 1)))
    out[a, b] = packed
    out[b, a] = packed
    return out


def save_problem(path, mol, nelectron=None, eri_packing="auto", **extras):
    """Write ``mol``'s integrals (and any extra arrays: ``oao_mo_coeff``, ``theta``, RDMs, ...).
    ``eri_packing``: ``"s8"`` stores the unique elements of the 8-fold symmetric ERI tensor (error if it is not
    symmetric), ``"dense"`` the full tensor, ``"auto"`` packs when the symmetry holds."""
    ne = nelectron if nelectron is not None else getattr(mol, "nelectron", None)
    if ne is None:
        raise ValueError("nelectron is required (mol has no .nelectron)")
    if eri_packing not in ("auto", "s8", "dense"):
        raise ValueError("eri_packing must be 'auto', 's8' or 'dense'")
    eri = np.asarray(mol.int2e_ao)
    if eri_packing != "dense":
        try:
            eri = pack_eri_s8(eri)
        except ValueError:
            if eri_packing == "s8":
                raise
    data = dict(format_version=np.asarray(FORMAT_VERSION), int1e_ao=np.asarray(mol.int1e_ao),
                int2e_ao=eri, overlap=np.asarray(mol.overlap),
                oao_coeff=np.asarray(mol.oao_coeff), nuc=np.asarray(float(mol.nuc)), nelectron=np.asarray(int(ne)))
    for k, v in extras.items():
        if k in data:
            raise ValueError(f"extra array name {k!r} collides with a required field")
        data[k] = np.asarray(v.detach().cpu() if hasattr(v, "detach") else v)
    np.savez_compressed(path, **data)


def load_problem(path, eri="dense"):
    """Returns ``(ArrayMol, extras dict)``.  ``eri="packed"`` (s8 files only): the integrals stay 8-fold packed
    (``mol.int2e_packed8``, ``mol.int2e_ao`` is None); ``OO_energy`` then runs the symmetric class path from them."""
    if eri not in ("dense", "packed"):
        raise ValueError("eri must be 'dense' or 'packed'")
    with np.load(path) as d:
        missing = [k for k in _REQUIRED if k not in d.files]
        if missing:
            raise ValueError(f"{path}: missing fields {missing}")
        if int(d["format_version"]) != FORMAT_VERSION:
            raise ValueError(f"{path}: unsupported format version {int(d['format_version'])}")
        stored, want_packed = d["int2e_ao"], eri == "packed"
        nao = d["int1e_ao"].shape[0]
        if want_packed and stored.ndim != 1:
            raise ValueError(f"{path}: eri='packed' needs an s8-packed file (this one stores the dense tensor)")
        dense = None if want_packed else (unpack_eri_s8(stored, nao) if stored.ndim == 1 else stored)
        packed8 = s8_to_packed8(stored, nao) if want_packed else None
        mol = ArrayMol(d["int1e_ao"], dense, d["overlap"], d["oao_coeff"], float(d["nuc"]), int(d["nelectron"]),
                       int2e_packed8=packed8)
        extras = {k: d[k] for k in d.files if k not in _REQUIRED and k != "format_version"}
    return mol, extras


def save_trajectory(path, oao_mo_coeff, theta=None, energies=None, **meta):
    """Checkpoint of an optimisation / a loop over geometries: ``oao_mo_coeff`` is a sequence of ``(N, N)``
    matrices (one per iteration or geometry, e.g. the fourth return value of ``OO_pqc.full_optimization``),
    ``theta`` the matching circuit parameters, ``energies`` the energy list; ``meta``: any further arrays."""
    tonp = lambda v: np.asarray(v.detach().cpu() if hasattr(v, "detach") else v)
    data = dict(format_version=np.asarray(FORMAT_VERSION), oao_mo_coeff=np.stack([tonp(c) for c in oao_mo_coeff]))
    if theta is not None:
        data["theta"] = np.stack([tonp(t).reshape(-1) for t in theta])
        if data["theta"].shape[0] != data["oao_mo_coeff"].shape[0]:
            raise ValueError("theta and oao_mo_coeff trajectories differ in length")
    if energies is not None:
        data["energies"] = np.asarray([float(e) for e in energies])
    for k, v in meta.items():
        if k in data:
            raise ValueError(f"extra array name {k!r} collides with a trajectory field")
        data[k] = tonp(v)
    np.savez_compressed(path, **data)


def load_trajectory(path):
    """Returns a dict with ``oao_mo_coeff (T, N, N)`` and whatever else :func:`save_trajectory` stored."""
    with np.load(path) as d:
        if "oao_mo_coeff" not in d.files or int(d["format_version"]) != FORMAT_VERSION:
            raise ValueError(f"{path}: not a trajectory file of format version {FORMAT_VERSION}")
        return {k: d[k] for k in d.files if k != "format_version"}

