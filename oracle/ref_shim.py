"""TEST INFRASTRUCTURE ONLY -- import the VERBATIM reference hot path from
``/root/reference`` (this container) or, where that path does not exist (the GPU box), from the
unmodified install under ``baseline/_ref`` (``pip install --no-deps --target baseline/_ref /root/reference``,
git-ignored, shipped with the snapshot).

The reference needs ``pennylane`` (for ``pennylane.math``), ``pyscf`` and
``openfermion`` at import time; none is installed.  This module registers
* a ``pennylane.math`` namespace backed by torch implementing exactly the
  functions the hot path calls (semantics per PennyLane >= 0.31: ``zeros`` /
  ``eye`` / ``ones`` without ``like`` return numpy, ``set_index`` writes in place
  for torch, ``expm`` -> ``torch.linalg.matrix_exp``, ``einsum`` -> ``torch.einsum``),
* empty ``pyscf`` / ``openfermion`` stubs,
and loads ``moldata_pyscf.py``, ``utils/active_space.py``, ``utils/newton_raphson.py``
and ``oo_energy.py`` by path without executing ``auto_oo/__init__.py``.
No reference source is copied into this repository.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("AUTO_OO_REFERENCE", "/root/reference")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the unmodified reference as installed by ``pip install --no-deps --target baseline/_ref /root/reference``
# (git-ignored; it travels to the GPU box with the snapshot, /root/reference does not)
_INSTALLED = os.path.join(_REPO, "baseline", "_ref", "auto_oo")
_SRC = os.path.join(REF_ROOT, "src", "auto_oo")
if not os.path.isfile(os.path.join(_SRC, "oo_energy.py")) and os.path.isfile(os.path.join(_INSTALLED, "oo_energy.py")):
    _SRC = _INSTALLED


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_SRC, "oo_energy.py"))


def _as_torch(x):
    if torch.is_tensor(x):
        return x
    return torch.as_tensor(np.asarray(x))


def _build_math():
    m = types.ModuleType("pennylane.math")

    def _like(like):
        return like if isinstance(like, str) else None

    def array(x, like=None, **kw):
        if _like(like) == 'torch':
            return x.clone() if torch.is_tensor(x) else torch.tensor(np.asarray(x))
        return np.array(x)

    def zeros(shape, like=None, **kw):
        return torch.zeros(shape, dtype=torch.float64) if _like(like) == 'torch' else np.zeros(shape)

    def ones(shape, like=None, **kw):
        return torch.ones(shape, dtype=torch.float64) if _like(like) == 'torch' else np.ones(shape)

    def eye(n, like=None, **kw):
        return torch.eye(n, dtype=torch.float64) if _like(like) == 'torch' else np.eye(n)

    def convert_like(a, b):
        if torch.is_tensor(b):
            return _as_torch(a).to(b.device)
        return np.asarray(a)

    def get_interface(x):
        return 'torch' if torch.is_tensor(x) else 'numpy'

    def einsum(spec, *ops, **kw):
        if any(torch.is_tensor(o) for o in ops):
            ops = [_as_torch(o) for o in ops]
            return torch.einsum(spec.replace(' ', ''), *ops)
        return np.einsum(spec, *ops)

    def set_index(a, idx, v):
        if torch.is_tensor(a):
            if isinstance(idx, tuple):
                idx = tuple(torch.as_tensor(np.asarray(i)) if not torch.is_tensor(i) else i
                            for i in idx)
            elif not torch.is_tensor(idx):
                idx = torch.as_tensor(np.asarray(idx))
            a[idx] = v
            return a
        a[idx] = v
        return a

    def transpose(a, axes=None):
        if torch.is_tensor(a):
            return a.T if axes is None else a.permute(*axes)
        return np.transpose(a, axes)

    def sum_(a, axis=None, **kw):
        if torch.is_tensor(a):
            return torch.sum(a) if axis is None else torch.sum(a, dim=axis)
        return np.sum(a, axis=axis)

    def shape(a):
        return tuple(a.shape)

    def zeros_like(a):
        return torch.zeros_like(a) if torch.is_tensor(a) else np.zeros_like(a)

    def reshape(a, s):
        return a.reshape(s)

    def flatten(a):
        return a.reshape(-1)

    def concatenate(xs, axis=0):
        if any(torch.is_tensor(x) for x in xs):
            return torch.cat([_as_torch(x) for x in xs], dim=axis)
        return np.concatenate(xs, axis=axis)

    def dot(a, b):
        if torch.is_tensor(a) or torch.is_tensor(b):
            return torch.dot(_as_torch(a), _as_torch(b))
        return np.dot(a, b)

    def diag(a):
        return torch.diag(a) if torch.is_tensor(a) else np.diag(a)

    def expm(a):
        if torch.is_tensor(a):
            return torch.linalg.matrix_exp(a)
        import scipy.linalg
        return scipy.linalg.expm(a)

    def allclose(a, b, rtol=1e-5, atol=1e-8):
        a = a.detach().numpy() if torch.is_tensor(a) else np.asarray(a)
        b = b.detach().numpy() if torch.is_tensor(b) else np.asarray(b)
        return np.allclose(a, b, rtol=rtol, atol=atol)

    def prod(a):
        return int(np.prod(a))

    def min_(a):
        return a.min()

    def max_(a):
        return a.max()

    def abs_(a):
        return a.abs() if torch.is_tensor(a) else np.abs(a)

    def log(a):
        return torch.log(a) if torch.is_tensor(a) else np.log(a)

    linalg = types.SimpleNamespace(
        eigh=lambda a: torch.linalg.eigh(a) if torch.is_tensor(a) else np.linalg.eigh(a))

    for name, fn in dict(
            array=array, zeros=zeros, ones=ones, eye=eye, convert_like=convert_like,
            get_interface=get_interface, einsum=einsum, set_index=set_index,
            transpose=transpose, sum=sum_, shape=shape, zeros_like=zeros_like,
            reshape=reshape, flatten=flatten, concatenate=concatenate, dot=dot, diag=diag,
            expm=expm, allclose=allclose, prod=prod, min=min_, max=max_, abs=abs_, log=log,
            ix_=np.ix_, linalg=linalg).items():
        setattr(m, name, fn)
    return m


_loaded = None


def load_reference():
    """Returns a namespace with the verbatim reference modules
    (``oo_energy``, ``active_space``, ``newton_raphson``, ``moldata_pyscf``)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REF_ROOT} or {_INSTALLED}")

    saved = {k: sys.modules.get(k) for k in
             ("pennylane", "pennylane.math", "pyscf", "openfermion",
              "auto_oo", "auto_oo.utils", "auto_oo.moldata_pyscf",
              "auto_oo.utils.active_space", "auto_oo.utils.newton_raphson",
              "auto_oo.oo_energy")}

    pl = types.ModuleType("pennylane")
    pl.math = _build_math()
    pl.__path__ = []
    sys.modules["pennylane"] = pl
    sys.modules["pennylane.math"] = pl.math
    for stub in ("pyscf", "openfermion"):
        mod = types.ModuleType(stub)
        mod.__path__ = []
        sys.modules[stub] = mod
    # moldata_pyscf does `from pyscf import gto, scf, mcscf, fci`
    for sub in ("gto", "scf", "mcscf", "fci"):
        setattr(sys.modules["pyscf"], sub, types.ModuleType(f"pyscf.{sub}"))

    pkg = types.ModuleType("auto_oo")
    pkg.__path__ = [_SRC]
    sys.modules["auto_oo"] = pkg
    upkg = types.ModuleType("auto_oo.utils")
    upkg.__path__ = [os.path.join(_SRC, "utils")]
    sys.modules["auto_oo.utils"] = upkg

    def _load(modname, relpath):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(_SRC, relpath))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    ns = types.SimpleNamespace()
    ns.moldata_pyscf = _load("auto_oo.moldata_pyscf", "moldata_pyscf.py")
    ns.active_space = _load("auto_oo.utils.active_space", os.path.join("utils", "active_space.py"))
    ns.newton_raphson = _load("auto_oo.utils.newton_raphson",
                              os.path.join("utils", "newton_raphson.py"))
    ns.oo_energy = _load("auto_oo.oo_energy", "oo_energy.py")
    ns.math = pl.math
    _loaded = ns
    # leave the auto_oo.* entries registered (the reference modules refer to each other lazily);
    # restore third-party names so the stubs never shadow a real install
    for k in ("pennylane", "pennylane.math", "pyscf", "openfermion"):
        if saved[k] is not None:
            sys.modules[k] = saved[k]
    return ns


_drivers = {}


def load_reference_oo_pqc(oo_energy_module=None):
    """The VERBATIM ``oo_pqc.py`` of the reference (``OO_pqc`` with ``full_optimization``, ``full_gradient``,
    ``full_hessian`` ...; oo_pqc.py:30-207) executed with ``auto_oo.oo_energy`` bound to ``oo_energy_module``:
    ``None`` -> the verbatim reference module (an all-reference CPU run); ``auto_oo_b200.oo_energy`` -> the
    reference's own driver class derived from the CUDA ``OO_energy`` ("the drivers run unchanged").
    ``auto_oo.utils.newton_raphson`` is the verbatim module in both cases; ``auto_oo.pqc`` (PennyLane circuits,
    out of scope) is a stub that only provides the ``Parameterized_circuit`` name used in an annotation."""
    ref = load_reference()
    key = "reference" if oo_energy_module is None else oo_energy_module.__name__
    if key in _drivers:
        return _drivers[key]
    target = ref.oo_energy if oo_energy_module is None else oo_energy_module
    names = ("pennylane", "pennylane.math", "auto_oo.oo_energy", "auto_oo.pqc")
    saved = {k: sys.modules.get(k) for k in names}
    pl = types.ModuleType("pennylane")
    pl.math = ref.math
    pl.__path__ = []
    pqc_stub = types.ModuleType("auto_oo.pqc")
    pqc_stub.Parameterized_circuit = type("Parameterized_circuit", (), {})
    sys.modules.update({"pennylane": pl, "pennylane.math": ref.math, "auto_oo.oo_energy": target,
                        "auto_oo.pqc": pqc_stub})
    try:
        modname = "auto_oo.oo_pqc__" + key.replace(".", "_")
        spec = importlib.util.spec_from_file_location(modname, os.path.join(_SRC, "oo_pqc.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
    finally:
        for k in names:
            if saved[k] is not None:
                sys.modules[k] = saved[k]
            else:
                sys.modules.pop(k, None)
        sys.modules["auto_oo.oo_energy"] = ref.oo_energy
    _drivers[key] = mod
    return mod


class FakeMol:
    """Duck-typed stand-in for ``Moldata_pyscf`` (moldata_pyscf.py:19-56): the
    attributes ``OO_energy.__init__`` reads (oo_energy.py:154-165)."""

    def __init__(self, int1e_ao, int2e_ao, overlap, oao_coeff, nuc, nelec):
        self.int1e_ao = np.asarray(int1e_ao)
        self.int2e_ao = np.asarray(int2e_ao)
        self.overlap = np.asarray(overlap)
        self.oao_coeff = np.asarray(oao_coeff)
        self.nuc = float(nuc)
        self.nao = self.int1e_ao.shape[0]
        self.nelectron = int(nelec)

    def get_active_space_idx(self, ncas, nelecas):
        nelecore = self.nelectron - nelecas
        if nelecore % 2 == 1:
            raise ValueError('odd number of core electrons')
        occ_idx = np.arange(nelecore // 2)
        act_idx = (occ_idx[-1] + 1 + np.arange(ncas) if len(occ_idx) > 0 else np.arange(ncas))
        virt_idx = np.arange(act_idx[-1] + 1, self.nao)
        return occ_idx, act_idx, virt_idx
