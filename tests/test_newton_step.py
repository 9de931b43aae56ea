"""CPU: our NewtonStep against known answers of the verbatim reference (tests/golden/newton_step.npz,
made by oracle/make_golden.py) and the reference's own synthetic objectives
(test/utils/test_newton_raphson.py:99-130)."""
import os

import numpy as np
import pytest
import torch

from auto_oo_b200.utils.newton_raphson import NewtonStep, split_list_shapes
from auto_oo_b200.oo_energy import vector_to_skew_symmetric, skew_symmetric_to_vector
from helpers import GOLDEN


def test_newton_step_known_answer():
    d = np.load(os.path.join(GOLDEN, "newton_step.npz"))
    a, x0 = torch.as_tensor(d["a"]), torch.as_tensor(d["x0"])

    def f(x):
        return 0.5 * x @ a @ x + 0.25 * torch.sum(x ** 4) + torch.sum(x)

    opt = NewtonStep(verbose=0)
    dp, lam = opt.newton_step(torch.as_tensor(d["grad"]), torch.as_tensor(d["hess"]))
    assert abs(lam - float(d["lowest"])) < 1e-12
    assert np.abs(dp.numpy() - d["dp"]).max() < 1e-10
    newx, lam2 = opt.damped_newton_step(f, (x0,), torch.as_tensor(d["grad"]), torch.as_tensor(d["hess"]))
    assert np.abs(newx.numpy() - d["new_x"]).max() < 1e-10
    assert lam2 == lam


def test_split_list_shapes():
    chunks = split_list_shapes(torch.arange(10.), [(2, 3), (4,)])
    assert chunks[0].shape == (2, 3) and chunks[1].shape == (4,)
    assert torch.equal(chunks[1], torch.arange(6., 10.))


def test_diagonalise_by_rotations():
    """reference function_type_a: minimise the off-diagonal weight of U^T A U over kappa."""
    torch.manual_seed(0)
    for dim in (2, 4, 8):
        A = torch.randn(dim, dim, dtype=torch.float64)
        A = A + A.T
        target = torch.linalg.eigvalsh(A)

        def cost(x):
            U = torch.linalg.matrix_exp(-vector_to_skew_symmetric(x))
            M = U.T @ A @ U
            return -torch.sum(torch.diagonal(M) * torch.arange(1, dim + 1, dtype=torch.float64))

        x = torch.zeros(dim * (dim - 1) // 2, dtype=torch.float64)
        Utot = torch.eye(dim, dtype=torch.float64)
        opt = NewtonStep(verbose=0)
        for _ in range(60):
            Ar = Utot.T @ A @ Utot

            def c(x):
                U = torch.linalg.matrix_exp(-vector_to_skew_symmetric(x))
                return -torch.sum(torch.diagonal(U.T @ Ar @ U) * torch.arange(1, dim + 1, dtype=torch.float64))

            g = torch.autograd.functional.jacobian(c, x)
            if g.abs().max() < 1e-10:
                break
            h = torch.autograd.functional.hessian(c, x)
            step, _ = opt.damped_newton_step(c, (x,), g, h)
            Utot = Utot @ torch.linalg.matrix_exp(-vector_to_skew_symmetric(step))
        diag = torch.diagonal(Utot.T @ A @ Utot)
        assert torch.allclose(torch.sort(diag).values, target, atol=1e-8)


def test_log_barrier_line_search():
    """reference function_type_b: 1-D objective where the full Newton step must be damped."""
    def f(x):
        return x[0] ** 2 - torch.log(1.0 - x[0]) - torch.log(1.0 + x[0]) + 3.0 * x[0]

    x = torch.tensor([0.9], dtype=torch.float64)
    opt = NewtonStep(verbose=0)
    for _ in range(30):
        g = torch.autograd.functional.jacobian(f, x)
        h = torch.autograd.functional.hessian(f, x)
        x, _ = opt.damped_newton_step(f, (x,), g, h)
        assert abs(x.item()) < 1.0
    assert torch.autograd.functional.jacobian(f, x).abs().item() < 1e-8


def test_speculative_line_search_equals_sequential():
    """speculate=k evaluates k step lengths per batched call but must accept the same step.
    f = sum sqrt(1 + x^2): the pure Newton step -x^3 overshoots for |x| > 1, so the search must damp."""
    def f(x):
        return torch.sum(torch.sqrt(1.0 + x ** 2))

    calls = []

    def batched(plist):
        calls.append(len(plist))
        return torch.stack([f(p[0]) for p in plist])

    f.batched = batched
    x_seq = torch.tensor([3.0, -2.0, 1.5], dtype=torch.float64)
    x_spec = x_seq.clone()
    for _ in range(10):
        g = torch.autograd.functional.jacobian(f, x_seq)
        h = torch.autograd.functional.hessian(f, x_seq)
        x_seq, _ = NewtonStep(verbose=0).damped_newton_step(f, (x_seq,), g, h)
        g = torch.autograd.functional.jacobian(f, x_spec)
        h = torch.autograd.functional.hessian(f, x_spec)
        x_spec, _ = NewtonStep(verbose=0, speculate=4).damped_newton_step(f, (x_spec,), g, h)
        assert torch.equal(x_seq, x_spec)
    assert calls and max(calls) <= 4
    assert x_seq.abs().max().item() < 1e-6


@pytest.mark.parametrize("lmax,accepted", [(3, 0.0), (4, 2.0 ** -4)])
def test_speculative_line_search_gives_up_where_the_sequential_one_does(lmax, accepted):
    """f = x^2 from x = 1 along dp = -16: the first step length that decreases f is beta^4.  With lmax = 3 the
    sequential search (reference newton_raphson.py:165-176) evaluates beta^4 but reports failure (t = 0); the
    batched search must not accept it either."""
    def f(x):
        return torch.sum(x ** 2)

    f.batched = lambda plist: torch.stack([f(p[0]) for p in plist])
    x = torch.tensor([1.0], dtype=torch.float64)
    g = torch.tensor([2.0], dtype=torch.float64)
    dp = torch.tensor([-16.0], dtype=torch.float64)
    for spec in (0, 2, 4, 8):
        newx, _ = NewtonStep(verbose=0, lmax=lmax, speculate=spec).backtracking(f, (x,), dp, g)
        assert torch.equal(newx, x + accepted * dp), (spec, newx)
