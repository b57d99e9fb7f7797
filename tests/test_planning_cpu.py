"""CPU tests of the host-side planning logic (no GPU needed): the smoothness Objective against the dense
matrix the reference builds (planning.jl:7-20), the batched torch form of it, EqConst stacking of
ConfigurationConstraints (planning.jl:72-88, 140-176) against the oracle restatement, scipynize."""
import numpy as np
import torch

from kinematics_jl_b200 import planning as P
from oracle import ref_model as R


def test_objective_matches_reference_matrix():
    rng = np.random.default_rng(0)
    for n_wp, n_dof in ((3, 2), (10, 8), (64, 11)):
        w = rng.uniform(0.5, 2.0, n_dof)
        A = R.objective_matrix(n_wp, w)                       # planning.jl:7-20, restated in the oracle
        F = P.Objective(n_wp, w)
        assert np.array_equal(F.A, A)
        xi = rng.standard_normal(n_wp * n_dof)
        grad = np.zeros(n_wp * n_dof)
        val = F(xi, grad)
        val_o, grad_o = R.objective(A, xi)
        np.testing.assert_allclose(val, val_o, rtol=1e-14)
        np.testing.assert_allclose(grad, grad_o, rtol=1e-14, atol=1e-14)
        assert F(xi, np.zeros(0)) == val                       # empty grad: value only (planning.jl:26)


def test_smoothness_objective_batched_matches_dense_xi_A_xi():
    rng = np.random.default_rng(1)
    n_wp, n_dof, n_prob = 64, 8, 5
    w = rng.uniform(0.5, 2.0, n_dof)
    A = R.objective_matrix(n_wp, w)
    X = torch.as_tensor(rng.standard_normal((n_prob, n_wp, n_dof)))
    val, grad = P.smoothness_objective(X, n_wp, w)
    for p in range(n_prob):
        xi = X[p].reshape(-1).numpy()
        val_o, grad_o = R.objective(A, xi)
        np.testing.assert_allclose(float(val[p]), val_o, rtol=1e-12)
        np.testing.assert_allclose(grad[p].reshape(-1).numpy(), grad_o, rtol=1e-12, atol=1e-12)


def test_eq_const_configuration_rows_vs_oracle():
    rng = np.random.default_rng(2)
    n_wp, n_dof = 10, 8
    qs, qg, xi = rng.standard_normal(n_dof), rng.standard_normal(n_dof), rng.standard_normal(n_wp * n_dof)
    H = P.EqConst(n_wp, [P.ConfigurationConstraint(1, n_dof, qs), P.ConfigurationConstraint(n_wp, n_dof, qg)])
    val, jac = H(xi)

    class _M:                                                   # eq_const only needs n_dof_extra for config rows
        n_dof_extra = 0
    val_o, jac_o = R.eq_const(_M(), [None] * n_dof, xi, n_wp, [("config", 1, qs), ("config", n_wp, qg)])
    assert np.array_equal(val, val_o) and np.array_equal(jac, jac_o)
    assert jac.shape == (n_dof * n_wp, 2 * n_dof) and np.count_nonzero(jac) == 2 * n_dof
    # scipynize: value closure + transposed Jacobian (planning.jl:198-205); nloptize: negated (planning.jl:178-185)
    h, dh = P.scipynize(H)
    assert np.array_equal(h(xi), val) and np.array_equal(dh(xi), jac.T)
    nv, nj = P.nloptize(H)(xi)
    assert np.array_equal(nv, -val) and np.array_equal(nj, -jac)
    # batched form: row blocks only, expanded by dense_batch
    X = torch.as_tensor(rng.standard_normal((3, n_wp, n_dof)))
    Hb = P.EqConst(n_wp, [P.ConfigurationConstraint(1, n_dof, torch.as_tensor(qs)), P.ConfigurationConstraint(n_wp, n_dof, torch.as_tensor(qg))])
    vb, blocks = Hb(X)
    dense = Hb.dense_batch(blocks)
    for p in range(3):
        v1, j1 = H(X[p].reshape(-1).numpy())
        np.testing.assert_allclose(vb[p].numpy(), v1, rtol=0, atol=0)
        assert np.array_equal(dense[p].numpy(), j1)


def test_straight_trajectory_and_objective_scipynize():
    q0, q1 = np.zeros(4), np.arange(4.0)
    xi = P.create_straight_trajectory(q0, q1, 5)               # planning.jl:304-308
    np.testing.assert_allclose(xi.reshape(5, 4)[2], 0.5 * q1)
    F = P.Objective(5, np.ones(4))
    f, df = P.scipynize(F)
    assert abs(f(xi)) < 1e-25                                  # a straight line has zero acceleration
    np.testing.assert_allclose(df(xi), 0.0, atol=1e-12)
