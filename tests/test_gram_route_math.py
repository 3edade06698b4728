"""The linear algebra behind the engine's simple update (csrc/itn_linalg.cu), restated in NumPy -- no device needed.

The reference QR-factorises the absorbed site tensor (src/apply.jl:70-76).  The engine works on Gram matrices instead and
repairs their squared condition number in two places; these tests pin the identities those kernels rely on:

  * CholeskyQR2 (k_chol twice + k_su_compose): R = R2 R1 with R1 = chol(A^H A), R2 = chol((A R1^-1)^H (A R1^-1)) restores
    kappa eps where one pass gives kappa^2 eps;
  * thin sides (k_thin_pinv2): R^+ = Y (Y^H Y)^-1 Gp^T with Y = R^H conj(Gp) is a right inverse of R for ANY invertible Gp,
    so an inaccurate Cholesky factor of R R^H costs nothing.
"""
import numpy as np
import pytest


def graded(rng, m, n, kappa, cplx=True):
    def rnd(a, b):
        x = rng.standard_normal((a, b))
        return x + 1j * rng.standard_normal((a, b)) if cplx else x
    r = min(m, n)
    u, _ = np.linalg.qr(rnd(m, r))
    v, _ = np.linalg.qr(rnd(n, r))
    return (u * np.logspace(0, -np.log10(kappa), r)) @ v.conj().T


@pytest.mark.parametrize("kappa", [1e2, 1e4, 1e6])
def test_cholesky_qr2_restores_orthogonality(kappa):
    rng = np.random.default_rng(5)
    a = graded(rng, 400, 12, kappa)
    r1 = np.linalg.cholesky(a.conj().T @ a).conj().T          # A^H A = R1^H R1
    q1 = a @ np.linalg.inv(r1)                                 # the explicit inverse, as the rebuild kernel applies it
    e1 = np.linalg.norm(q1.conj().T @ q1 - np.eye(12))
    r2 = np.linalg.cholesky(q1.conj().T @ q1).conj().T
    r = r2 @ r1
    rp = np.linalg.inv(r1) @ np.linalg.inv(r2)                 # R^+ = R1^+ R2^+ (k_su_compose)
    q = a @ rp
    e2 = np.linalg.norm(q.conj().T @ q - np.eye(12))
    eps = np.finfo(float).eps
    assert e2 < 50 * kappa * eps                               # kappa eps from forming A R^+ with an explicit inverse
    if kappa >= 1e4:
        assert e1 > 30 * e2                                    # one pass: kappa^2 eps
    assert np.linalg.norm(q @ r - a) < 50 * kappa * eps * np.linalg.norm(a)
    # the Gram matrix of A is reproduced by the composed factor to working precision
    assert np.linalg.norm(r.conj().T @ r - a.conj().T @ a) < 1e-13 * np.linalg.norm(a) ** 2


@pytest.mark.parametrize("kappa", [1e2, 1e4, 1e6])
def test_thin_side_right_inverse_is_insensitive_to_the_cholesky_factor(kappa):
    rng = np.random.default_rng(9)
    x, n = 6, 16                                               # X outer states < n = d chi: a degree-2 site
    r = graded(rng, x, n, kappa)
    g = r @ r.conj().T
    # k_chol's convention: conj(G) = L L^H, Gp = L^-H, so that G^-1 = conj(Gp) Gp^T
    l = np.linalg.cholesky(g.conj())
    gp = np.linalg.inv(l).conj().T
    # spoil Gp the way a kappa^2-conditioned Cholesky would (relative 1e-6 here, far worse than rounding)
    gp_bad = gp * (1 + 1e-6 * rng.standard_normal(gp.shape))
    for gpx in (gp, gp_bad):
        y = r.conj().T @ gpx.conj()
        rp = y @ np.linalg.inv(y.conj().T @ y) @ gpx.T
        assert np.linalg.norm(r @ rp - np.eye(x)) < 100 * kappa * np.finfo(float).eps
        # columns of R^+ stay in the row space of R: it is the Moore-Penrose inverse
        assert np.linalg.norm(rp - np.linalg.pinv(r)) < 1e-6 * np.linalg.norm(np.linalg.pinv(r))
    # the plain formula R^H G^-1 with the spoiled factor is off by the size of the perturbation
    plain = r.conj().T @ (gp_bad.conj() @ gp_bad.T)
    assert np.linalg.norm(r @ plain - np.eye(x)) > 1e-7


def test_pivot_ratio_bounds_the_condition_number_from_below():
    # k_chol flags a side for the second pass when min pivot / max pivot < 1e-4; the ratio of the Cholesky pivots never
    # exceeds kappa(C), so a flagged matrix is ill conditioned (no false alarms on well-conditioned sides)
    rng = np.random.default_rng(13)
    for kappa in (3.0, 30.0, 1e3, 1e5):
        a = graded(rng, 200, 10, kappa)
        c = a.conj().T @ a
        d = np.real(np.diag(np.linalg.cholesky(c))) ** 2
        assert d.max() / d.min() <= np.linalg.cond(c) * (1 + 1e-8)
