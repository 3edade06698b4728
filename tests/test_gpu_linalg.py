"""GPU checks of the batched one-sided Jacobi SVD behind `factorize_svd` in simple_update_bp (src/apply.jl:81-88).

Reference for the numbers: LAPACK through NumPy (the reference calls LAPACK through NDTensors' `svd`).  Singular values
must agree to 1e-12 relative to the largest one (the gate path's tolerance is 1e-10); the specialised m, n <= 64 kernels
(variant 0: odd-even ordering, variant 2: round-robin) are also compared with the shape-generic kernel (variant 1) on the
same bytes."""
import numpy as np
import pytest

import itn_b200 as E

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return E.Context(0)


def random_batch(rng, b, m, n, dtype, decay=None):
    a = rng.standard_normal((b, m, n))
    if np.dtype(dtype).kind == "c":
        a = a + 1j * rng.standard_normal((b, m, n))
    if decay is not None:  # graded columns: singular values spread over many orders of magnitude
        a = a * (decay ** np.arange(n))[None, None, :]
    return a.astype(dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("shape", [(64, 64), (64, 63), (33, 17), (16, 64), (64, 2), (5, 3), (64, 32), (48, 64), (7, 1)])
def test_singular_values_match_lapack(ctx, dtype, shape):
    m, n = shape
    rng = np.random.default_rng(100 * m + n)
    a = random_batch(rng, 5, m, n, dtype)
    ref = np.stack([np.concatenate([np.linalg.svd(x, compute_uv=False), np.zeros(max(0, n - m))]) for x in a])
    for variant in (0, 1, 2, 3):
        sig, us, _ = E.svd_batch(a, variant=variant, want_us=True, ctx=ctx)
        scale = ref[:, :1]
        assert np.max(np.abs(sig - ref) / scale) < 1e-12, (variant, np.max(np.abs(sig - ref) / scale))
        for x, u, s in zip(a, us, sig):
            # U Sigma = A V with V unitary: same Gram matrix on the row side, orthogonal columns with norms sigma
            assert np.linalg.norm(u @ u.conj().T - x @ x.conj().T) < 1e-11 * np.linalg.norm(x) ** 2
            g = u.conj().T @ u
            off = g - np.diag(np.diag(g))
            assert np.linalg.norm(off) < 1e-11 * np.linalg.norm(x) ** 2
            assert np.allclose(np.sort(np.sqrt(np.abs(np.diag(g))))[::-1], s, rtol=0, atol=1e-11 * s[0])


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("shape", [(128, 64), (64, 128), (100, 50), (33, 100), (128, 7), (128, 128), (70, 70)])
def test_bond_matrices_up_to_128(ctx, dtype, shape):
    # heavy-hex chi = 32 bond matrices: 128 x 64 and 64 x 128 run on the m <= 128 / n <= 128 instances of the odd-even
    # kernel (a wide matrix has n - m null columns), anything larger in both directions on the shape-generic kernel
    m, n = shape
    rng = np.random.default_rng(7 * m + n)
    a = random_batch(rng, 3, m, n, dtype)
    ref = np.stack([np.concatenate([np.linalg.svd(x, compute_uv=False), np.zeros(max(0, n - m))]) for x in a])
    for variant in (0, 1):
        sig, us, _ = E.svd_batch(a, variant=variant, want_us=True, ctx=ctx)
        assert np.max(np.abs(sig - ref) / ref[:, :1]) < 1e-12, (variant, np.max(np.abs(sig - ref) / ref[:, :1]))
        for x, u in zip(a, us):
            assert np.linalg.norm(u @ u.conj().T - x @ x.conj().T) < 1e-11 * np.linalg.norm(x) ** 2


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_graded_and_rank_deficient(ctx, dtype):
    rng = np.random.default_rng(5)
    a = random_batch(rng, 4, 64, 64, dtype, decay=0.6)  # sigma_max / sigma_min ~ 1e14
    a[1, :, 40:] = 0.0                                   # exact zero columns
    a[2] = a[2, :, :1] @ np.ones((1, 64))                # rank one
    a[3] = 0.0                                           # zero matrix
    ref = np.stack([np.linalg.svd(x, compute_uv=False) for x in a])
    s0, _, _ = E.svd_batch(a, variant=0, ctx=ctx)
    s1, _, _ = E.svd_batch(a, variant=1, ctx=ctx)
    s2, _, _ = E.svd_batch(a, variant=2, ctx=ctx)
    scale = np.maximum(ref[:, :1], 1e-300)
    assert np.max(np.abs(s0 - ref) / scale) < 1e-12
    assert np.max(np.abs(s0 - s1) / scale) < 1e-12
    assert np.max(np.abs(s0 - s2) / scale) < 1e-12


def test_gate_sized_batch_matches_generic_kernel(ctx):
    rng = np.random.default_rng(11)
    a = random_batch(rng, 300, 64, 64, np.complex128)
    s0, _, t0 = E.svd_batch(a, variant=0, ctx=ctx)
    s1, _, t1 = E.svd_batch(a, variant=1, ctx=ctx)
    s2, _, t2 = E.svd_batch(a, variant=2, ctx=ctx)
    assert np.max(np.abs(s0 - s1)) < 1e-12 * np.max(s1)
    assert np.max(np.abs(s0 - s2)) < 1e-12 * np.max(s1)
    ref = np.linalg.svd(a[:8], compute_uv=False)
    assert np.max(np.abs(s0[:8] - ref)) < 1e-12 * np.max(ref)
    print(f"\n300 x (64 x 64 c128): odd-even {t0:.3f} ms, round-robin {t2:.3f} ms, generic {t1:.3f} ms")
