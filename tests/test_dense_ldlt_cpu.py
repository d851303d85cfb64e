"""The algorithm of dense_solve_kernel, executed on the CPU: blocked right-looking L D L' with 8-column
panels and UNSCALED columns (column k keeps c_ik = l_ik d_k, no square root), the right-hand side eliminated along with
the columns, and the 32-row blocked backward substitution — a line-by-line numpy model of the index arithmetic, checked
against numpy.linalg.solve on SPD systems of the sizes the kernel sees (not multiples of 8 or 32 included).  The CUDA
kernel itself is pinned by tests/test_ba_gpu.py (reference goldens, test_direct_and_iterative_reduced_solves_agree)."""
import numpy as np
import pytest

NB = 8


def model_solve(S, b):
    n = len(b)
    L = np.tril(S).astype(np.float64)          # the packed lower triangle of the kernel
    y = b.astype(np.float64).copy()
    invd = np.zeros(n); Lp = np.zeros((n, NB))
    for k0 in range(0, n, NB):
        nb = min(NB, n - k0)
        for c in range(nb):                    # panel: one column at a time
            kc = k0 + c
            dk = L[kc, kc]
            assert dk > 0
            inv = 1.0 / dk; invd[kc] = inv
            yk = y[kc]
            for i in range(kc + 1, n):
                li = L[i, kc] * inv
                Lp[i, c] = li; y[i] -= li * yk
                for sl in range(c + 1, nb):    # the panel columns to its right
                    j = k0 + sl
                    if j <= i:
                        L[i, j] -= li * L[j, kc]
        t0 = k0 + nb
        for j in range(t0, n):                 # trailing matrix: one pass per panel
            cj = np.array([L[j, k0 + c] if c < nb else 0.0 for c in range(NB)])
            for i in range(j, n):
                L[i, j] -= float(Lp[i] @ cj)
    acc = np.zeros(n)
    b1 = n
    while b1 > 0:                              # backward: 32 unknowns at a time from the bottom
        b0 = b1 - 32 if b1 > 32 else 0
        for kk in range(b1 - b0 - 1, -1, -1):
            k = b0 + kk
            y[k] = (y[k] - acc[k]) * invd[k]
            for i in range(b0, k):
                acc[i] += L[k, i] * y[k]
        for i in range(b0):
            acc[i] += sum(L[k, i] * y[k] for k in range(b0, b1))
        b1 -= 32
    return y


@pytest.mark.parametrize("n", [1, 7, 8, 9, 33, 68, 100])
def test_blocked_ldlt_model(n):
    rng = np.random.default_rng(n)
    J = rng.normal(size=(3 * n + 5, n))
    S = J.T @ J + 1e-3 * np.eye(n)
    S[n // 2, :] = 0; S[:, n // 2] = 0; S[n // 2, n // 2] = 1.0      # a constant coordinate: identity row, as the masks leave it
    b = rng.normal(size=n); b[n // 2] = 0.0
    x = model_solve(S, b)
    ref = np.linalg.solve(S, b)
    assert np.allclose(x, ref, rtol=1e-9, atol=1e-12 * np.abs(ref).max())
