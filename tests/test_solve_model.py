"""CPU model of the solve kernel's elimination (RegSolve in k3d_kernels.cu).

The kernel borders the (n + mdt) kriging system with the right-hand side b and the data vector
v, stores it mirrored packed-lower, spreads it 2-D cyclic over the 32 lanes of a warp (entry
(r, c) -> lane (r & 3) * 8 + (c & 7), register m[r >> 2][c >> 3]) and eliminates the pivots
R .. 2 without pivoting, handing each pivot row to the other lanes through a permuted shared
row; what is left in entries (1, 1) and (1, 0) is the kriging variance and minus the estimate.
This restates that data movement lane by lane in numpy -- same layout, same permuted row, same
blocks skipped -- and checks it against numpy.linalg.solve on the systems krige3d.py:470-614
builds (test infrastructure; the product has no CPU path)."""
import numpy as np
import pytest


def rowpos(c):
    return (c & 3) * 12 + ((c >> 2) & 1) * 6 + (c >> 3)


def pk(r, c):
    return r * (r + 1) // 2 + c


def bordered_packed(A, b, v, cbb):
    """The kernel's layout: row 0 = v, row 1 = b, then the constraint rows, then the samples
    in reverse (sample i in row R - i); mirrored packed lower triangle."""
    N = A.shape[0]
    R = N + 1
    full = np.zeros((R + 1, R + 1))
    idx = [R - i for i in range(N)]          # system row/col i -> bordered row
    for i in range(N):
        for j in range(N):
            full[idx[i], idx[j]] = A[i, j]
        full[idx[i], 1] = full[1, idx[i]] = b[i]
        full[idx[i], 0] = full[0, idx[i]] = v[i]
    full[1, 1] = cbb
    Pm = np.zeros(pk(R, R) + 1)
    for r in range(R + 1):
        for c in range(r + 1):
            Pm[pk(r, c)] = full[r, c]
    return Pm, R


def regsolve_model(Pm, R, AMAX):
    """RegSolve<AMAX>::run lane by lane."""
    BMAX = ((4 * (AMAX - 1) + 3) >> 3) + 1
    m = np.zeros((32, AMAX, BMAX))
    for lane in range(32):
        li, lj = lane >> 3, lane & 7
        for a in range(AMAX):
            for b in range(BMAX):
                if 8 * b <= 4 * a + 3:
                    r, c = li + 4 * a, lj + 8 * b
                    m[lane, a, b] = Pm[pk(r, c)] if (c <= r and r <= R) else 0.0
    minpiv = np.inf
    for M in range(R, 1, -1):
        AP, q = M >> 2, M & 3
        BW = ((4 * AP + 3) >> 3) + 1
        row = np.full(48, np.nan)
        for lane in range(32):               # the owner lanes publish the row, permuted
            li, lj = lane >> 3, lane & 7
            if li == q:
                for b in range(BW):
                    row[(lj & 3) * 12 + (lj >> 2) * 6 + b] = m[lane, AP, b]
        d = row[rowpos(M)]
        minpiv = min(minpiv, abs(d))
        ninv = -1.0 / d
        A = AP if q == 0 else AP + 1          # the first row of a block skips that block
        if A == 0:
            continue
        B = ((4 * (A - 1) + 3) >> 3) + 1
        for lane in range(32):
            li, lj = lane >> 3, lane & 7
            u = [row[(lj & 3) * 12 + (lj >> 2) * 6 + b] * ninv for b in range(B)]
            for a in range(A):
                w = row[li * 12 + (a & 1) * 6 + (a >> 1)]   # row li + 4 a of the lane's run
                assert rowpos(li + 4 * a) == li * 12 + (a & 1) * 6 + (a >> 1)
                for b in range(B):
                    if 8 * b <= 4 * a + 3:
                        m[lane, a, b] += w * u[b]
    var = m[9, 0, 0]                          # (r, c) = (1, 1): lane 1 * 8 + 1
    est = -m[8, 0, 0]                         # (r, c) = (1, 0): lane 1 * 8 + 0
    return est, var, minpiv


def kriging_system(rng, n, mdt):
    """A positive definite covariance block with mdt constraint rows (unbiasedness + drift
    monomials), as krige3d.py:748-801 and 493-583 build them."""
    p = rng.uniform(0, 10, (n, 3))
    h = np.linalg.norm(p[:, None, :] - p[None, :, :], axis=2)
    C = 0.1 * (h == 0) + 0.9 * np.exp(-3.0 * h / 8.0)
    x0 = rng.uniform(3, 7, 3)
    b = 0.9 * np.exp(-3.0 * np.linalg.norm(p - x0, axis=1) / 8.0)
    N = n + mdt
    A = np.zeros((N, N))
    A[:n, :n] = C
    rhs = np.zeros(N)
    rhs[:n] = b
    if mdt > 0:
        A[n, :n] = A[:n, n] = 1.0
        rhs[n] = 1.0
        x, y, z = ((p - x0) * 0.2).T          # krige3d.py:493-583: the nine drift monomials
        mono = [x, y, z, x * x, y * y, z * z, x * y, x * z, y * z]
        for k in range(mdt - 1):
            A[n + 1 + k, :n] = A[:n, n + 1 + k] = mono[k]
    v = np.zeros(N)
    v[:n] = rng.normal(size=n)
    return A, rhs, v


@pytest.mark.parametrize("n,mdt,AMAX", [(32, 1, 9), (16, 0, 5), (32, 10, 11), (7, 1, 5), (30, 4, 9),
                                        (1, 1, 5), (2, 0, 5), (18, 0, 5), (33, 1, 9), (31, 10, 11)])
def test_register_elimination_model_against_numpy(n, mdt, AMAX):
    rng = np.random.default_rng(100 * n + mdt)
    A, b, v = kriging_system(rng, n, mdt)
    cbb = 1.0
    Pm, R = bordered_packed(A, b, v, cbb)
    assert R <= 4 * AMAX - 1
    est, var, minpiv = regsolve_model(Pm, R, AMAX)
    s = np.linalg.solve(A, b)
    assert abs(est - s @ v) <= 1e-10 * max(1.0, abs(s @ v))
    assert abs(var - (cbb - s @ b)) <= 1e-10
    assert minpiv > 1e-9


def test_permuted_row_is_a_bijection_with_aligned_runs():
    pos = [rowpos(c) for c in range(48)]
    assert sorted(pos) == list(range(48))
    for lj in range(8):       # a lane's columns lj, lj + 8, ... are contiguous and 16-byte aligned
        run = [rowpos(lj + 8 * b) for b in range(6)]
        assert run == list(range(run[0], run[0] + 6)) and run[0] % 2 == 0
    for li in range(4):       # its rows li + 8 e and li + 4 + 8 e are the runs of lanes li, li + 4
        assert [rowpos(li + 8 * e) for e in range(6)] == [rowpos(li) + e for e in range(6)]
        assert [rowpos(li + 4 + 8 * e) for e in range(6)] == [rowpos(li + 4) + e for e in range(6)]


def test_duplicate_samples_give_a_tiny_pivot():
    """Repair D16: two identical samples leave a pivot of rounding size, which the kernel
    compares with 1e-12 x maxcov."""
    rng = np.random.default_rng(5)
    A, b, v = kriging_system(rng, 12, 1)
    A[3, :] = A[5, :]
    A[:, 3] = A[:, 5]
    b[3] = b[5]
    Pm, R = bordered_packed(A, b, v, 1.0)
    with np.errstate(all='ignore'):
        _, _, minpiv = regsolve_model(Pm, R, 5)
    assert minpiv <= 1e-12
