"""Neighbour sets for the esti_plane pin (include/common_lib.h:236-269 -> Eigen colPivHouseholderQr, common_lib.h:252) and two
independent restatements of what Eigen computes: scipy's pivoted QR (LAPACK geqp3) and scipy lstsq (LAPACK gelsd, SVD)."""
import numpy as np
import scipy.linalg as sl


def neighbour_sets(n, seed=7):
    """[n, 5, 3] float32: points near random planes at |p| up to 500 m (in-plane spread 0.1 .. 0.6 m, out-of-plane noise 0 .. 3 cm),
    a tenth of them nearly collinear (second in-plane direction squeezed to 1e-3 .. 1e-6 of the first)."""
    rng = np.random.default_rng(seed)
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    u = np.cross(nrm, rng.normal(size=(n, 3)))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    v = np.cross(nrm, u)
    c = rng.uniform(-1, 1, (n, 3)) * rng.choice([1.0, 20.0, 150.0, 500.0], (n, 1))
    spread = rng.uniform(0.1, 0.6, (n, 1, 1))
    s = rng.uniform(-1, 1, (n, 5, 1)) * spread
    t = rng.uniform(-1, 1, (n, 5, 1)) * spread
    squeeze = np.ones((n, 1, 1))
    k = n // 10
    squeeze[:k] = 10.0 ** rng.uniform(-6, -3, (k, 1, 1))
    w = rng.normal(size=(n, 5, 1)) * rng.choice([0.0, 0.003, 0.01, 0.03], (n, 1, 1))
    p = c[:, None, :] + s * u[:, None, :] + t * squeeze * v[:, None, :] + w * nrm[:, None, :]
    return np.ascontiguousarray(p, np.float32)


def plane_qr_pivot(nb):
    """A x = -1 by Householder QR with column pivoting (LAPACK geqp3 through scipy): x -> (n, d) as esti_plane normalises it."""
    A = nb.astype(np.float64)
    Q, R, P = sl.qr(A, mode="economic", pivoting=True)
    y = sl.solve_triangular(R, Q.T @ (-np.ones(5)))
    x = np.zeros(3)
    x[P] = y
    nn = np.linalg.norm(x)
    return np.array([x[0] / nn, x[1] / nn, x[2] / nn, 1.0 / nn]), np.abs(np.diag(R))


def plane_svd(nb):
    x = sl.lstsq(nb.astype(np.float64), -np.ones(5), lapack_driver="gelsd")[0]
    nn = np.linalg.norm(x)
    return np.array([x[0] / nn, x[1] / nn, x[2] / nn, 1.0 / nn])
