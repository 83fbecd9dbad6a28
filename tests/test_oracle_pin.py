"""The OpenCV primitives restated in oracle/tri_oracle.c against the real ones in the cv2 wheel:
cv::invert(DECOMP_SVD) (one-sided Jacobi), cv::solve / cv::invert(DECOMP_EIG) (Jacobi eigen solver)."""
import ctypes as C

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
import oracle_py as O  # noqa: E402


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("n_views", [2, 3, 4, 8, 16, 32])
def test_pinv_svd_matches_cv_invert(n_views):
    rng = np.random.default_rng(n_views)
    for trial in range(20):
        A = rng.normal(0, 1000.0, (2 * n_views, 3)) + rng.normal(0, 1.0, (2 * n_views, 3)) * 1e-3
        want = cv2.invert(A, flags=cv2.DECOMP_SVD)[1]
        got = np.zeros((3, 2 * n_views))
        O.lib().orc_pinv_svd(_p(np.ascontiguousarray(A)), 2 * n_views, 3, _p(got))
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


def test_eig_solve_and_inverse_diag_match_cv():
    rng = np.random.default_rng(5)
    for trial in range(200):
        J = rng.normal(0, 1.0, (rng.integers(2, 9), 3))
        A = J.T @ J + (np.eye(3) * rng.uniform(0, 2) if trial % 3 else 0)  # incl. rank-deficient (2 rays)
        A = (A + A.T) / 2
        b = rng.normal(0, 1.0, (3, 1))
        want = cv2.solve(A, b, flags=cv2.DECOMP_EIG)[1].ravel()
        got = np.zeros(3)
        O.lib().orc_eig_solve3(_p(np.ascontiguousarray(A)), _p(np.ascontiguousarray(b)), _p(got))
        assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max()), trial
        wd = np.diag(cv2.invert(A, flags=cv2.DECOMP_EIG)[1])
        gd = np.zeros(3)
        O.lib().orc_eig_inv_diag3(_p(np.ascontiguousarray(A)), _p(gd))
        assert np.abs(gd - wd).max() <= 1e-12 * max(1.0, np.abs(wd).max()), trial
