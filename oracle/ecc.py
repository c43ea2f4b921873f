"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

numpy restatement of cv::findTransformECC (OpenCV video/src/ecc.cpp) for MOTION_AFFINE with
no input mask and gaussFiltSize = 5, exactly the call register_pixel makes
(cpp/lib/registration.cpp:43-64: 5-argument overload, criteria COUNT+EPS 50 / 1e-3).
OpenCV is not vendored in the reference; this follows its published algorithm and is pinned
against cv2.findTransformECC (tests/test_oracle_golden.py, tests/golden/ecc_golden.npz).

Steps (ecc.cpp): blur template and image with GaussianBlur(5x5, sigma 0) -> fixed kernel
[1,4,6,4,1]/16, BORDER_REFLECT_101; all-ones mask -> blur -> *0.5/0.95 -> u8 (stays 1);
gx = filter2D(I, [-.5,0,.5]), gy likewise transposed; per iteration: warp I, gx, gy
(INTER_LINEAR | WARP_INVERSE_MAP, the fixed-point model of upsp_oracle.c) and the mask
(INTER_NEAREST); masked mean / std (float64); zero-mean under the mask; affine Jacobian columns
[gx X, gy X, gx Y, gy Y, gx, gy]; H = J^T J (f32 6x6 from f64 dots); rho = <T~,I~>/(|T~||I~|);
lambda = (|I~|^2 - ip.H^-1 ip)/(corr - tp.H^-1 ip); dp = H^-1 J^T (lambda T~ - I~).
"""
from __future__ import annotations

import numpy as np

from . import oracle as orc

_K5 = np.array([1, 4, 6, 4, 1], np.float32) / np.float32(16)


def _reflect101(n, idx):
    idx = np.abs(idx)
    return np.where(idx >= n, 2 * (n - 1) - idx, idx)


def gaussian_blur5(img: np.ndarray) -> np.ndarray:
    """cv::GaussianBlur(img, (5,5), 0) on f32: separable [1,4,6,4,1]/16, row pass then column
    pass, BORDER_REFLECT_101, float accumulation."""
    img = np.asarray(img, np.float32)
    h, w = img.shape
    cols = _reflect101(w, np.arange(-2, w + 2))
    p = img[:, cols]
    tmp = np.zeros_like(img)
    for k in range(5):
        tmp += _K5[k] * p[:, k:k + w]
    rows = _reflect101(h, np.arange(-2, h + 2))
    p = tmp[rows, :]
    out = np.zeros_like(img)
    for k in range(5):
        out += _K5[k] * p[k:k + h, :]
    return out


def gradients(img: np.ndarray):
    """filter2D(img, -1, [-0.5, 0, 0.5]) and its transpose, BORDER_REFLECT_101."""
    h, w = img.shape
    c = _reflect101(w, np.arange(-1, w + 1))
    r = _reflect101(h, np.arange(-1, h + 1))
    gx = np.float32(0.5) * (img[:, c[2:]] - img[:, c[:-2]])
    gy = np.float32(0.5) * (img[r[2:], :] - img[r[:-2], :])
    return gx.astype(np.float32), gy.astype(np.float32)


def find_transform_ecc(template32: np.ndarray, input32: np.ndarray, max_iters=50, eps=1e-3,
                       return_trace=False):
    """Returns (M[2,3] f32, rho, n_iterations).  Raises ValueError where OpenCV throws."""
    T = gaussian_blur5(template32)
    I = gaussian_blur5(input32)
    gx, gy = gradients(I)
    h, w = T.shape
    X, Y = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    ones = np.ones(input32.shape, np.float32)
    M = np.eye(2, 3, dtype=np.float32)
    rho, last_rho = -1.0, -eps
    it = 0
    trace = []
    while it < max_iters and abs(rho - last_rho) >= eps:
        it += 1
        Iw = orc.warp_affine(I, M, 1)
        gxw = orc.warp_affine(gx, M, 1)
        gyw = orc.warp_affine(gy, M, 1)
        mask = orc.warp_affine(ones, M, 0) != 0
        cnt = int(mask.sum())
        imean = float(Iw[mask].astype(np.float64).sum() / cnt)
        tmean = float(T[mask].astype(np.float64).sum() / cnt)
        istd = np.sqrt(max((Iw[mask].astype(np.float64) ** 2).sum() / cnt - imean ** 2, 0.0))
        tstd = np.sqrt(max((T[mask].astype(np.float64) ** 2).sum() / cnt - tmean ** 2, 0.0))
        Iz = Iw.copy()
        Iz[mask] = Iw[mask] - np.float32(imean)
        Tz = np.zeros_like(T)
        Tz[mask] = T[mask] - np.float32(tmean)
        tnorm = np.sqrt(cnt * tstd * tstd)
        inorm = np.sqrt(cnt * istd * istd)
        J = [gxw * X, gyw * X, gxw * Y, gyw * Y, gxw, gyw]
        H = np.zeros((6, 6), np.float32)
        for i in range(6):
            for j in range(i, 6):
                H[i, j] = H[j, i] = np.float32(np.dot(J[i].ravel().astype(np.float64), J[j].ravel().astype(np.float64)))
        Hinv = np.linalg.inv(H.astype(np.float64)).astype(np.float32)
        corr = float(np.dot(Tz.ravel().astype(np.float64), Iz.ravel().astype(np.float64)))
        last_rho = rho
        rho = corr / (inorm * tnorm)
        if np.isnan(rho):
            raise ValueError("NaN encountered.")
        proj = lambda A: np.array([np.float32(np.dot(Jk.ravel().astype(np.float64), A.ravel().astype(np.float64)))
                                   for Jk in J], np.float32)
        ip, tp = proj(Iz), proj(Tz)
        iph = Hinv @ ip
        lam_n = inorm * inorm - float(np.dot(ip.astype(np.float64), iph.astype(np.float64)))
        lam_d = corr - float(np.dot(tp.astype(np.float64), iph.astype(np.float64)))
        if lam_d <= 0.0:
            raise ValueError("The algorithm stopped before its convergence.")
        lam = lam_n / lam_d
        err = (np.float32(lam) * Tz - Iz).astype(np.float32)
        dp = Hinv @ proj(err)
        M[0, 0] += dp[0]
        M[1, 0] += dp[1]
        M[0, 1] += dp[2]
        M[1, 1] += dp[3]
        M[0, 2] += dp[4]
        M[1, 2] += dp[5]
        trace.append((rho, M.copy()))
    if return_trace:
        return M, rho, it, trace
    return M, rho, it
