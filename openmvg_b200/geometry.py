"""Python binding of the geometric-filter C-ABI (tests / bench plumbing).

Mirrors GeometricFilter_FMatrix_AC::Robust_estimation per pair (reference:
matching_image_collection/F_ACRobust.hpp:45-106) over a CSR of putative matches."""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import check, lib

_vp = ctypes.c_void_p


FUNDAMENTAL, HOMOGRAPHY = 0, 1


def fundamental_acransac(offsets, xI, xJ, image_size, precision: float = 4.0, max_iterations: int = 2048, device: int = 0):
    return acransac(FUNDAMENTAL, offsets, xI, xJ, image_size, precision, max_iterations, device)


def homography_acransac(offsets, xI, xJ, image_size, precision: float = 4.0, max_iterations: int = 2048, device: int = 0):
    """GeometricFilter_HMatrix_AC::Robust_estimation per pair (H_ACRobust.hpp:46-112); `F` of the result holds H."""
    return acransac(HOMOGRAPHY, offsets, xI, xJ, image_size, precision, max_iterations, device)


def acransac(model, offsets, xI, xJ, image_size, precision: float = 4.0, max_iterations: int = 2048, device: int = 0):
    """offsets[n_pairs+1]; xI, xJ [n_matches, 2] float64; image_size [n_pairs, 4] (wI, hI, wJ, hJ).
    -> list of dicts per pair: inliers (uint32 indices into the pair's matches), F [3,3], error_max, min_nfa"""
    off = np.ascontiguousarray(offsets, np.uint64); n_pairs = len(off) - 1
    xI = np.ascontiguousarray(xI, np.float64).reshape(-1, 2); xJ = np.ascontiguousarray(xJ, np.float64).reshape(-1, 2)
    sz = np.ascontiguousarray(image_size, np.int32).reshape(-1, 4)
    nm = int(off[-1]); assert len(xI) == nm == len(xJ) and len(sz) == n_pairs
    inl = np.zeros(max(nm, 1), np.uint32); ninl = np.zeros(max(n_pairs, 1), np.uint32); F = np.zeros((max(n_pairs, 1), 9)); st = np.zeros((max(n_pairs, 1), 2))
    p = lambda a: a.ctypes.data_as(_vp)   # noqa: E731
    check(lib().omvg_geom_acransac(int(device), int(model), ctypes.c_uint64(n_pairs), p(off), p(xI), p(xJ), p(sz), ctypes.c_double(precision),
                                               ctypes.c_uint32(max_iterations), p(inl), p(ninl), p(F), p(st)))
    out = []
    for k in range(n_pairs):
        a = int(off[k])
        out.append(dict(inliers=inl[a:a + int(ninl[k])].copy(), F=F[k].reshape(3, 3).copy(), error_max=float(st[k, 0]), min_nfa=float(st[k, 1])))
    return out
