"""GPU: the per-pair a-contrario RANSAC kernel (omvg_geom_fundamental_acransac) through the C ABI
against the oracle and the reference goldens: IDENTICAL inlier lists (same MT19937 sample sequence, same NFA
decisions), (errorMax, minNFA) to 1e-9, F equal up to scale."""
import numpy as np
import pytest

import checkers as ck
from openmvg_b200 import synth
from test_oracle_geom import GOLD, GOLD_H, case_inputs, case_inputs_h, check_against_gold

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def geometry():
    from openmvg_b200 import geometry as g
    return g


def test_reference_golden_cases_in_one_launch(geometry):
    """All golden pairs as ONE CSR call (one CTA per pair, pairs with <= 7 matches included)."""
    xs = [case_inputs(c) for c in GOLD]
    off = np.concatenate([[0], np.cumsum([len(x[0]) for x in xs])]).astype(np.uint64)
    res = geometry.fundamental_acransac(off, np.concatenate([x[0] for x in xs]), np.concatenate([x[1] for x in xs]), np.array([x[2] for x in xs]), 4.0, 2048)
    for r, c in zip(res, GOLD):
        check_against_gold(r, c)


@pytest.mark.parametrize("iterations", [64, 2048])
def test_many_pairs_against_oracle(geometry, iterations):
    rng = np.random.default_rng(5)
    pairs = []
    for k in range(60):
        n = int(rng.integers(0, 1200)); of = float(rng.uniform(0.0, 0.8))
        xI, xJ, _ = synth.two_view_matches(max(n, 1), of, seed=100 + k, wh=(1600, 1200))
        pairs.append((xI[:n], xJ[:n]))
    off = np.concatenate([[0], np.cumsum([len(p[0]) for p in pairs])]).astype(np.uint64)
    sz = np.tile(np.array([1600, 1200, 1600, 1200], np.int32), (len(pairs), 1))
    res = geometry.fundamental_acransac(off, np.concatenate([p[0] for p in pairs]), np.concatenate([p[1] for p in pairs]), sz, 4.0, iterations)
    kept = 0
    for (xI, xJ), r in zip(pairs, res):
        o = ck.oracle_acransac_fundamental(xI, xJ, (1600, 1200, 1600, 1200), 4.0, iterations) if len(xI) else dict(inliers=np.zeros(0, np.uint32), error_max=0.0, min_nfa=0.0)
        assert np.array_equal(r["inliers"], o["inliers"]), (len(xI), len(r["inliers"]), len(o["inliers"]))
        if len(o["inliers"]):
            kept += 1
            assert abs(r["error_max"] - o["error_max"]) <= 1e-9 * o["error_max"] and abs(r["min_nfa"] - o["min_nfa"]) <= 1e-9 * abs(o["min_nfa"])
            F = r["F"] / np.linalg.norm(r["F"]); G = o["F"] / np.linalg.norm(o["F"])
            assert min(np.abs(F - G).max(), np.abs(F + G).max()) <= 1e-7
    assert kept > 20


def test_rejects_what_it_does_not_implement(geometry):
    from openmvg_b200._lib import OmvgError
    xI, xJ, _ = synth.two_view_matches(50, 0.2, seed=1)
    with pytest.raises(OmvgError):
        geometry.fundamental_acransac([0, 50], xI, xJ, [[1000, 1000, 1000, 1000]], precision=float("inf"))


def test_homography_goldens_and_random_pairs(geometry):
    """OMVG_GEOM_HOMOGRAPHY: the reference goldens in one launch, then 40 random planar pairs against the oracle."""
    xs = [case_inputs_h(c) for c in GOLD_H]
    off = np.concatenate([[0], np.cumsum([len(x[0]) for x in xs])]).astype(np.uint64)
    res = geometry.homography_acransac(off, np.concatenate([x[0] for x in xs]), np.concatenate([x[1] for x in xs]), np.array([x[2] for x in xs]), 4.0, 2048)
    for r, c in zip(res, GOLD_H):
        check_against_gold(r, c)
    rng = np.random.default_rng(9)
    pairs = []
    for k in range(40):
        n = int(rng.integers(0, 1000)); of = float(rng.uniform(0.0, 0.8))
        xI, xJ, _ = synth.two_view_matches(max(n, 1), of, seed=300 + k, wh=(1600, 1200), planar=True)
        pairs.append((xI[:n], xJ[:n]))
    off = np.concatenate([[0], np.cumsum([len(p[0]) for p in pairs])]).astype(np.uint64)
    sz = np.tile(np.array([1600, 1200, 1600, 1200], np.int32), (len(pairs), 1))
    res = geometry.homography_acransac(off, np.concatenate([p[0] for p in pairs]), np.concatenate([p[1] for p in pairs]), sz, 4.0, 2048)
    kept = 0
    for (xI, xJ), r in zip(pairs, res):
        o = ck.oracle_acransac_homography(xI, xJ, (1600, 1200, 1600, 1200), 4.0, 2048) if len(xI) else dict(inliers=np.zeros(0, np.uint32))
        assert np.array_equal(r["inliers"], o["inliers"]), (len(xI), len(r["inliers"]), len(o["inliers"]))
        kept += len(o["inliers"]) > 10
    assert kept > 15
