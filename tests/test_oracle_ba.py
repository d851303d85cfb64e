"""CPU: the BA oracle against the committed golden outputs of the reference's own
Bundle_Adjustment_Ceres::Adjust (tests/golden/reference_outputs.json) and, when oracle/_ref is
present, the compiled reference itself."""
import json
import os

import numpy as np
import pytest

import checkers as ck
from openmvg_b200 import synth

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_outputs.json")))
SMALL = [c for c in GOLD["ba"] if c["scene"]["n_cams"] <= 100]


@pytest.mark.parametrize("case", SMALL, ids=lambda c: c["name"])
def test_golden_final_cost(case):
    s = synth.ba_scene(**case["scene"])
    o = ck.oracle_ba_solve(s, **case["opts"])
    assert o["usable"] == case["ok"]
    assert abs(o["initial_cost"] - case["initial_cost"]) <= 1e-12 * case["initial_cost"]
    # the golden cost is recomputed from the scene Adjust RETURNS (write-back rules included)
    returned = ck.oracle_ba_cost(s, o["poses"], o["intrinsics"], o["points"], use_loss=case["opts"].get("use_loss", 1))
    assert abs(returned - case["final_cost"]) <= 1e-9 * case["final_cost"]        # observed: ~1e-15
    if case["opts"].get("extrinsics_opt", 6) != 2:
        assert abs(o["final_cost"] - case["final_cost"]) <= 1e-9 * case["final_cost"]
    assert o["iterations"] == case["iterations"] and o["successful"] == case["successful"] and o["unsuccessful"] == case["unsuccessful"]
    assert ("Function tolerance" in case["termination"]) == (o["termination"] == 0)


def test_cost_matches_reference_formula():
    """1/2 sum rho(|r|^2), Huber a=16 (sfm_data_BA_ceres.cpp:249; loss_function.cc:47-61)."""
    s = synth.ba_scene(6, 100, 4, seed=1, outlier_frac=0.3)
    _, r, *_ = ck.oracle_ba_eval(s, use_loss=0)
    sq = (r ** 2).sum(1)
    want = 0.5 * np.where(sq <= 256.0, sq, 32.0 * np.sqrt(sq) - 256.0).sum()
    assert abs(ck.oracle_ba_cost(s) - want) <= 1e-12 * want
    assert (sq > 256).any() and (sq <= 256).any()


def test_jacobian_against_finite_differences():
    s = synth.ba_scene(4, 30, 4, seed=9, model=4)
    s["intrinsics"][:, 3:] = s["gt_dist"][3:]
    _, r0, Ji, Jc, Jp = ck.oracle_ba_eval(s, use_loss=0)
    h = 1e-6
    for name, J, arr, width in (("poses", Jc, "poses", 6), ("points", Jp, "points", 3), ("intrinsics", Ji, "intrinsics", 8)):
        for k in range(width):
            s2 = dict(s); a = s[arr].copy(); a[:, k] += h; s2[arr] = a
            _, r1, *_ = ck.oracle_ba_eval(s2, use_loss=0)
            fd = (r1 - r0) / h
            assert np.abs(fd - J[:, :, k]).max() <= 2e-4 * max(1.0, np.abs(J[:, :, k]).max()), (name, k)


@pytest.mark.skipif(not ck.have_ref_ba(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("kw", [dict(), dict(intrinsics_opt=1), dict(extrinsics_opt=4), dict(structure_opt=0)])
def test_against_compiled_reference(kw):
    s = synth.ba_scene(12, 400, 5, seed=13, outlier_frac=0.01)
    r = ck.ref_ba_adjust(s, threads=2, **kw)
    o = ck.oracle_ba_solve(s, **kw)
    assert r["ok"] and o["usable"]
    assert abs(o["final_cost"] - r["final_cost"]) <= 1e-9 * r["final_cost"]
    assert o["iterations"] == r["iterations"]
    assert np.abs(o["points"] - r["points"]).max() < 1e-7 and np.abs(o["poses"][:, 3:] - r["poses"][:, 3:]).max() < 1e-7
    from scipy.spatial.transform import Rotation   # angle-axis is ambiguous at theta ~ pi: compare the rotations
    Ro = Rotation.from_rotvec(o["poses"][:, :3]).as_matrix(); Rr = Rotation.from_rotvec(r["poses"][:, :3]).as_matrix()
    assert np.abs(Ro - Rr).max() < 1e-7


# ---- ground control points and pose-centre priors
@pytest.mark.parametrize("case", GOLD.get("ba_ext", []), ids=lambda c: c["name"])
def test_golden_gcp_and_priors(case):
    """Oracle vs the reference Adjust with control points / motion priors (golden = cost of the scene
    the reference returned, GCP and prior terms included)."""
    s = ck.golden_ext_scene(case)
    o = ck.oracle_ba_solve(s, **case["opts"])
    assert o["usable"] and case["ok"]
    assert abs(o["final_cost"] - case["final_cost"]) <= 1e-9 * case["final_cost"], (o["final_cost"], case["final_cost"])
    assert o["iterations"] == case["iterations"]
    if s.get("point_fixed") is not None:                      # GCP landmarks are constant
        f = s["point_fixed"].astype(bool)
        assert np.array_equal(o["points"][f], s["points"][f])
    ret = dict(s); ret.update(poses=o["poses"], intrinsics=o["intrinsics"], points=o["points"])
    assert abs(ck.oracle_ba_cost(ret, use_loss=case["opts"].get("use_loss", 1)) - o["final_cost"]) <= 1e-12 * o["final_cost"]


@pytest.mark.skipif(not ck.have_ref_ba(), reason="oracle/_ref not built")
def test_gcp_and_priors_against_compiled_reference():
    s = synth.add_priors(synth.add_gcp(synth.ba_scene(12, 300, 5, seed=3), 6, weight=15.0), sigma=0.02)   # GCPs live in the priors' frame
    r = ck.ref_ba_adjust_ex(s)
    t, fit, cen = ck.ref_ba_register_priors(s)
    assert fit == r["prior_fit"] and fit > 0
    o = ck.oracle_ba_solve(t)
    assert abs(o["final_cost"] - r["final_cost"]) <= 1e-9 * r["final_cost"]
    assert o["iterations"] == r["iterations"]
    # the reference undoes only the centring (sfm_data_BA_ceres.cpp:572): same for the oracle's result
    assert np.abs((o["points"] + cen) - r["points"]).max() <= 1e-7


def test_zero_weight_removes_observation_exactly():
    """Weight 0 must equal deleting the observation (the rejection loop relies on it)."""
    s = synth.ba_scene(10, 300, 5, seed=4, outlier_frac=0.05)
    rng = np.random.default_rng(0)
    keep = rng.random(len(s["obs_view"])) > 0.1
    w = dict(s); w["obs_weight"] = keep.astype(np.float64)
    d = dict(s); d["obs_view"] = np.ascontiguousarray(s["obs_view"][keep]); d["obs_point"] = np.ascontiguousarray(s["obs_point"][keep]); d["obs_xy"] = np.ascontiguousarray(s["obs_xy"][keep])
    a = ck.oracle_ba_solve(w); b = ck.oracle_ba_solve(d)
    assert abs(a["final_cost"] - b["final_cost"]) <= 1e-9 * b["final_cost"] and a["iterations"] == b["iterations"]   # (OpenMP atomics reorder sums)


@pytest.mark.skipif(not ck.have_ref_ba(), reason="oracle/_ref not built")
@pytest.mark.parametrize("kw", [dict(intrinsics_opt=1), dict(extrinsics_opt=4), dict(structure_opt=0), dict(use_loss=0), dict(model=3)],
                         ids=lambda k: "_".join(f"{a}{b}" for a, b in k.items()))
def test_gcp_priors_option_mixes_against_compiled_reference(kw):
    """Control points + motion priors under the option mixes of Optimize_Options and another camera model."""
    kw = dict(kw); model = kw.pop("model", 1)
    s = synth.add_priors(synth.add_gcp(synth.ba_scene(14, 400, 5, seed=6, model=model), 5, weight=12.0), sigma=0.015)
    ref_kw = {k: v for k, v in kw.items() if k in ("intrinsics_opt", "extrinsics_opt", "structure_opt", "use_loss")}
    r = ck.ref_ba_adjust_ex(s, **ref_kw)
    t, fit, cen = ck.ref_ba_register_priors(s)
    o = ck.oracle_ba_solve(t, **kw)
    assert r["ok"] and o["usable"]
    assert abs(o["final_cost"] - r["final_cost"]) <= 1e-8 * r["final_cost"], (o["final_cost"], r["final_cost"])
    assert o["iterations"] == r["iterations"]
