"""BA parity on the GPU: kernels and the full LM solve (through the C-ABI) against the oracle.
Tolerances: Jacobian / residual entries 1e-9 relative to the block scale (analytic vs dual-number
AD differ by rounding only); final cost 1e-6 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

import checkers as ck
from openmvg_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ba():
    from openmvg_b200 import ba as m
    return m


@pytest.mark.parametrize("model", [1, 2, 3, 4, 5, 7])
def test_eval_matches_oracle(ba, model):
    s = synth.ba_scene(12, 300, 5, seed=3, model=model, n_intrinsics=2, outlier_frac=0.05)
    s["intrinsics"][:, 3:] += s["gt_dist"][3:] * 0.9                    # non-trivial distortion parameters
    ctx = ba.BAContext(s)
    cost, r, Ji, Jc, Jp = ctx.debug_eval()
    oc, orr, oJi, oJc, oJp = ck.oracle_ba_eval(s)
    ctx.close()
    assert abs(cost - oc) <= 1e-12 * oc
    for a, b, name in ((r, orr, "r"), (Jp, oJp, "Jp"), (Jc, oJc, "Jc"), (Ji, oJi, "Ji")):
        scale = np.abs(b).max() + 1e-300
        assert np.abs(a - b).max() <= 1e-9 * scale, (name, np.abs(a - b).max(), scale)
    assert (np.abs(orr).max(axis=1) ** 2).max() > 256          # the Huber outlier branch was exercised


@pytest.mark.parametrize("cfg", [(10, 500, 4), (30, 1500, 8), (100, 5000, 10)])
def test_solve_matches_oracle(ba, cfg):
    s = synth.ba_scene(*cfg)
    g = ba.solve(s)
    o = ck.oracle_ba_solve(s)
    assert g["ok"] and o["usable"]
    assert abs(g["initial_cost"] - o["initial_cost"]) <= 1e-12 * o["initial_cost"]
    assert abs(g["final_cost"] - o["final_cost"]) <= 1e-6 * o["final_cost"], (g["final_cost"], o["final_cost"])
    assert g["iterations"] == o["iterations"] and g["successful_steps"] == o["successful"]
    assert g["termination"] == o["termination"]
    # the returned parameters reproduce the reported cost
    c = ck.oracle_ba_cost(s, g["poses"], g["intrinsics"], g["points"])
    assert abs(c - g["final_cost"]) <= 1e-10 * c


@pytest.mark.parametrize("opts", [
    dict(intrinsics_opt=1, extrinsics_opt=6, structure_opt=1),      # intrinsics NONE
    dict(intrinsics_opt=14, extrinsics_opt=4, structure_opt=1),     # ADJUST_TRANSLATION only (global SfM pass 1)
    dict(intrinsics_opt=2, extrinsics_opt=2, structure_opt=1),      # focal only, rotation only
    dict(intrinsics_opt=14, extrinsics_opt=6, structure_opt=0),     # structure fixed
    dict(intrinsics_opt=1, extrinsics_opt=1, structure_opt=1),      # only points move
    dict(intrinsics_opt=14, extrinsics_opt=6, structure_opt=1, use_loss=0),
])
def test_option_mix_matches_oracle(ba, opts):
    s = synth.ba_scene(16, 800, 6, seed=5, outlier_frac=0.02)
    g = ba.solve(s, **opts)
    o = ck.oracle_ba_solve(s, **opts)
    assert g["ok"] and o["usable"]
    assert abs(g["final_cost"] - o["final_cost"]) <= 1e-6 * o["final_cost"], (g["final_cost"], o["final_cost"])
    assert g["iterations"] == o["iterations"]


@pytest.mark.parametrize("model", [2, 3, 4, 5, 7])
def test_distortion_models_solve(ba, model):
    s = synth.ba_scene(14, 700, 6, seed=8, model=model)
    g = ba.solve(s)
    o = ck.oracle_ba_solve(s)
    assert g["ok"] and o["usable"]
    assert abs(g["final_cost"] - o["final_cost"]) <= 1e-6 * o["final_cost"], (g["final_cost"], o["final_cost"])


@pytest.mark.parametrize("n_intr,model", [(3, 1), (2, 2)])
def test_multiple_intrinsic_groups(ba, n_intr, model):
    """Points seen through several intrinsic groups: the general (per-observation-pair) border path."""
    s = synth.ba_scene(18, 900, 6, seed=6, model=model, n_intrinsics=n_intr)
    g = ba.solve(s)
    o = ck.oracle_ba_solve(s)
    assert g["ok"] and o["usable"]
    assert abs(g["final_cost"] - o["final_cost"]) <= 1e-6 * o["final_cost"], (g["final_cost"], o["final_cost"])
    assert g["iterations"] == o["iterations"]


def test_pcg_variants_agree(ba, monkeypatch):
    """Two-level block-PCG (default) and plain block-Jacobi PCG with the border inside the iteration."""
    monkeypatch.setenv("OMVG_BA_DENSE_MAX", "0"); monkeypatch.setenv("OMVG_BA_DENSE2_MAX", "0")     # (40 cameras would be solved directly)
    s = synth.ba_scene(40, 2000, 8, seed=12)
    a = ba.solve(s)
    monkeypatch.setenv("OMVG_BA_PCG1", "1")
    b = ba.solve(s)
    assert abs(a["final_cost"] - b["final_cost"]) <= 1e-7 * a["final_cost"]
    assert a["iterations"] == b["iterations"] and a["pcg_iterations"] < b["pcg_iterations"]


def test_context_reset_is_reproducible(ba):
    s = synth.ba_scene(20, 1000, 6, seed=2)
    ctx = ba.BAContext(s)
    a = ctx.run(); ctx.reset(); b = ctx.run()
    # cost/gradient/model reductions are fixed-order; the Schur accumulation uses FP64 atomics, so two
    # runs agree to rounding (not bitwise)
    assert abs(a["final_cost"] - b["final_cost"]) <= 1e-12 * a["final_cost"] and a["iterations"] == b["iterations"]
    ctx.close()


def test_config2_against_reference_golden(ba):
    """BASELINE.json configs[1] (1000 cams / 100k pts / 1M obs) at full size: final cost within 1e-6 of the
    reference's own Bundle_Adjustment_Ceres::Adjust (golden generated by tests/golden/make_golden.py), same
    Ceres iteration count; the returned scene reproduces the reported cost (size-independent property)."""
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_outputs.json")))
    case = [c for c in gold["ba"] if c["name"] == "config2_1000_100k_1M"][0]
    s = synth.ba_scene(**case["scene"])
    g = ba.solve(s)
    assert g["ok"]
    assert abs(g["initial_cost"] - case["initial_cost"]) <= 1e-12 * case["initial_cost"]
    assert abs(g["final_cost"] - case["final_cost"]) <= 1e-6 * case["final_cost"], (g["final_cost"], case["final_cost"])
    assert g["iterations"] == case["iterations"]
    c = ck.oracle_ba_cost(s, g["poses"], g["intrinsics"], g["points"])
    assert abs(c - g["final_cost"]) <= 1e-10 * c


@pytest.mark.parametrize("name", ["c1", "c30", "c100", "outliers", "intr_none", "translation_only", "focal_rotation",
                                  "structure_fixed", "points_only", "no_loss", "radial1", "radial3", "brown", "fisheye",
                                  "spherical", "spherical_rotation_only", "intr40", "intr200_radial3", "intr5_brown"])
def test_reference_golden_cases(ba, name):
    """Every golden case of the reference (option mixes, camera models): the cost of the RETURNED scene within 1e-6."""
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_outputs.json")))
    case = [c for c in gold["ba"] if c["name"] == name][0]
    s = synth.ba_scene(**case["scene"])
    g = ba.solve(s, **case["opts"])
    assert g["ok"] == case["ok"]
    returned = ck.oracle_ba_cost(s, g["poses"], g["intrinsics"], g["points"], use_loss=case["opts"].get("use_loss", 1))
    assert abs(returned - case["final_cost"]) <= 1e-6 * case["final_cost"], (returned, case["final_cost"])
    assert g["iterations"] == case["iterations"]


def test_spherical_and_pinhole_views_in_one_scene(ba):
    """CAMERA_SPHERICAL (functor.hpp:662-760: no intrinsic block) next to a pinhole group in the same problem."""
    s = synth.ba_scene(12, 500, 6, seed=4, n_intrinsics=2)
    sp = synth.ba_scene(12, 500, 6, seed=4, n_intrinsics=2, model=7)
    s["intr_model"] = np.array([1, 7], np.int32); s["intrinsics"][1] = sp["intrinsics"][1]
    sphere_obs = s["view_intr"][s["obs_view"]] == 1
    s["obs_xy"][sphere_obs] = sp["obs_xy"][sphere_obs]
    g = ba.solve(s); o = ck.oracle_ba_solve(s)
    assert g["ok"] and o["usable"] and sphere_obs.any() and (~sphere_obs).any()
    assert abs(g["final_cost"] - o["final_cost"]) <= 1e-6 * o["final_cost"] and g["iterations"] == o["iterations"]
    assert np.array_equal(g["intrinsics"][1], s["intrinsics"][1])          # {w, h} are constants, never written back


def test_many_intrinsic_groups_and_many_free_columns(ba):
    """One intrinsic group per view (80 groups, 240 free columns) and 6 Brown groups (48 free columns > 32): the
    whole-system PCG with the aggregated coarse space on the camera rows; same result and iteration count as the oracle."""
    for kw in (dict(n_cams=80, n_points=3000, obs_per_point=6, seed=3, n_intrinsics=80), dict(n_cams=30, n_points=1500, obs_per_point=8, seed=3, n_intrinsics=6, model=4)):
        s = synth.ba_scene(**kw)
        g = ba.solve(s); o = ck.oracle_ba_solve(s)
        assert g["ok"] and o["usable"]
        assert abs(g["final_cost"] - o["final_cost"]) <= 1e-6 * o["final_cost"], (kw, g["final_cost"], o["final_cost"])
        assert g["iterations"] == o["iterations"]
        assert g["pcg_iterations"] < 1500 * g["lm_steps"], g["pcg_iterations"]      # (measured ~700 per LM step: functional, not fast)


def test_concurrent_adjust_from_four_host_threads(ba):
    """Some engines call Adjust from inside an OpenMP region (relative_pose_engine.cpp:84,189; stellar_solver.cpp:471
    under sfm_stellar_engine.cpp:379): four host threads solve different small scenes at once through the C ABI
    (one context, one stream each); results equal the sequential ones."""
    import threading
    scenes = [synth.ba_scene(6 + 2 * k, 300 + 50 * k, 4, seed=20 + k) for k in range(8)]
    seq = [ba.solve(s) for s in scenes]
    out = [None] * len(scenes); err = []

    def work(lo):
        try:
            for rep in range(3):
                for k in range(lo, len(scenes), 4):
                    out[k] = ba.solve(scenes[k])
        except Exception as e:                                   # noqa: BLE001
            err.append(e)
    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in th]; [t.join() for t in th]
    assert not err, err
    for a, b in zip(seq, out):
        assert b["ok"] and a["iterations"] == b["iterations"]
        assert abs(a["final_cost"] - b["final_cost"]) <= 1e-11 * a["final_cost"]


def test_residual_norms_for_outlier_rejection(ba):
    """omvg_ba_residual_norms == |residual| of the reference's IntrinsicBase::residual at the refined scene
    (what RemoveOutliers_PixelResidualError thresholds, sfm_data_filters.cpp:40-73)."""
    s = synth.ba_scene(12, 600, 5, seed=4, outlier_frac=0.05)
    ctx = ba.BAContext(s)
    ctx.run()
    norms = ctx.residual_norms()
    poses, intr, pts = ctx.download(); ctx.close()
    s2 = dict(s); s2["poses"], s2["intrinsics"], s2["points"] = poses, intr, pts
    _, r, *_ = ck.oracle_ba_eval(s2, use_loss=0)
    want = np.sqrt((r ** 2).sum(1))
    assert np.abs(norms - want).max() <= 1e-9 * max(1.0, want.max())
    assert (want > 4.0).any() and (want < 4.0).any()


# ---- ground control points, pose-centre priors, resident rejection loop
def _ext_cases():
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_outputs.json")))
    return gold.get("ba_ext", [])


@pytest.mark.parametrize("case", _ext_cases(), ids=lambda c: c["name"])
def test_gcp_and_priors_against_reference_golden(ba, case):
    """Control points (fixed landmarks, weighted residuals, no loss) and pose-centre priors (Huber(fit^2))
    through the C-ABI: cost of the returned scene within 1e-6 of the reference Adjust's, same iterations,
    and the oracle agrees."""
    s = ck.golden_ext_scene(case)
    g = ba.solve(s, **case["opts"])
    assert g["ok"] and case["ok"]
    ret = dict(s); ret.update(poses=g["poses"], intrinsics=g["intrinsics"], points=g["points"])
    returned = ck.oracle_ba_cost(ret, use_loss=case["opts"].get("use_loss", 1))
    assert abs(returned - case["final_cost"]) <= 1e-6 * case["final_cost"], (returned, case["final_cost"])
    assert abs(g["final_cost"] - returned) <= 1e-10 * returned
    assert g["iterations"] == case["iterations"]
    o = ck.oracle_ba_solve(s, **case["opts"])
    assert abs(g["initial_cost"] - o["initial_cost"]) <= 1e-12 * o["initial_cost"]
    assert abs(g["final_cost"] - o["final_cost"]) <= 1e-7 * o["final_cost"]
    if s.get("point_fixed") is not None:
        f = s["point_fixed"].astype(bool)
        assert np.array_equal(g["points"][f], s["points"][f])       # GCP landmarks untouched


def test_weighted_eval_matches_oracle(ba):
    """Residual / Jacobian rows of weighted, loss-free observations and fixed landmarks."""
    s = synth.add_gcp(synth.ba_scene(12, 300, 5, seed=3, model=3, outlier_frac=0.05), 5, weight=7.0)
    ctx = ba.BAContext(s)
    cost, r, Ji, Jc, Jp = ctx.debug_eval()
    ctx.close()
    # the oracle's eval entry has no extension arguments: weight rows by hand, check flags through the cost
    plain = {k: v for k, v in s.items() if k not in ("obs_weight", "obs_no_loss", "point_fixed")}
    oc, orr, oJi, oJc, oJp = ck.oracle_ba_eval(plain, use_loss=0)
    g = s["obs_no_loss"].astype(bool)
    w = s["obs_weight"][g][:, None]
    assert np.abs(r[g] - w * orr[g]).max() <= 1e-9 * np.abs(w * orr[g]).max()
    assert np.abs(Jc[g] - w[:, :, None] * oJc[g]).max() <= 1e-9 * np.abs(w[:, :, None] * oJc[g]).max()
    assert np.abs(Jp[g]).max() == 0.0                           # fixed landmark: no point columns
    assert abs(cost - ck.oracle_ba_cost(s)) <= 1e-12 * cost


def test_rejection_loop_on_resident_scene(ba):
    """BA -> residual norms -> weight-0 rejection -> BA on ONE device context equals solving the scene with
    those observations deleted (sequential_SfM.cpp:1226-1243 loop without rebuilding the problem)."""
    s = synth.ba_scene(20, 1500, 6, seed=6, outlier_frac=0.05)
    ctx = ba.BAContext(s)
    a = ctx.run()
    norms = ctx.residual_norms()
    keep = norms <= 4.0                                          # RemoveOutliers_PixelResidualError(4.0)
    assert 0 < (~keep).sum() < 0.2 * len(keep)
    p1, i1, x1 = ctx.download()
    ctx.commit(); ctx.set_obs_weights(keep.astype(np.float64)); ctx.reset()
    b = ctx.run()
    p2, i2, x2 = ctx.download(); ctx.close()
    d = dict(s); d.update(poses=p1, intrinsics=i1, points=x1,
                          obs_view=np.ascontiguousarray(s["obs_view"][keep]), obs_point=np.ascontiguousarray(s["obs_point"][keep]),
                          obs_xy=np.ascontiguousarray(s["obs_xy"][keep]))
    o = ck.oracle_ba_solve(d)
    assert b["ok"] and o["usable"]
    assert abs(b["initial_cost"] - o["initial_cost"]) <= 1e-12 * o["initial_cost"]
    assert abs(b["final_cost"] - o["final_cost"]) <= 1e-7 * o["final_cost"], (b["final_cost"], o["final_cost"])
    assert b["iterations"] == o["iterations"] and b["final_cost"] < 0.5 * a["final_cost"]


def test_device_side_rejection_equals_the_reference_rule(ba):
    """omvg_ba_reject_outliers = RemoveOutliers_PixelResidualError(4.0, 2) (sfm_data_filters.cpp:40-73) evaluated on
    the device: same removed observations as thresholding the residual norms on the host, tracks left with < 2
    observations removed whole (not counted as outliers), second call removes nothing; then BA on the reduced
    resident scene equals the oracle on the scene with those observations and tracks deleted."""
    s = synth.ba_scene(20, 1500, 3, seed=6, outlier_frac=0.12)            # short tracks: some fall below 2 observations
    ctx = ba.BAContext(s)
    a = ctx.run()
    norms = ctx.residual_norms()
    removed, pt_removed, n_out, n_tracks = ctx.reject_outliers(4.0, 2)
    out = norms > 4.0
    assert n_out == int(out.sum()) > 0
    alive = np.bincount(s["obs_point"][~out], minlength=len(s["points"]))
    want_pts = alive < 2
    assert n_tracks == int(want_pts.sum()) > 0 and np.array_equal(pt_removed, want_pts)
    want_removed = out | want_pts[s["obs_point"]]
    assert np.array_equal(removed, want_removed)
    r2, p2, n2, t2 = ctx.reject_outliers(4.0, 2)                           # idempotent at the same parameters
    assert n2 == 0 and t2 == 0 and not r2.any() and not p2.any()
    p1, i1, x1 = ctx.download()
    b = ctx.run()
    ctx.close()
    keep = ~want_removed
    remap = np.cumsum(~want_pts) - 1
    d = dict(s); d.update(poses=p1, intrinsics=i1, points=np.ascontiguousarray(x1[~want_pts]),
                          obs_view=np.ascontiguousarray(s["obs_view"][keep]), obs_point=np.ascontiguousarray(remap[s["obs_point"][keep]].astype(np.int32)),
                          obs_xy=np.ascontiguousarray(s["obs_xy"][keep]))
    o = ck.oracle_ba_solve(d)
    assert b["ok"] and o["usable"]
    assert abs(b["initial_cost"] - o["initial_cost"]) <= 1e-12 * o["initial_cost"]
    assert abs(b["final_cost"] - o["final_cost"]) <= 1e-7 * o["final_cost"], (b["final_cost"], o["final_cost"])
    assert b["iterations"] == o["iterations"] and b["final_cost"] < a["final_cost"]


def test_remove_points_on_resident_scene(ba):
    """omvg_ba_remove_points (the host-side angle test's verdict): the named tracks vanish from the problem."""
    s = synth.ba_scene(12, 600, 5, seed=4)
    ctx = ba.BAContext(s)
    mask = np.zeros(len(s["points"]), np.uint8); mask[::7] = 1
    removed, nt = ctx.remove_points(mask)
    assert nt == int(mask.sum()) and np.array_equal(removed, mask[s["obs_point"]].astype(bool))
    b = ctx.run(); ctx.close()
    keep = ~removed; remap = np.cumsum(mask == 0) - 1
    d = dict(s); d.update(points=np.ascontiguousarray(s["points"][mask == 0]), obs_view=np.ascontiguousarray(s["obs_view"][keep]),
                          obs_point=np.ascontiguousarray(remap[s["obs_point"][keep]].astype(np.int32)), obs_xy=np.ascontiguousarray(s["obs_xy"][keep]))
    o = ck.oracle_ba_solve(d)
    assert abs(b["final_cost"] - o["final_cost"]) <= 1e-7 * o["final_cost"] and b["iterations"] == o["iterations"]


def test_direct_and_iterative_reduced_solves_agree(ba, monkeypatch):
    """Up to 220 reduced unknowns one CTA solves the reduced camera system directly (dense L D L', as Ceres' Cholesky does),
    up to 640 it is inverted explicitly (blocked Gauss-Jordan over all SMs); above, and with OMVG_BA_DENSE_MAX=0 /
    OMVG_BA_DENSE2_MAX=0, the two-level PCG runs.  Same LM trajectory either way."""
    for cams, model in ((12, 1), (30, 3), (35, 1), (50, 1), (80, 2)):   # 6 * 35 + 8 = 218 unknowns <= 220: one CTA; up to 640: explicit inverse
        s = synth.ba_scene(cams, 60 * cams, 8, seed=5 + cams, model=model)
        monkeypatch.delenv("OMVG_BA_DENSE_MAX", raising=False); monkeypatch.delenv("OMVG_BA_DENSE2_MAX", raising=False)
        a = ba.solve(s)
        monkeypatch.setenv("OMVG_BA_DENSE_MAX", "0"); monkeypatch.setenv("OMVG_BA_DENSE2_MAX", "0")
        b = ba.solve(s)
        assert a["pcg_iterations"] == 0 and b["pcg_iterations"] > 0, (cams, a["pcg_iterations"], b["pcg_iterations"])
        assert a["iterations"] == b["iterations"]
        assert abs(a["final_cost"] - b["final_cost"]) <= 1e-8 * a["final_cost"]
        o = ck.oracle_ba_solve(s)
        assert abs(a["final_cost"] - o["final_cost"]) <= 1e-9 * o["final_cost"] and a["iterations"] == o["iterations"], (a["final_cost"], o["final_cost"])


def test_tight_pcg_reproduces_oracle_to_rounding(ba, monkeypatch):
    """The only approximation in the iterative path is the inner PCG tolerance (default 1e-8, final cost within
    ~3e-9 of the exact-solve reference over all golden cases; north_star allows 1e-6).  Tightened to 1e-12
    the result agrees with the exact (dense Cholesky) oracle to rounding."""
    monkeypatch.setenv("OMVG_BA_DENSE_MAX", "0"); monkeypatch.setenv("OMVG_BA_DENSE2_MAX", "0")
    s = synth.ba_scene(30, 1500, 8)
    g = ba.solve(s, pcg_tolerance=1e-12)
    o = ck.oracle_ba_solve(s)
    assert abs(g["final_cost"] - o["final_cost"]) <= 1e-11 * o["final_cost"], (g["final_cost"], o["final_cost"])
    assert g["iterations"] == o["iterations"]


@pytest.mark.parametrize("cfg", [(40, 120, 36), (64, 50, 64)])
def test_many_observations_per_landmark(ba, cfg):
    """Landmarks with more than 32 observations take the per-observation Schur kernel (the warp-per-landmark
    kernel stages at most 32): both paths in one solve."""
    s = synth.ba_scene(*cfg, seed=2)
    extra = synth.ba_scene(cfg[0], 400, 5, seed=3)             # plus ordinary short tracks
    n0 = len(s["points"])
    s["points"] = np.ascontiguousarray(np.concatenate([s["points"], extra["points"]]))
    s["obs_view"] = np.ascontiguousarray(np.concatenate([s["obs_view"], extra["obs_view"]]))
    s["obs_point"] = np.ascontiguousarray(np.concatenate([s["obs_point"], extra["obs_point"] + n0]).astype(np.int32))
    s["obs_xy"] = np.ascontiguousarray(np.concatenate([s["obs_xy"], extra["obs_xy"]]))
    g = ba.solve(s)
    o = ck.oracle_ba_solve(s)
    assert g["ok"] and o["usable"]
    assert abs(g["final_cost"] - o["final_cost"]) <= 1e-6 * o["final_cost"], (g["final_cost"], o["final_cost"])
    assert g["iterations"] == o["iterations"]


@pytest.mark.parametrize("kw", [dict(intrinsics_opt=1), dict(extrinsics_opt=4), dict(extrinsics_opt=2), dict(structure_opt=0), dict(use_loss=0), dict(model=3),
                                dict(model=5, n_intrinsics=2)], ids=lambda k: "_".join(f"{a}{b}" for a, b in k.items()))
def test_gcp_priors_option_mixes_match_oracle(ba, kw):
    """Control points + pose-centre priors under the option mixes of Optimize_Options, other camera models and
    several intrinsic groups: same flat problem to the GPU path and to the oracle."""
    kw = dict(kw); model = kw.pop("model", 1); ni = kw.pop("n_intrinsics", 1)
    s = synth.add_priors(synth.add_gcp(synth.ba_scene(14, 400, 5, seed=6, model=model, n_intrinsics=ni), 5, weight=12.0), sigma=0.015)
    s["prior_huber_a"] = 0.03 ** 2
    g = ba.solve(s, **kw)
    o = ck.oracle_ba_solve(s, **kw)
    assert g["ok"] and o["usable"]
    assert abs(g["initial_cost"] - o["initial_cost"]) <= 1e-12 * o["initial_cost"]
    assert abs(g["final_cost"] - o["final_cost"]) <= 1e-6 * o["final_cost"], (g["final_cost"], o["final_cost"])
    assert g["iterations"] == o["iterations"]
    ret = dict(s); ret.update(poses=g["poses"], intrinsics=g["intrinsics"], points=g["points"])
    assert abs(ck.oracle_ba_cost(ret, use_loss=kw.get("use_loss", 1)) - g["final_cost"]) <= 1e-9 * g["final_cost"] or kw.get("extrinsics_opt") == 2
