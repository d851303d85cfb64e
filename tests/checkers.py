"""ctypes bindings for the CHECKERS: oracle/_build/liboracle.so (our CPU restatement) and
oracle/_ref/libref_{match,ba}.so (the reference itself, compiled by oracle/Makefile).

Test infrastructure only — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg; never by the product package.
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
REF_MATCH_SO = os.path.join(ROOT, "oracle", "_ref", "libref_match.so")
REF_BA_SO = os.path.join(ROOT, "oracle", "_ref", "libref_ba.so")
REF_GEOM_SO = os.path.join(ROOT, "oracle", "_ref", "libref_geom.so")

_P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731

# Ceres defaults as overridden by sfm_data_BA_ceres.cpp:477-493 (solver.h:62-138)
BA_DEFAULT_OPTS = dict(
    intrinsics_opt=14, extrinsics_opt=6, structure_opt=1, use_loss=1, huber_a=16.0,
    max_num_iterations=50, function_tolerance=1e-6, gradient_tolerance=1e-10,
    parameter_tolerance=1e-8, initial_radius=1e4, max_radius=1e16, min_radius=1e-32,
    min_relative_decrease=1e-3, min_lm_diagonal=1e-6, max_lm_diagonal=1e32,
    max_consecutive_invalid_steps=5)
BA_OPT_ORDER = list(BA_DEFAULT_OPTS.keys())


def build_oracle():
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return ORACLE_SO


_cache = {}


def oracle():
    if "o" not in _cache:
        lib = ctypes.CDLL(build_oracle())
        lib.oracle_match_pair.restype = ctypes.c_int64
        lib.oracle_fnv1a_ij.restype = ctypes.c_uint64
        lib.oracle_ba_cost.restype = ctypes.c_double
        lib.oracle_ba_eval.restype = ctypes.c_double
        lib.oracle_ba_cost_ex.restype = ctypes.c_double
        _cache["o"] = lib
    return _cache["o"]


def have_ref_match():
    return os.path.exists(REF_MATCH_SO)


def have_ref_ba():
    return os.path.exists(REF_BA_SO)


def ref_match():
    if "rm" not in _cache:
        lib = ctypes.CDLL(REF_MATCH_SO)
        lib.ref_match_pair.restype = ctypes.c_int64
        lib.ref_match_collection.restype = ctypes.c_int64
        _cache["rm"] = lib
    return _cache["rm"]


def ref_ba():
    if "rb" not in _cache:
        _cache["rb"] = ctypes.CDLL(REF_BA_SO)
    return _cache["rb"]


# ----------------------------------------------------------------------------- MATCH
def oracle_match_pair(di, dj, ratio=0.8):
    ni, nj = len(di), len(dj)
    out = np.zeros(2 * max(nj, 1), np.uint32)
    scratch = np.zeros(3 * max(nj, 1), np.int32)
    n = oracle().oracle_match_pair(_P(di), ni, _P(dj), nj, ctypes.c_float(ratio), _P(out), _P(scratch))
    return out[:2 * n].reshape(-1, 2).copy()


def oracle_top2(di, dj):
    nj = len(dj)
    d1 = np.zeros(nj, np.int32); i1 = np.zeros(nj, np.uint32); d2 = np.zeros(nj, np.int32)
    rc = oracle().oracle_top2(_P(di), len(di), _P(dj), nj, _P(d1), _P(i1), _P(d2))
    return rc, d1, i1, d2


def ref_match_pair(di, dj, ratio=0.8):
    out = np.zeros(2 * max(len(dj), 1), np.uint32)
    n = ref_match().ref_match_pair(_P(di), len(di), _P(dj), len(dj), ctypes.c_float(ratio), _P(out))
    return out[:2 * n].reshape(-1, 2).copy()


def ref_match_collection(descs, pi, pj, ratio=0.8):
    counts = np.array([len(d) for d in descs], np.uint32)
    row_start = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.uint64)
    allrows = np.ascontiguousarray(np.concatenate([d.reshape(-1, 128) for d in descs]) if len(descs) else np.zeros((0, 128), np.uint8))
    pi = np.ascontiguousarray(pi, np.uint32); pj = np.ascontiguousarray(pj, np.uint32)
    offsets = np.zeros(len(pi) + 1, np.uint64)
    cap = int(sum(int(counts[j]) for j in pj)) + 1
    ij = np.zeros(2 * cap, np.uint32)
    n = ref_match().ref_match_collection(_P(allrows), _P(row_start), _P(counts), len(descs), _P(pi), _P(pj),
                                         ctypes.c_uint64(len(pi)), ctypes.c_float(ratio), _P(offsets), _P(ij),
                                         ctypes.c_uint64(cap))
    assert n >= 0
    return offsets, ij[:2 * n].reshape(-1, 2).copy()


def oracle_match_collection(descs, pi, pj, ratio=0.8):
    """Matcher_Regions.cpp:32-107 over the oracle's pair routine; CSR in the order of (pi,pj)."""
    offsets = [0]; rows = []
    for i, j in zip(pi, pj):
        m = oracle_match_pair(descs[int(i)], descs[int(j)], ratio)
        rows.append(m); offsets.append(offsets[-1] + len(m))
    ij = np.concatenate(rows) if rows else np.zeros((0, 2), np.uint32)
    return np.array(offsets, np.uint64), ij.astype(np.uint32)


# ----------------------------------------------------------------------------- BA
def _ba_args(s, poses, intr, pts):
    return (len(poses), _P(poses), len(intr), _P(intr), _P(s["intr_model"]), len(pts), _P(pts),
            len(s["view_pose"]), _P(s["view_pose"]), _P(s["view_intr"]), ctypes.c_long(len(s["obs_view"])),
            _P(s["obs_view"]), _P(s["obs_point"]), _P(s["obs_xy"]))


def _opt(s, key, dtype):
    a = s.get(key)
    return None if a is None else np.ascontiguousarray(a, dtype)


def has_ext(s):
    return any(s.get(k) is not None for k in ("obs_weight", "obs_no_loss", "point_fixed", "prior_pose"))


def _ba_ext_args(s, with_fixed=True):
    """Extension arrays of a scene (GCP weights / flags / fixed landmarks, pose-centre priors)."""
    keep = [_opt(s, "obs_weight", np.float64), _opt(s, "obs_no_loss", np.uint8), _opt(s, "point_fixed", np.uint8),
            _opt(s, "prior_pose", np.int32), _opt(s, "prior_center", np.float64), _opt(s, "prior_weight", np.float64)]
    p = lambda a: None if a is None else _P(a)  # noqa: E731
    npri = 0 if keep[3] is None else len(keep[3])
    head = [p(keep[0]), p(keep[1])] + ([p(keep[2])] if with_fixed else [])
    return keep, head + [npri, p(keep[3]), p(keep[4]), p(keep[5]), ctypes.c_double(float(s.get("prior_huber_a", 0.0)))]


def oracle_ba_cost(s, poses=None, intr=None, pts=None, use_loss=1, huber_a=16.0):
    poses = s["poses"] if poses is None else poses
    intr = s["intrinsics"] if intr is None else intr
    pts = s["points"] if pts is None else pts
    if has_ext(s):
        keep, ext = _ba_ext_args(s, with_fixed=False)
        return oracle().oracle_ba_cost_ex(*_ba_args(s, poses, intr, pts), *ext, int(use_loss), ctypes.c_double(huber_a))
    return oracle().oracle_ba_cost(*_ba_args(s, poses, intr, pts), int(use_loss), ctypes.c_double(huber_a))


def oracle_ba_eval(s, use_loss=1, huber_a=16.0):
    n = len(s["obs_view"])
    r = np.zeros((n, 2)); Ji = np.zeros((n, 2, 8)); Jc = np.zeros((n, 2, 6)); Jp = np.zeros((n, 2, 3))
    c = oracle().oracle_ba_eval(*_ba_args(s, s["poses"], s["intrinsics"], s["points"]), int(use_loss),
                                ctypes.c_double(huber_a), _P(r), _P(Ji), _P(Jc), _P(Jp))
    return c, r, Ji, Jc, Jp


def oracle_ba_solve(s, **kw):
    o = dict(BA_DEFAULT_OPTS); o.update(kw)
    opts = np.array([float(o[k]) for k in BA_OPT_ORDER])
    poses = s["poses"].copy(); intr = s["intrinsics"].copy(); pts = s["points"].copy()
    summ = np.zeros(16); trace = np.zeros((128, 4))
    if has_ext(s):
        keep, ext = _ba_ext_args(s)
        rc = oracle().oracle_ba_solve_ex(*_ba_args(s, poses, intr, pts), *ext, _P(opts), _P(summ), _P(trace), 128)
    else:
        rc = oracle().oracle_ba_solve(*_ba_args(s, poses, intr, pts), _P(opts), _P(summ), _P(trace), 128)
    return dict(rc=rc, poses=poses, intrinsics=intr, points=pts, initial_cost=summ[0], final_cost=summ[1],
                iterations=int(summ[2]), successful=int(summ[3]), unsuccessful=int(summ[4]),
                termination=int(summ[5]), usable=bool(summ[6]), trace=trace[:int(summ[7])], lm_steps=int(summ[8]))


def ref_ba_adjust(s, intrinsics_opt=14, extrinsics_opt=6, structure_opt=1, threads=0, use_loss=1):
    opts = np.array([intrinsics_opt, extrinsics_opt, structure_opt, threads, use_loss], np.int32)
    poses = s["poses"].copy(); intr = s["intrinsics"].copy(); pts = s["points"].copy()
    out = np.zeros(4); rep = ctypes.create_string_buffer(1 << 16)
    rc = ref_ba().ref_ba_adjust(*_ba_args(s, poses, intr, pts), _P(opts), _P(out), rep, 1 << 16)
    text = rep.value.decode(errors="replace")

    def grab(pat, cast=float, default=None):
        m = re.search(pat, text)
        return cast(m.group(1)) if m else default
    return dict(rc=rc, ok=bool(out[0]), initial_cost=out[1], final_cost=out[2], wall_s=out[3], poses=poses,
                intrinsics=intr, points=pts, report=text,
                iterations=grab(r"Minimizer iterations\s+(\d+)", int),
                successful=grab(r"Successful steps\s+(\d+)", int),
                unsuccessful=grab(r"Unsuccessful steps\s+(\d+)", int),
                minimizer_s=grab(r"\nMinimizer\s+([0-9.]+)"),
                preprocessor_s=grab(r"Preprocessor\s+([0-9.]+)"),
                linear_solver_s=grab(r"Linear solver\s+([0-9.]+)"),
                jacobian_s=grab(r"Jacobian evaluation\s+([0-9.]+)"),
                residual_s=grab(r"Residual evaluation\s+([0-9.]+)"),
                threads=grab(r"\nThreads\s+\d+\s+(\d+)", int),
                termination=grab(r"Termination:\s+(.*)", str))


def _split_gcp(s):
    """Flat scene with fixed landmarks -> (regular scene view, GCP arrays) for the reference driver:
    fixed landmarks become SfM_Data::control_points, their observations the GCP observations."""
    fixed = np.asarray(s.get("point_fixed", np.zeros(len(s["points"]), np.uint8))).astype(bool)
    reg_idx = np.flatnonzero(~fixed); gcp_idx = np.flatnonzero(fixed)
    remap = np.full(len(fixed), -1, np.int64); remap[reg_idx] = np.arange(len(reg_idx))
    gmap = np.full(len(fixed), -1, np.int64); gmap[gcp_idx] = np.arange(len(gcp_idx))
    og = fixed[s["obs_point"]]
    reg = dict(s)
    reg["points"] = np.ascontiguousarray(s["points"][reg_idx])
    reg["obs_view"] = np.ascontiguousarray(s["obs_view"][~og]); reg["obs_xy"] = np.ascontiguousarray(s["obs_xy"][~og])
    reg["obs_point"] = np.ascontiguousarray(remap[s["obs_point"][~og]].astype(np.int32))
    w = 20.0
    if og.any():
        ws = np.asarray(s["obs_weight"])[og]
        assert np.all(ws == ws[0]) and np.all(np.asarray(s["obs_no_loss"])[og] == 1), "GCP observations: one weight, no loss"
        w = float(ws[0])
    if "obs_weight" in s:
        assert np.all(np.asarray(s["obs_weight"])[~og] == 1.0) and np.all(np.asarray(s["obs_no_loss"])[~og] == 0)
    gcp = dict(X=np.ascontiguousarray(s["points"][gcp_idx]), obs_view=np.ascontiguousarray(s["obs_view"][og]),
               obs_gcp=np.ascontiguousarray(gmap[s["obs_point"][og]].astype(np.int32)),
               obs_xy=np.ascontiguousarray(s["obs_xy"][og]), weight=w)
    return reg, gcp, reg_idx, gcp_idx


def _prior_views(s):
    """The reference hangs a prior on a VIEW; scenes here use view v <-> pose v."""
    if s.get("prior_pose") is None:
        return 0, None, None, None
    assert np.array_equal(s["view_pose"], np.arange(len(s["view_pose"]))), "prior tests need view_pose = identity"
    pv = np.ascontiguousarray(s["prior_pose"], np.int32)
    return len(pv), pv, np.ascontiguousarray(s["prior_center"], np.float64).copy(), np.ascontiguousarray(s["prior_weight"], np.float64)


def ref_ba_adjust_ex(s, intrinsics_opt=14, extrinsics_opt=6, structure_opt=1, threads=0, use_loss=1):
    """The reference Adjust with control points and/or pose-centre priors."""
    reg, g, reg_idx, gcp_idx = _split_gcp(s)
    opts = np.array([intrinsics_opt, extrinsics_opt, structure_opt, threads, use_loss], np.int32)
    poses = s["poses"].copy(); intr = s["intrinsics"].copy(); pts = reg["points"].copy(); gX = g["X"].copy()
    npri, pv, pc, pw = _prior_views(s)
    out = np.zeros(8); rep = ctypes.create_string_buffer(1 << 16)
    rc = ref_ba().ref_ba_adjust_ex(*_ba_args(reg, poses, intr, pts), len(gX), _P(gX), ctypes.c_long(len(g["obs_view"])),
                                   _P(g["obs_view"]), _P(g["obs_gcp"]), _P(g["obs_xy"]), ctypes.c_double(g["weight"]),
                                   npri, None if pv is None else _P(pv), None if pc is None else _P(pc),
                                   None if pw is None else _P(pw), _P(opts), _P(out), rep, 1 << 16)
    text = rep.value.decode(errors="replace")
    allpts = s["points"].copy(); allpts[reg_idx] = pts; allpts[gcp_idx] = gX
    m = re.search(r"Minimizer iterations\s+(\d+)", text)
    return dict(rc=rc, ok=bool(out[0]), initial_cost=out[1], final_cost=out[2], wall_s=out[3], prior_fit=out[4],
                poses=poses, intrinsics=intr, points=allpts, prior_center=pc, report=text,
                iterations=int(m.group(1)) if m else None)


def ref_ba_register_priors(s):
    """The registration Adjust performs before its solve when priors are used, done by the
    reference's own functions; returns the registered scene (prior_huber_a = fit^2) and the centroid."""
    reg, g, reg_idx, gcp_idx = _split_gcp(s)
    poses = s["poses"].copy(); intr = s["intrinsics"].copy(); pts = reg["points"].copy(); gX = g["X"].copy()
    npri, pv, pc, pw = _prior_views(s)
    out = np.zeros(4)
    rc = ref_ba().ref_ba_register_priors(*_ba_args(reg, poses, intr, pts), len(gX), _P(gX), ctypes.c_long(len(g["obs_view"])),
                                         _P(g["obs_view"]), _P(g["obs_gcp"]), _P(g["obs_xy"]),
                                         npri, _P(pv), _P(pc), _P(pw), _P(out))
    assert rc == 0
    t = dict(s)
    if out[0] >= 0:
        allpts = s["points"].copy(); allpts[reg_idx] = pts; allpts[gcp_idx] = gX
        t.update(poses=poses, points=allpts, prior_center=pc, prior_huber_a=float(out[0]) ** 2)
    return t, float(out[0]), out[1:4].copy()


# ----------------------------------------------------------------------------- golden GCP / prior cases
def golden_ext_scene(case, registered=True):
    """Scene of a tests/golden 'ba_ext' case from its seeds.  With registered=True the prior cases come
    back in the frame the reference's LM solved in (its own LMedS registration + centring, stored in
    tests/golden/reference_ba_ext.npz) with prior_huber_a = fit^2."""
    from openmvg_b200 import synth
    s = synth.ba_scene(**case["scene"])
    if case.get("gcp"):
        s = synth.add_gcp(s, **case["gcp"])
    if case.get("priors"):
        s = synth.add_priors(s, **case["priors"])
        if registered:
            z = np.load(os.path.join(ROOT, "tests", "golden", "reference_ba_ext.npz"))
            s["poses"] = np.ascontiguousarray(z[case["name"] + "_poses"])
            s["points"] = np.ascontiguousarray(z[case["name"] + "_points"])
            s["prior_center"] = np.ascontiguousarray(z[case["name"] + "_prior_center"])
            s["prior_huber_a"] = float(case["prior_fit"]) ** 2
    return s


# ----------------------------------------------------------------------------- cascade hashing (M9 / N2)
def ref_cascade_projections():
    P = np.zeros((128, 128), np.float32); S = np.zeros((6, 10, 128), np.float32)
    ref_match().ref_cascade_projections(_P(P), _P(S))
    return P, S


def cascade_projections():
    """The projections CascadeHasher::Init draws, from the committed fixture (generated with std::mt19937 +
    std::normal_distribution by the compiled reference driver; tests/golden/make_golden.py)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "cascade_projections.npz"))
    return np.ascontiguousarray(z["primary"]), np.ascontiguousarray(z["secondary"])


def oracle_cascade_zero_mean(descs, used):
    means = np.zeros((len(used), 128), np.float32)
    for t, k in enumerate(used):
        d = np.ascontiguousarray(descs[int(k)].reshape(-1, 128))
        oracle().oracle_cascade_image_mean(_P(d), len(d), _P(means[t]))
    zm = np.zeros(128, np.float32)
    oracle().oracle_cascade_zero_mean(_P(means), len(used), _P(zm))
    return zm


def oracle_cascade_hash(d, zm, P, S):
    d = np.ascontiguousarray(d.reshape(-1, 128))
    codes = np.zeros((len(d), 4), np.uint32); bids = np.zeros((len(d), 6), np.uint16)
    oracle().oracle_cascade_hash(_P(d), len(d), _P(zm), _P(P), _P(S), _P(codes), _P(bids))
    return codes, bids


def oracle_cascade_match_pair(di, hi, dj, hj, ratio=0.8):
    di = np.ascontiguousarray(di.reshape(-1, 128)); dj = np.ascontiguousarray(dj.reshape(-1, 128))
    out = np.zeros(2 * max(len(dj), 1), np.uint32)
    oracle().oracle_cascade_match_pair.restype = ctypes.c_int64
    n = oracle().oracle_cascade_match_pair(_P(di), len(di), _P(hi[0]), _P(hi[1]), _P(dj), len(dj), _P(hj[0]), _P(hj[1]),
                                           ctypes.c_float(ratio), _P(out))
    return out[:2 * n].reshape(-1, 2).copy()


def oracle_cascade_collection(descs, pi, pj, ratio=0.8, P=None, S=None):
    """Cascade_Hashing_Matcher_Regions.cpp:38-226 over the oracle: CSR in the order of (pi, pj), each row in
    ascending query (j) order."""
    if P is None:
        P, S = cascade_projections()
    used = sorted(set(int(x) for x in pi) | set(int(x) for x in pj))
    zm = oracle_cascade_zero_mean(descs, used)
    hashed = {k: oracle_cascade_hash(descs[k], zm, P, S) for k in used}
    offsets = [0]; rows = []
    for i, j in zip(pi, pj):
        i = int(i); j = int(j)
        m = oracle_cascade_match_pair(descs[i], hashed[i], descs[j], hashed[j], ratio) if len(descs[i]) and len(descs[j]) else np.zeros((0, 2), np.uint32)
        rows.append(m); offsets.append(offsets[-1] + len(m))
    ij = np.concatenate(rows) if rows else np.zeros((0, 2), np.uint32)
    return np.array(offsets, np.uint64), ij.astype(np.uint32), zm, hashed


def _flat(descs):
    counts = np.array([len(d) for d in descs], np.uint32)
    row_start = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.uint64)
    allrows = np.ascontiguousarray(np.concatenate([d.reshape(-1, 128) for d in descs]) if len(descs) else np.zeros((0, 128), np.uint8))
    return counts, row_start, allrows


def ref_cascade_collection(descs, pi, pj, ratio=0.8):
    counts, row_start, allrows = _flat(descs)
    pi = np.ascontiguousarray(pi, np.uint32); pj = np.ascontiguousarray(pj, np.uint32)
    offsets = np.zeros(len(pi) + 1, np.uint64)
    cap = int(sum(int(counts[j]) for j in pj)) + 1
    ij = np.zeros(2 * cap, np.uint32)
    ref_match().ref_cascade_collection.restype = ctypes.c_int64
    n = ref_match().ref_cascade_collection(_P(allrows), _P(row_start), _P(counts), len(descs), _P(pi), _P(pj),
                                           ctypes.c_uint64(len(pi)), ctypes.c_float(ratio), _P(offsets), _P(ij), ctypes.c_uint64(cap))
    assert n >= 0
    return offsets, ij[:2 * n].reshape(-1, 2).copy()


def ref_cascade_zero_mean(descs, used):
    counts, row_start, allrows = _flat(descs)
    used = np.ascontiguousarray(used, np.uint32); zm = np.zeros(128, np.float32)
    ref_match().ref_cascade_zero_mean(_P(allrows), _P(row_start), _P(counts), _P(used), len(used), _P(zm))
    return zm


def ref_cascade_hash(d, zm):
    d = np.ascontiguousarray(d.reshape(-1, 128))
    codes = np.zeros((len(d), 4), np.uint32); bids = np.zeros((len(d), 6), np.uint16)
    ref_match().ref_cascade_hash(_P(d), len(d), _P(zm), _P(codes), _P(bids))
    return codes, bids


# ----------------------------------------------------------------------------- file formats (N3)
def ref_save_descs(path, d):
    d = np.ascontiguousarray(d, np.uint8).reshape(-1, 128)
    assert ref_match().ref_save_descs(os.fsencode(path), _P(d), len(d)) == 0


def ref_save_matches(path, pI, pJ, offsets, ij):
    pI = np.ascontiguousarray(pI, np.uint32); pJ = np.ascontiguousarray(pJ, np.uint32)
    off = np.ascontiguousarray(offsets, np.uint64); m = np.ascontiguousarray(ij, np.uint32).reshape(-1)
    assert ref_match().ref_save_matches(os.fsencode(path), ctypes.c_uint64(len(pI)), _P(pI), _P(pJ), _P(off), _P(m)) == 0


def ref_load_matches(path, cap_pairs=100000, cap_m=10000000):
    pI = np.zeros(cap_pairs, np.uint32); pJ = np.zeros(cap_pairs, np.uint32); off = np.zeros(cap_pairs + 1, np.uint64); ij = np.zeros(2 * cap_m, np.uint32)
    ref_match().ref_load_matches.restype = ctypes.c_int64
    n = ref_match().ref_load_matches(os.fsencode(path), ctypes.c_uint64(cap_pairs), _P(pI), _P(pJ), _P(off), ctypes.c_uint64(cap_m), _P(ij))
    assert n >= 0, n
    return pI[:n].copy(), pJ[:n].copy(), off[:n + 1].copy(), ij[:2 * int(off[n])].reshape(-1, 2).copy()


def per_pair_digest(offsets, ij):
    """(count u16, FNV-1a u64 of the (i,j) rows) per pair of a CSR result — the form the whole-M1 / sampled-M2
    goldens are stored in (tests/golden/make_golden_m1m2.py)."""
    off = np.asarray(offsets, np.int64); m = np.ascontiguousarray(ij, np.uint32).reshape(-1, 2)
    n = len(off) - 1
    counts = np.diff(off).astype(np.uint16); h = np.zeros(n, np.uint64)
    f = oracle().oracle_fnv1a_ij
    base = m.ctypes.data
    for k in range(n):
        h[k] = f(ctypes.c_void_p(base + 8 * int(off[k])), ctypes.c_int64(int(off[k + 1] - off[k])))
    return counts, h


# ----------------------------------------------------------------------------- geometric filtering (N4)
def have_ref_geom():
    return os.path.exists(REF_GEOM_SO)


def ref_geom():
    if "rg" not in _cache:
        _cache["rg"] = ctypes.CDLL(REF_GEOM_SO)
    return _cache["rg"]


def _acransac(fn, xI, xJ, wh, precision, iterations, with_trace):
    xI = np.ascontiguousarray(xI, np.float64); xJ = np.ascontiguousarray(xJ, np.float64)
    n = len(xI)
    inl = np.zeros(max(n, 1), np.uint32); F = np.zeros(9); stats = np.zeros(2); trace = np.zeros(3, np.int32)
    args = [_P(xI), _P(xJ), n, int(wh[0]), int(wh[1]), int(wh[2]), int(wh[3]), ctypes.c_double(precision), ctypes.c_uint(iterations), _P(inl), _P(F), _P(stats)]
    if with_trace:
        args.append(_P(trace))
    k = fn(*args)
    return dict(inliers=inl[:k].copy(), F=F.reshape(3, 3), error_max=float(stats[0]), min_nfa=float(stats[1]), trace=trace)


def ref_acransac_fundamental(xI, xJ, wh=(1000, 1000, 1000, 1000), precision=4.0, iterations=2048):
    """The reference's ACRANSAC + ACKernelAdaptor<SevenPointSolver, EpipolarDistanceError, UnnormalizerT> (F_ACRobust.hpp:66-83)."""
    return _acransac(ref_geom().ref_acransac_fundamental, xI, xJ, wh, precision, iterations, False)


def oracle_acransac_fundamental(xI, xJ, wh=(1000, 1000, 1000, 1000), precision=4.0, iterations=2048):
    return _acransac(oracle().oracle_acransac_fundamental, xI, xJ, wh, precision, iterations, True)


def ref_acransac_homography(xI, xJ, wh=(1000, 1000, 1000, 1000), precision=4.0, iterations=2048):
    """The reference's ACRANSAC + ACKernelAdaptor<FourPointSolver, AsymmetricError, UnnormalizerI> (H_ACRobust.hpp:78-95)."""
    return _acransac(ref_geom().ref_acransac_homography, xI, xJ, wh, precision, iterations, False)


def oracle_acransac_homography(xI, xJ, wh=(1000, 1000, 1000, 1000), precision=4.0, iterations=2048):
    return _acransac(oracle().oracle_acransac_homography, xI, xJ, wh, precision, iterations, True)
