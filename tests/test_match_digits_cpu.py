"""The exactness argument of the digit-slice MATCH kernel, executed on the CPU.

A numpy model of what match_dig2_kernel + match_finalize_kernel<true> compute — the 32 signed digits of c0 - ceil(|b|^2/2),
the accumulator acc = q.b - h + c0, one key per 32-column chunk, the two-sided parity bound and the exact-scan fallbacks —
is checked against the plain-C oracle (oracle/match_oracle.c) on random, near-tie and extreme-norm descriptors.  The CUDA
kernels are tested against the same oracle in tests/test_match_gpu.py; this file pins the ARGUMENT, the GPU tests pin
the implementation."""
import numpy as np
import pytest

import checkers as ck
from openmvg_b200 import synth

DIG_PAD = -128 * 2 - 128 * 255 * 30
DIG_VMIN, DIG_VMAX = -979328, 971676
WEIGHTS = np.array([1] + [255] * 15 + [1] + [255] * 15, np.int64)      # the constant A tile of the fifth K-slice


def encode_digits(v):
    """32 int8 digits with sum(WEIGHTS * d) == v, as prep_digits_kernel writes them (byte 0: v - 255 M, byte 16: 0,
    the other 30 bytes: M spread evenly)."""
    v = int(v)
    M = (v + 128) // 255
    d0 = v - 255 * M
    qd, rem = divmod(M, 30)
    d = np.zeros(32, np.int64)
    d[0] = d0
    k = 0
    for pos in range(32):
        if pos in (0, 16):
            continue
        d[pos] = qd + (1 if k < rem else 0)
        k += 1
    return d


def test_digits_represent_every_value_in_range():
    rng = np.random.default_rng(0)
    vals = np.concatenate([[DIG_VMIN, DIG_VMAX, 0, -1, 1, 127, -128, 254, 255, -255, -256], rng.integers(DIG_VMIN, DIG_VMAX + 1, 20000)])
    for v in vals:
        d = encode_digits(v)
        assert d.min() >= -128 and d.max() <= 127
        assert int((WEIGHTS * d).sum()) == int(v)
    assert int((WEIGHTS * np.full(32, -128)).sum()) == DIG_PAD and DIG_PAD < DIG_VMIN      # padding sits below every real column


def model_match(db, q, ratio, group_rows=32):
    """db = database image (I), q = query image (J).  Returns the matches (i, j) the digit-slice path produces and how
    many queries needed the whole-image scan."""
    fratio = np.float32(ratio) * np.float32(ratio)
    db = db.astype(np.int64); q = q.astype(np.int64)
    nb = (db * db).sum(1); nq = (q * q).sum(1)
    h = (nb + 1) // 2
    spread_ok = len(db) == 0 or int(h.max() - h.min()) <= DIG_VMAX - DIG_VMIN
    assert spread_ok, "collection not eligible for the digit slice (the GPU falls back to the key-arithmetic kernel)"
    c0 = max(0, int(h.max()) + DIG_VMIN)
    v = c0 - h
    assert v.min() >= DIG_VMIN and v.max() <= DIG_VMAX
    n_pad = (-len(db)) % 256
    acc = q @ db.T + v[None, :]                                  # what the five MMAs leave in TMEM (exact integers)
    acc = np.concatenate([acc, np.full((len(q), n_pad), DIG_PAD, np.int64)], 1)      # zero descriptors x padding digits
    assert np.abs(acc).max() * 256 + 255 < 2 ** 31               # key = 256 * acc + group fits int32
    n_chunks = acc.shape[1] // 32
    A = acc.reshape(len(q), n_chunks, 32).max(2)                 # running max per 32-column chunk
    keys = A * 256 + (np.arange(n_chunks) * 32 // group_rows)[None, :]
    order = np.sort(keys, 1)
    K1, K2 = order[:, -1], (order[:, -2] if n_chunks > 1 else np.full(len(q), -2 ** 31))
    out, full_scans = [], 0
    d_all = nq[:, None] + nb[None, :] - 2 * (q @ db.T)           # exact distances, only used by the re-scans below

    def f32(x):
        return np.float32(int(x))

    def exact_top2(row, lo, hi):
        d = d_all[row, lo:hi]
        i1 = int(np.argmin(d)); d1 = int(d[i1])
        rest = np.delete(d, i1)
        return d1, lo + i1, (int(rest.min()) if len(rest) else 2 ** 31 - 1)

    for r in range(len(q)):
        A1, g1, A2 = int(K1[r]) >> 8, int(K1[r]) & 255, int(K2[r]) >> 8
        D1 = int(nq[r]) - 2 * A1 + 2 * c0
        none2 = A2 <= DIG_PAD
        ub2 = 2 ** 31 - 1 if none2 else int(nq[r]) - 2 * A2 + 2 * c0
        lo2 = ub2 if none2 else ub2 - 1
        if not (f32(max(D1 - 1, 0)) < fratio * f32(ub2)):        # weakest case: d1 >= D1 - 1, d2 <= D2
            continue
        full = (not none2) and A1 == A2
        keep, idx = False, 0
        if not full:
            lo, hi = g1 * group_rows, min((g1 + 1) * group_rows, len(db))
            bd1, idx, bd2 = exact_top2(r, lo, hi)
            assert bd1 in (D1 - 1, D1)                           # the parity bound
            k_hi = f32(bd1) < fratio * f32(min(ub2, bd2)); k_lo = f32(bd1) < fratio * f32(min(lo2, bd2))
            if k_hi == k_lo:
                keep = bool(k_hi)
            else:
                full = True
        if full:
            full_scans += 1
            bd1, idx, bd2 = exact_top2(r, 0, len(db))
            keep = bool(f32(bd1) < fratio * f32(bd2))
        if keep:
            out.append((idx, r))
    return np.array(out, np.uint32).reshape(-1, 2), full_scans


def near_ties(seed):
    rng = np.random.default_rng(seed)
    base = synth.descriptors(1, 24, seed=seed + 1)[0].astype(np.int64)
    rows = []
    for b in base:
        for _ in range(10):
            v = b.copy(); k = rng.integers(0, 4); idx = rng.choice(128, size=k, replace=False)
            v[idx] += rng.choice([-1, 1], size=k); rows.append(np.clip(v, 0, 255))
    rows = np.array(rows, np.uint8)
    return rows[rng.permutation(len(rows))], rows[rng.permutation(len(rows))][:150]


@pytest.mark.parametrize("ratio", [0.8, 0.95, 1.0])
def test_model_equals_oracle(ratio):
    rng = np.random.default_rng(4)
    cases = [tuple(synth.descriptors(2, [300, 260], seed=7)),
             (rng.integers(0, 256, (200, 128)).astype(np.uint8), rng.integers(0, 256, (140, 128)).astype(np.uint8)),   # |b|^2 ~ 2.8e6: c0 > 0
             (np.full((40, 128), 255, np.uint8), rng.integers(200, 256, (50, 128)).astype(np.uint8)),                  # largest norms: c0 = 3.2e6
             (synth.descriptors(1, 20, seed=3)[0], synth.descriptors(1, 70, seed=5)[0]),                               # one real chunk: "no second chunk"
             near_ties(1), near_ties(2)]
    scans = 0
    for db, q in cases:
        got, n_full = model_match(db, q, ratio)
        want = ck.oracle_match_pair(db, q, ratio)
        assert np.array_equal(got, want), (len(got), len(want))
        scans += n_full
    if ratio == 1.0:
        assert scans > 0          # the near-tie collections do reach the whole-image scan
