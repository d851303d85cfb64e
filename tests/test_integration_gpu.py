"""GPU: the selection hook as an artefact.  integration/openmvg_b200.patch is applied (by oracle/Makefile, target
`patched`) to copies of the two reference sources it touches and compiled against /root/reference/src where it lies:
  * oracle/_ref/cm_stock   = the reference's openMVG_main_ComputeMatches, unmodified
  * oracle/_ref/cm_b200    = the same program with the patch: -n BRUTEFORCEL2_B200 / FASTCASCADEHASHINGL2_B200
  * oracle/_ref/patched_ba_test = a caller of the PATCHED Bundle_Adjustment_Ceres::Adjust (no caller edit)
The matches files the two ComputeMatches binaries write from the same .feat/.desc files must be byte-identical
(brute force), resp. >= 99.9 % identical (cascade hashing: float hashing order)."""
import os
import subprocess

import numpy as np
import pytest

from openmvg_b200 import matching, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
need = pytest.mark.skipif(not all(os.path.exists(os.path.join(REF, b)) for b in ("cm_stock", "cm_b200", "patched_ba_test", "make_sfm_data")),
                          reason="oracle/_ref patched binaries not built (needs /root/reference at build time: make -C oracle patched)")


def make_fixture(tmp_path, counts, seed=21):
    """sfm_data.json (written by the reference's own sfm::Save) + one .feat / .desc per view, as ComputeFeatures leaves them."""
    m = tmp_path / "matches"; m.mkdir()
    subprocess.check_call([os.path.join(REF, "make_sfm_data"), str(m / "sfm_data.json"), str(len(counts))])
    descs = synth.descriptors(len(counts), counts, seed=seed)
    rng = np.random.default_rng(seed)
    for k, d in enumerate(descs):
        matching.write_desc_file(str(m / f"img_{k:04d}.desc"), d)
        xy = rng.uniform(0, 1000, (len(d), 2))                      # distinct positions (the cascade shim drops equal coordinates)
        with open(m / f"img_{k:04d}.feat", "w") as f:
            for x, y in xy:
                f.write(f"{x:.3f} {y:.3f} 1.5 0.25\n")
    return m


def run_cm(exe, m, method, out):
    p = subprocess.run([os.path.join(REF, exe), "-i", str(m / "sfm_data.json"), "-o", str(m / out), "-n", method, "-f", "1"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
    return (m / out).read_bytes()


@need
@pytest.mark.gpu
def test_patched_compute_matches_is_byte_identical(tmp_path):
    m = make_fixture(tmp_path, [900, 850, 0, 700, 1, 1200])
    want = run_cm("cm_stock", m, "BRUTEFORCEL2", "matches.putative.stock.bin")
    got = run_cm("cm_b200", m, "BRUTEFORCEL2_B200", "matches.putative.b200.bin")
    assert len(want) > 1000 and got == want
    # the text format too
    want_t = run_cm("cm_stock", m, "BRUTEFORCEL2", "matches.putative.stock.txt")
    got_t = run_cm("cm_b200", m, "BRUTEFORCEL2_B200", "matches.putative.b200.txt")
    assert got_t == want_t
    # the patch leaves the stock methods in place
    assert run_cm("cm_b200", m, "BRUTEFORCEL2", "matches.putative.b200cpu.bin") == want


@need
@pytest.mark.gpu
def test_patched_compute_matches_cascade(tmp_path):
    m = make_fixture(tmp_path, [1500, 1400, 1300, 1200])
    run_cm("cm_stock", m, "FASTCASCADEHASHINGL2", "matches.putative.stock.txt")
    run_cm("cm_b200", m, "FASTCASCADEHASHINGL2_B200", "matches.putative.b200.txt")

    def rows(path):
        out = set(); pair = None
        with open(path) as f:
            toks = f.read().split()
        i = 0
        while i < len(toks):
            I, J, n = int(toks[i]), int(toks[i + 1]), int(toks[i + 2]); i += 3
            for _ in range(n):
                out.add((I, J, int(toks[i]), int(toks[i + 1]))); i += 2
        return out
    a, b = rows(m / "matches.putative.stock.txt"), rows(m / "matches.putative.b200.txt")
    assert len(a) > 500 and len(a & b) >= 0.999 * len(a | b)


@need
@pytest.mark.gpu
def test_patched_bundle_adjustment_ceres_routes_to_the_gpu():
    """The engines' own call (Bundle_Adjustment_Ceres by name, linear_solver_type_ poked) on the patched translation
    unit: GPU by default, Ceres with OPENMVG_B200_DISABLE=1; same final cost within 1e-6."""
    exe = os.path.join(REF, "patched_ba_test")
    env = dict(os.environ); env.pop("OPENMVG_B200_DISABLE", None)
    env["OMVG_BA_TIMING"] = "1"                             # the library prints its host phases to stderr: proof of which solver ran
    g = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=env)
    env["OPENMVG_B200_DISABLE"] = "1"
    c = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=env)
    assert g.returncode == 0 and c.returncode == 0, g.stdout + g.stderr + c.stdout + c.stderr
    assert "falling back" not in g.stdout + g.stderr
    assert "[omvg_ba timing] solve/run" in g.stderr and "[omvg_ba timing]" not in c.stderr
    cg = float(g.stdout.split("cost")[-1]); cc = float(c.stdout.split("cost")[-1])
    assert abs(cg - cc) <= 1e-6 * cc                        # two different solvers, one answer (the small scene is solved
                                                            # directly on the GPU too, so the costs may agree to every printed digit)
