"""GPU: the drop-in boundary through openMVG's own types.  oracle/_ref/dropin_test (built in the
container from the reference's sources in place + the host shims in openmvg_b200/host/) runs the
reference Matcher_Regions / Bundle_Adjustment_Ceres and the B200 subclasses on identical inputs
through the abstract Matcher / Bundle_Adjustment interfaces."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "dropin_test")


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/dropin_test not built (needs /root/reference at build time)")
def test_dropin_through_openmvg_types():
    p = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    print(p.stdout[-2000:], p.stderr[-2000:])
    assert p.returncode == 0 and "DROPIN OK" in p.stdout
    assert "IDENTICAL" in p.stdout
