"""The reference's OWN test-suite on top of this engine's HBM row store, by injection: the
unmodified `bigsi` package (installed under baseline/_ref/ by oracle/install_reference.py, or /root/reference in the
build container) gets bigsi_b200.ref_storage.B200Storage registered in its STORAGE_DICT and a "b200" config appended to
bigsi.tests.base.CONFIGS; its tests/{bloom,graph,matrix,storage} + tests/scoring.py then run unmodified (in a
subprocess: the mmh3 / bitarray stand-ins of oracle/ref_shims must not leak into this process)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_present():
    return any(os.path.isdir(os.path.join(p, "bigsi", "tests", "graph"))
               for p in (os.environ.get("BIGSI_REFERENCE_ROOT", "/root/reference"), os.path.join(ROOT, "baseline", "_ref")))


@pytest.mark.gpu
def test_reference_suite_on_hbm_row_store():
    if not _reference_present():
        pytest.skip("the reference package is not installed (oracle/install_reference.py)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_reference_tests.py"), "--engine", "b200"],
                       capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) >= 23, tail  # the same 23 tests that pass on the reference's dict storage


def test_reference_suite_on_dict_storage_pins_the_shims():
    """CPU: the same suite on the plain dict storage -- the stand-ins for mmh3 / bitarray behave like the real ones
    as far as the reference's own tests can tell (the golden vectors were generated through them)."""
    if not _reference_present():
        pytest.skip("the reference package is not installed (oracle/install_reference.py)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_reference_tests.py")], capture_output=True, text=True,
                       timeout=1500)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) >= 23
