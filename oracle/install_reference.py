"""TEST / BASELINE INFRASTRUCTURE ONLY -- installs the UNMODIFIED reference package into the git-ignored baseline/_ref/
(it travels to the GPU box with the snapshot; /root/reference does not exist there):

    python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>

from a copy under /tmp (the reference tree is read-only and setuptools writes build files next to setup.py), then adds
the test directories its setup.py leaves out of the wheel (bigsi/tests/{bloom,graph,matrix,storage}: the reference's
own test-suite, run against this engine by oracle/run_reference_tests.py).  --no-deps: mmh3 / bitarray / redis are not
installable here; oracle/ref_shims/ stands in for them (oracle/ref_harness.py).  Nothing under bigsi_b200/ imports it.
Called by __graft_entry__.build() when /root/reference is present.
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("BIGSI_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def install(force=False):
    marker = os.path.join(DST, "bigsi", "tests", "graph", "test_end_to_end.py")
    if os.path.exists(marker) and not force:
        return DST
    if not os.path.isdir(os.path.join(SRC, "bigsi")):
        raise RuntimeError("reference tree %s not present" % SRC)
    tmp = tempfile.mkdtemp(prefix="bigsi_ref_")
    try:
        copy = os.path.join(tmp, "reference")
        shutil.copytree(SRC, copy, ignore=shutil.ignore_patterns(".git", "example-data", "__pycache__"))
        if os.path.isdir(DST):
            shutil.rmtree(DST)
        os.makedirs(DST)
        subprocess.run([sys.executable, "-m", "pip", "install", "-q", "--no-index", "--no-build-isolation", "--no-deps",
                        "--find-links", "/opt/wheelhouse", "--target", DST, copy], check=True)
        for sub in ("bloom", "graph", "matrix", "storage"):
            shutil.copytree(os.path.join(SRC, "bigsi", "tests", sub), os.path.join(DST, "bigsi", "tests", sub),
                            ignore=shutil.ignore_patterns("__pycache__"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return DST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
