"""TEST INFRASTRUCTURE ONLY -- runs the reference's OWN test-suite
(/root/reference/bigsi/tests/{bloom,graph,matrix,storage} + tests/scoring.py) against the
shimmed import (oracle/ref_harness.py) with the dict storage injected into
bigsi.tests.base.CONFIGS, to pin the stand-ins (mmh3/bitarray shims) before golden
vectors are generated from them.  Build-container only.

    python oracle/run_reference_tests.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_harness import REFERENCE_ROOT, dict_config, load_reference  # noqa: E402


def main():
    load_reference()
    import pytest
    import bigsi.tests.base as base

    base.CONFIGS.append(dict_config("reftests", **base.PARAMETERS))
    t = os.path.join(REFERENCE_ROOT, "bigsi", "tests")
    args = ["-q", "-p", "no:cacheprovider", "--rootdir", "/tmp", "-o", "python_files=test_*.py scoring.py",
            os.path.join(t, "bloom"), os.path.join(t, "graph"), os.path.join(t, "matrix"),
            os.path.join(t, "storage"), os.path.join(t, "scoring.py")]
    return pytest.main(args)


if __name__ == "__main__":
    sys.exit(main())
