"""TEST INFRASTRUCTURE ONLY -- runs the reference's OWN test-suite (bigsi/tests/{bloom,graph,matrix,storage} +
tests/scoring.py of the unmodified package: /root/reference, or its install under baseline/_ref/) against the
shimmed import (oracle/ref_harness.py) with one extra storage config appended to bigsi.tests.base.CONFIGS:

    python oracle/run_reference_tests.py                 # --engine dict: pins the stand-ins (mmh3 / bitarray shims)
    python oracle/run_reference_tests.py --engine b200   # the suite on top of the HBM row store (needs a GPU):
                                                         # bigsi_b200/ref_storage.py registered in STORAGE_DICT
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_harness import dict_config, load_reference  # noqa: E402


def main():
    engine = "dict"
    if "--engine" in sys.argv:
        engine = sys.argv[sys.argv.index("--engine") + 1]
    load_reference()
    import pytest
    import bigsi.tests.base as base
    from oracle import ref_harness

    if engine == "b200":
        from bigsi_b200 import ref_storage

        ref_storage.register()
        base.CONFIGS.append(ref_storage.b200_config("reftests", **base.PARAMETERS))
    else:
        base.CONFIGS.append(dict_config("reftests", **base.PARAMETERS))
    t = os.path.join(ref_harness.REFERENCE_ROOT, "bigsi", "tests")
    args = ["-q", "-p", "no:cacheprovider", "--rootdir", "/tmp", "-o", "python_files=test_*.py scoring.py",
            os.path.join(t, "bloom"), os.path.join(t, "graph"), os.path.join(t, "matrix"),
            os.path.join(t, "storage"), os.path.join(t, "scoring.py")]
    return pytest.main(args)


if __name__ == "__main__":
    sys.exit(main())
