"""McCortex .ctx ingest (bigsi_b200/cortex.py) against k-mers the reference's reader extracted
(tests/golden/make_golden.py:golden_cortex; bigsi/utils/cortex.py:23-27) and, where the reference tree is
present, against its shipped example graphs (BASELINE config 1)."""
import base64
import os

import numpy as np
import pytest

from bigsi_b200.cortex import ctx_kmer_array, extract_kmers_from_ctx, read_header, write_ctx
from tests.golden_util import load


def test_extract_matches_reference_reader(tmp_path):
    for i, case in enumerate(load("cortex.json")):
        p = tmp_path / ("g%d.ctx" % i)
        p.write_bytes(base64.b64decode(case["ctx_b64"]))
        for kk, want in case["extract"].items():
            assert list(extract_kmers_from_ctx(str(p), int(kk))) == want
        arr = ctx_kmer_array(str(p))
        assert arr.shape[1] == case["k"] and arr.dtype == np.uint8
        k, rec, start = read_header(p.read_bytes())
        assert (k, rec) == (case["k"], 8 + 5 * case["ncols"])


def test_write_read_round_trip_and_errors(tmp_path):
    p = str(tmp_path / "r.ctx")
    write_ctx(p, ["ACGTA", "TTTTT", "GGGCC"], sample_names=("a", "b"))
    assert list(extract_kmers_from_ctx(p, 5)) == ["ACGTA", "AAAAA", "GGCCC"]  # canonical forms
    write_ctx(p, [], kmer_size=7)
    assert list(extract_kmers_from_ctx(p, 7)) == []
    bad = tmp_path / "bad.ctx"
    bad.write_bytes(b"NOTCTX" + b"\0" * 40)
    with pytest.raises(ValueError):
        ctx_kmer_array(str(bad))


@pytest.mark.skipif(not os.path.isdir("/root/reference/example-data"), reason="reference tree not present")
def test_example_graphs_give_the_config1_kmers():
    c = load("config1.json")
    for name, want in zip(("test1.ctx", "test2.ctx", "kmers.ctx"), c["sample_kmers"]):
        got = sorted(set(extract_kmers_from_ctx(os.path.join("/root/reference/example-data", name), c["k"])))
        assert got == want
