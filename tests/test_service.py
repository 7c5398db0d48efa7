"""The resident service layer (bigsi_b200/service.py) against renderings produced by the reference's own
request-handler functions (tests/golden/make_golden.py:golden_service; bigsi/__main__.py:41-72, 261-299)."""
import json

import numpy as np
import pytest

from bigsi_b200 import service
from tests.golden_util import load


def test_d_to_csv_matches_reference_renderings():
    g = load("service.json")
    flags = ((True, True), (True, False), (False, True), (False, False))
    for case in g["cases"]:
        for (wh, cr), want in zip(flags, case["csv"]):
            for d, w in zip(case["responses"], want):
                assert service.d_to_csv(d, wh, cr) == w
        assert "\n".join(service.d_to_csv(d, False, False) for d in case["responses"]) == case["bulk_csv"]
    assert service.d_to_csv({"query": "ACGT", "results": []}) == ""
    assert service.d_to_csv({"query": "ACGT", "results": []}, False, False) == ""


def test_read_fasta(tmp_path):
    p = tmp_path / "q.fasta"
    p.write_text(">r1 first\nACGT\nTTGA\n\n>r2\nGGGG\n>empty\n>r3\nAC\nGT")
    assert service.read_fasta(str(p)) == [("r1 first", "ACGTTTGA"), ("r2", "GGGG"), ("empty", ""), ("r3", "ACGT")]


@pytest.mark.gpu
def test_bulk_search_matches_reference_bodies(tmp_path):
    import bigsi_b200 as B

    g = load("service.json")
    cfg = {"k": g["k"], "m": g["m"], "h": g["h"], "storage-config": {"filename": "service-golden"}}
    bigsi = B.BIGSI.build(cfg, [B.BIGSI.bloom(cfg, B.seq_to_kmers(s, g["k"])) for s in g["sample_seqs"]], g["samples"])
    fasta = tmp_path / "records.fasta"
    fasta.write_text("".join(">rec%d\n%s\n%s\n" % (i, s[:50], s[50:]) for i, s in enumerate(g["records"])))
    try:
        for case in g["cases"]:
            t, sc = case["threshold"], case["score"]
            assert service.bulk_search(cfg, str(fasta), t, sc, format="json") == case["bulk_json"]
            assert service.bulk_search(cfg, g["records"], t, sc, format="csv") == case["bulk_csv"]
            assert service.search(cfg, g["records"][0], t, sc) == case["search_json"]
            assert service.search(cfg, g["records"][0], t, sc, format="csv") == case["csv"][0][0]
            lines = []
            assert service.bulk_search(cfg, g["records"], t, sc, format="json", stream=True, write=lines.append) is None
            assert [json.loads(l) for l in lines] == case["responses"]
            lines = []
            service.bulk_search(cfg, g["records"], t, sc, format="csv", stream=True, write=lines.append)
            n = len(g["records"])
            assert lines[0] == case["csv"][1][0]                      # header, no final LF
            assert lines[1 : n - 1] == case["csv"][3][1 : n - 1]       # no header, no final LF
            assert lines[n - 1] == case["csv"][2][n - 1]               # the last record keeps its final LF
    finally:
        bigsi.delete()


@pytest.mark.gpu
def test_commands_ctx_to_bloom_to_build_to_search(tmp_path):
    """`bigsi bloom` / `build` / `insert` / `merge` as functions (bigsi/__main__.py:105-183): McCortex graphs ->
    .bloom files (GPU hashing) -> TSV build (GPU transpose) -> search, against the oracle on the k-mers the
    reference's reader extracted from the same graphs (tests/golden/cortex.json)."""
    import base64

    import bigsi_b200 as B
    from oracle import oracle as O

    cases = [c for c in load("cortex.json") if c["k"] == 31]
    k, m, h = 31, 4099, 3
    cfg = {"k": k, "m": m, "h": h, "storage-config": {"filename": "cmd-main"}}
    cfg2 = {"k": k, "m": m, "h": h, "storage-config": {"filename": "cmd-other"}}
    rows_tsv, oblooms, names = [], [], []
    for i, case in enumerate(cases):
        ctx = tmp_path / ("s%d.ctx" % i)
        ctx.write_bytes(base64.b64decode(case["ctx_b64"]))
        out = service.bloom(cfg, str(ctx), str(tmp_path / "blooms" / ("s%d.bloom" % i)))
        kmers = case["extract"]["31"]
        ob = O.OracleIndex.bloom(k, m, h, [O.canonical(x) for x in kmers])
        assert np.array_equal(np.fromfile(out, dtype=np.uint8), ob)  # the .bloom file IS the filter's bytes
        rows_tsv.append("%s\t%s" % (out, "sample%d" % i))
        oblooms.append(ob)
        names.append("sample%d" % i)
    tsv = tmp_path / "build.tsv"
    tsv.write_text("\n".join(rows_tsv) + "\n")
    try:
        assert service.build(cfg, from_file=str(tsv)) == {"result": "success"}
        bigsi = B.BIGSI(cfg)
        oix = O.OracleIndex.build(k, m, h, oblooms, names)
        assert np.array_equal(bigsi.index.download_rows(0, m), oix.rows)
        q = cases[0]["extract"]["31"][3]
        assert bigsi.search(q) == oix.search(q) and bigsi.search(q)
        with pytest.raises(ValueError):
            service.build(cfg, bloomfilters=["x"], from_file=str(tsv))
        # insert one more filter from its file, then merge a second index built from the same files
        assert service.insert(cfg, rows_tsv[0].split("\t")[0], "again") == {"result": "success"}
        assert B.BIGSI(cfg).num_samples == len(names) + 1
        service.build(cfg2, bloomfilters=[r.split("\t")[0] for r in rows_tsv])  # samples default to the paths
        service.merge(cfg, cfg2)
        merged = B.BIGSI(cfg)
        assert merged.num_samples == 2 * len(names) + 1
        res = merged.search(q)
        assert [r["sample_name"] for r in res][:1] == ["sample0"] and len(res) >= 3  # sample0, again, the merged copy
    finally:
        for c in (cfg, cfg2):
            try:
                B.BIGSI(c).delete()
            except KeyError:
                pass
