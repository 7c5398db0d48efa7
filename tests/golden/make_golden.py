"""Generates tests/golden/*.json from the UNMODIFIED reference package.

Runs only in the build container (needs /root/reference); the reference is imported
through oracle/ref_harness.py (mmh3/bitarray/redis stand-ins + dict storage), whose
stand-ins are themselves pinned by running the reference's own tests
(oracle/run_reference_tests.py: 23 passed, 2 skipped).  The JSON files are committed
so the oracle and the CUDA path can be checked on machines without the reference.

    python tests/golden/make_golden.py

Reference entry points exercised (all under /root/reference/bigsi/):
  bloom/bloomfilter.py:5-13 (_hash, generate_hashes), utils/fncts.py:38-65,
  graph/bigsi.py:150-247 (bloom, build, search, insert), graph/index.py:42-49 (lookup),
  graph/metadata.py:33-38 (delete_sample), utils/cortex.py:23-27 (ctx k-mers).
"""
import base64
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_harness import REFERENCE_ROOT, dict_config, load_reference  # noqa: E402

load_reference()
from bigsi import BIGSI  # noqa: E402
from bigsi.bloom.bloomfilter import _hash, generate_hashes  # noqa: E402
from bigsi.utils import canonical, seq_to_kmers  # noqa: E402
from bigsi.utils.cortex import extract_kmers_from_ctx  # noqa: E402


def dump(name, obj):
    path = os.path.join(HERE, name)
    with open(path, "w") as f:
        json.dump(obj, f, separators=(",", ":"), sort_keys=False)
        f.write("\n")
    print("wrote", path, os.path.getsize(path), "bytes")


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def mutate(rng, s, nmut):
    s = list(s)
    for _ in range(nmut):
        i = rng.randrange(len(s))
        s[i] = rng.choice([c for c in "ACGT" if c != s[i]])
    return "".join(s)


# ---------------------------------------------------------------------------
def golden_hashes():
    rng = random.Random(1234)
    out = {"kat": [], "cases": [], "canonical": []}
    # the reference's own known-answer test, tests/bloom/test_create_bloomfilter.py:5-8
    for el, h, m in (("ATT", 3, 25), ("ATT", 1, 25), ("ATT", 2, 50)):
        out["kat"].append({"element": el, "h": h, "m": m, "set": sorted(generate_hashes(el, h, m))})
    ms = [25, 250, 1000, 2500, 99991, 25_000_000, 2**31 - 1]
    for k in (1, 2, 3, 4, 5, 13, 21, 31, 32, 33, 63):
        for _ in range(24):
            alpha = rng.choice(["ACGT", "ACGT", "ACGT", "ACGTN", "acgtACGT", "ACGTRYKM-"])
            kmer = rand_seq(rng, k, alpha)
            m = rng.choice(ms)
            h = rng.choice([1, 2, 3, 3, 5, 8])
            can = canonical(kmer)
            out["cases"].append({
                "kmer": kmer, "m": m, "h": h,
                # per-seed row ids of the CANONICAL k-mer (graph/index.py:62-70)
                "rows": [_hash(can, s, m) for s in range(h)],
            })
            out["canonical"].append([kmer, can])
    for kmer in ("TTT", "AAA", "ATC", "GAT", "ACGT", "TGCA", "NNN", "ACN", "NGT", "acg", "AcG"):
        out["canonical"].append([kmer, canonical(kmer)])
    dump("hashes.json", out)


# ---------------------------------------------------------------------------
def export_rows(bigsi):
    """The reference's own row bytes ("<row>:bitarray" values, storage/base.py:86-94)."""
    rows = [bytes(bigsi.storage[bigsi.storage.convert_to_bitarray_key(i)]) for i in range(bigsi.bloomfilter_size)]
    return base64.b64encode(b"".join(rows)).decode("ascii"), len(rows[0])


def run_queries(bigsi, queries):
    res = []
    for q in queries:
        seq, thr = q["seq"], q["threshold"]
        entry = {"seq": seq, "threshold": thr}
        try:
            entry["result"] = bigsi.search(seq, thr)
        except BaseException as e:  # the reference raises TypeError on queries shorter than k
            entry["raises"] = type(e).__name__
        res.append(entry)
    return res


def golden_search():
    rng = random.Random(4321)
    cases = []
    shapes = [
        # (k, m, h, n_samples, seq_len)
        (3, 1000, 3, 2, 9),
        (3, 250, 1, 2, 9),
        (5, 2500, 2, 3, 40),
        (7, 1000, 3, 7, 60),
        (11, 999, 3, 8, 80),
        (13, 1021, 5, 9, 100),
        (31, 1000, 3, 33, 150),
        (31, 5003, 3, 70, 200),
        (21, 4099, 2, 130, 120),
        (31, 1000, 3, 1, 100),
    ]
    for ci, (k, m, h, n, L) in enumerate(shapes):
        cfg = dict_config("golden%d" % ci, k, m, h)
        from bigsi.storage import get_storage

        get_storage(cfg).delete_all()
        base = rand_seq(rng, L)
        seqs = []
        for s in range(n):
            # samples are progressively mutated copies of one sequence so that
            # inexact thresholds give graded, non-trivial hit lists
            seqs.append(mutate(rng, base, rng.randrange(0, 6)) if s else base)
        samples = ["s%d" % s for s in range(n)]
        blooms = [BIGSI.bloom(cfg, seq_to_kmers(sq, k)) for sq in seqs]
        bigsi = BIGSI.build(cfg, blooms, samples)
        rows_b64, row_bytes = export_rows(bigsi)
        queries = []
        for thr in (1.0, 0.9, 0.5, 0.4, 0.0):
            queries.append({"seq": base, "threshold": thr})
            queries.append({"seq": seqs[-1], "threshold": thr})
            queries.append({"seq": mutate(rng, base, 3), "threshold": thr})
            queries.append({"seq": base[: k + 3], "threshold": thr})
        queries.append({"seq": rand_seq(rng, L), "threshold": 1.0})
        queries.append({"seq": rand_seq(rng, L), "threshold": 0.2})
        queries.append({"seq": base + base[::-1], "threshold": 0.6})
        queries.append({"seq": base[:k], "threshold": 1.0})          # a single k-mer
        queries.append({"seq": base[: k - 1], "threshold": 1.0})      # shorter than k
        queries.append({"seq": "N" * k + base[:k], "threshold": 0.3})  # non-ACGT bytes
        lookups_in = [base[:k], canonical(base[:k]), seqs[-1][-k:], rand_seq(rng, k)]
        lookup = {km: ba.to01() for km, ba in bigsi.lookup(lookups_in + lookups_in[:1]).items()}
        case = {
            "k": k, "m": m, "h": h, "samples": samples, "sample_seqs": seqs,
            "row_bytes": row_bytes, "rows_b64": rows_b64,
            "blooms_b64": [base64.b64encode(b.tobytes()).decode("ascii") for b in blooms],
            "queries": run_queries(bigsi, queries),
            "lookup_in": lookups_in + lookups_in[:1], "lookup": lookup,
        }
        # insert path (graph/bigsi.py:244-247): one more sample, then re-query
        if ci in (0, 3, 4, 6):
            new_seq = mutate(rng, base, 2)
            bigsi.insert(BIGSI.bloom(cfg, seq_to_kmers(new_seq, k)), "inserted")
            rows2_b64, row_bytes2 = export_rows(bigsi)
            case["insert"] = {
                "seq": new_seq, "sample": "inserted",
                "bloom_b64": base64.b64encode(BIGSI.bloom(cfg, seq_to_kmers(new_seq, k)).tobytes()).decode("ascii"),
                "row_bytes": row_bytes2, "rows_b64": rows2_b64,
                "queries": run_queries(bigsi, [{"seq": new_seq, "threshold": t} for t in (1.0, 0.5, 0.0)]),
                "num_samples": bigsi.num_samples,
            }
        # tombstone path (graph/metadata.py:33-38, graph/bigsi.py:186-190)
        if ci in (3, 6):
            bigsi.delete_sample(samples[0])
            case["delete"] = {
                "sample": samples[0],
                "queries": run_queries(bigsi, [{"seq": base, "threshold": t} for t in (1.0, 0.5, 0.0)]),
            }
        cases.append(case)
        bigsi.delete()
    # the reference's own end-to-end KATs (tests/graph/test_end_to_end.py:69-131)
    cfg = dict_config("golden_kat", 3, 1000, 3)
    from bigsi.storage import get_storage

    get_storage(cfg).delete_all()
    b1 = BIGSI.bloom(cfg, seq_to_kmers("ATACACAAT", 3))
    b2 = BIGSI.bloom(cfg, seq_to_kmers("ACAGAGAAC", 3))
    bg = BIGSI.build(cfg, [b1, b2], ["a", "b"])
    kat = {
        "exact": run_queries(bg, [{"seq": s, "threshold": 1.0} for s in ("ATACACAAT", "ACAGAGAAC", "ACAGTTAAC")]),
    }
    bg.delete()
    b2 = BIGSI.bloom(cfg, seq_to_kmers("ATACACAAC", 3))
    bg = BIGSI.build(cfg, [b1, b2], ["a", "b"])
    kat["inexact"] = run_queries(bg, [{"seq": "ACAGTTAAC", "threshold": 0.5}, {"seq": "ATACACAAT", "threshold": 0.5},
                                      {"seq": "ATACACAAT", "threshold": 0.0}])
    kat["inexact_lookup"] = {km: ba.to01() for km, ba in bg.lookup("AAT").items()}
    bg.delete()
    dump("search_cases.json", {"cases": cases, "reference_kat": kat})


# ---------------------------------------------------------------------------
def golden_config1():
    """BASELINE.json configs[0]: 3-sample index from example-data/*.ctx (k=31, m=1000, h=3),
    searched with example-data/query.fasta and kmers.txt.  (example-data/test-bigsi is a
    legacy v0.1 BerkeleyDB index the v0.3.8 code cannot open -- SURVEY.md section 0.)"""
    ex = os.path.join(REFERENCE_ROOT, "example-data")
    k, m, h = 31, 1000, 3
    cfg = dict_config("golden_config1", k, m, h)
    from bigsi.storage import get_storage

    get_storage(cfg).delete_all()
    samples = ["s1", "s2", "s3"]
    ctxs = ["test1.ctx", "test2.ctx", "kmers.ctx"]
    sample_kmers = [sorted(set(extract_kmers_from_ctx(os.path.join(ex, c), k))) for c in ctxs]
    blooms = [BIGSI.bloom(cfg, km) for km in sample_kmers]
    bigsi = BIGSI.build(cfg, blooms, samples)
    rows_b64, row_bytes = export_rows(bigsi)
    records, name, buf = [], None, []
    with open(os.path.join(ex, "query.fasta")) as f:
        for line in f:
            line = line.strip()
            if line.startswith(">"):
                if name is not None:
                    records.append((name, "".join(buf)))
                name, buf = line[1:], []
            elif line:
                buf.append(line)
    if name is not None:
        records.append((name, "".join(buf)))
    uniq_seqs = list(dict.fromkeys(s for _, s in records))
    with open(os.path.join(ex, "kmers.txt")) as f:
        kmers_txt = [l.strip() for l in f if l.strip()]
    queries = [{"seq": s, "threshold": t} for s in uniq_seqs for t in (1.0, 0.4, 0.0)]
    queries += [{"seq": km, "threshold": 1.0} for km in kmers_txt]
    out = {
        "k": k, "m": m, "h": h, "samples": samples, "sample_kmers": sample_kmers,
        "bloom_popcounts": [int(b.count()) for b in blooms],
        "row_bytes": row_bytes, "rows_b64": rows_b64,
        "fasta_record_names": [n for n, _ in records],
        "fasta_record_seq_index": [uniq_seqs.index(s) for _, s in records],
        "queries": run_queries(bigsi, queries),
    }
    bigsi.delete()
    dump("config1.json", out)


# ---------------------------------------------------------------------------
def _plain(o):
    """numpy scalars -> Python scalars so that the JSON round trip compares equal."""
    if isinstance(o, dict):
        return {k: _plain(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_plain(v) for v in o]
    if hasattr(o, "item") and not isinstance(o, (str, bytes)):
        return o.item()
    return o


def golden_scores():
    """score=True (graph/bigsi.py:232-239 -> scoring/score.py:96-116): the Scorer alone on random
    presence strings, and whole searches with score=True (k = 31: the Scorer hard-wires it)."""
    from bigsi.scoring import Scorer
    from bigsi.storage import get_storage

    rng = random.Random(9876)
    scorer_cases = []
    for db in (0, 1, 3, 70, 5 * 10 ** 5):
        sc = Scorer(db)
        for L in (1, 2, 3, 4, 5, 17, 34, 35, 36, 100, 257, 1200):
            for p1 in (0.0, 0.3, 0.8, 0.97, 1.0):
                # runs, not independent bits: presence strings are long stretches of 1 broken by gaps
                s, cur = [], "1" if rng.random() < p1 else "0"
                while len(s) < L:
                    run = rng.randrange(1, 60)
                    s.extend(cur * run)
                    cur = "1" if rng.random() < p1 else "0"
                s = "".join(s[:L])
                scorer_cases.append({"db_size": db, "s": s, "result": _plain(sc.score(s))})
    searches = []
    for ci, (k, m, h, n, L) in enumerate([(31, 5003, 3, 12, 220), (31, 1000, 2, 40, 120)]):
        cfg = dict_config("golden_score%d" % ci, k, m, h)
        get_storage(cfg).delete_all()
        base = rand_seq(rng, L)
        seqs = [base] + [mutate(rng, base, rng.randrange(1, 5)) for _ in range(n - 1)]
        samples = ["s%d" % i for i in range(n)]
        blooms = [BIGSI.bloom(cfg, seq_to_kmers(sq, k)) for sq in seqs]
        bigsi = BIGSI.build(cfg, blooms, samples)
        queries = []
        for thr in (1.0, 0.7, 0.3, 0.0):
            for q in (base, seqs[3], mutate(rng, base, 4), base + base[:50]):
                queries.append({"seq": q, "threshold": thr})
        queries.append({"seq": base[:k], "threshold": 1.0})       # one window: IndexError in the reference
        queries.append({"seq": base[: k + 1], "threshold": 1.0})  # two windows
        queries.append({"seq": rand_seq(rng, 80), "threshold": 1.0})  # no hit: nothing to score
        res = []
        for q in queries:
            entry = dict(q)
            try:
                entry["result"] = _plain(bigsi.search(q["seq"], q["threshold"], score=True))
            except BaseException as e:
                entry["raises"] = type(e).__name__
            res.append(entry)
        searches.append({"k": k, "m": m, "h": h, "samples": samples, "sample_seqs": seqs,
                         "blooms_b64": [base64.b64encode(b.tobytes()).decode("ascii") for b in blooms],
                         "queries": res})
        bigsi.delete()
    # the reference's own known-answer test (tests/scoring.py:10-31), read from its test file
    import ast
    import re

    src = open(os.path.join(REFERENCE_ROOT, "bigsi", "tests", "scoring.py")).read()
    kat_s = re.search(r's = "([01]+)"', src).group(1)
    kat_expected = ast.literal_eval(re.search(r"scorer\.score\(s\) == (\{.*?\})", src, re.S).group(1))
    kat = {"db_size": 5 * 10 ** 5, "s": kat_s, "expected_by_reference_test": kat_expected,
           "result": _plain(Scorer(5 * 10 ** 5).score(kat_s))}
    assert kat["result"] == kat_expected
    dump("scores.json", {"reference_kat": kat, "scorer": scorer_cases, "searches": searches})


def golden_kv_store():
    """The reference's complete key/value store (v0.3 schema) after build + insert +
    delete_sample: every key and value the reference wrote, for the import/export parity tests."""
    from bigsi.storage import get_storage
    from oracle.ref_harness import _DICT_STORES

    rng = random.Random(2468)
    out = []
    for ci, (k, m, h, n, L, extra) in enumerate([(5, 300, 2, 5, 60, 4), (31, 1000, 3, 13, 150, 1), (7, 257, 3, 8, 50, 0)]):
        cfg = dict_config("golden_kv%d" % ci, k, m, h)
        get_storage(cfg).delete_all()
        base = rand_seq(rng, L)
        seqs = [mutate(rng, base, rng.randrange(0, 6)) for _ in range(n)]
        samples = ["sample%d" % i for i in range(n)]
        if ci == 0:
            samples[2] = "0"  # a sample whose NAME is a colour number: int and string keys must not collide
        blooms = [BIGSI.bloom(cfg, seq_to_kmers(sq, k)) for sq in seqs]
        bigsi = BIGSI.build(cfg, blooms, samples)
        ins = []
        for j in range(extra):
            sq = mutate(rng, base, 2)
            bigsi.insert(BIGSI.bloom(cfg, seq_to_kmers(sq, k)), "ins%d" % j)
            ins.append(sq)
        if ci != 2:
            bigsi.delete_sample(samples[1])
        store = _DICT_STORES[cfg["storage-config"]["filename"]]
        kv = [[base64.b64encode(bytes(key)).decode("ascii"), base64.b64encode(bytes(val)).decode("ascii")]
              for key, val in sorted(store.items())]
        queries = run_queries(bigsi, [{"seq": s_, "threshold": t} for s_ in (base, seqs[0], seqs[-1]) for t in (1.0, 0.6, 0.0)])
        out.append({"k": k, "m": m, "h": h, "samples": samples, "sample_seqs": seqs, "inserted_seqs": ins,
                    "deleted": samples[1] if ci != 2 else None, "kv_b64": kv, "queries": queries,
                    "num_samples": bigsi.num_samples})
        bigsi.delete()
    dump("kv_store.json", out)


def golden_service():
    """The response renderings of the reference's request handlers (bigsi/__main__.py:41-72, 261-299).  The
    module itself cannot be imported here (hug, pyfasta, humanfriendly are absent), so the two pure functions
    `d_to_csv` and `search_bigsi` are taken out of its source with `ast` and executed unmodified."""
    import ast
    import csv
    import io
    from bigsi.storage import get_storage

    src = open(os.path.join(REFERENCE_ROOT, "bigsi", "__main__.py")).read()
    tree = ast.parse(src)
    ns = {"io": io, "csv": csv}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("d_to_csv", "search_bigsi"):
            exec(compile(ast.Module([node], []), "bigsi/__main__.py", "exec"), ns)
    d_to_csv, search_bigsi = ns["d_to_csv"], ns["search_bigsi"]
    rng = random.Random(1357)
    k, m, h, n, L = 31, 2003, 3, 9, 160
    cfg = dict_config("golden_service", k, m, h)
    get_storage(cfg).delete_all()
    base = rand_seq(rng, L)
    seqs = [base] + [mutate(rng, base, rng.randrange(1, 4)) for _ in range(n - 1)]
    samples = ["sample %d" % i for i in range(n)]  # names with a blank: quoted in CSV
    blooms = [BIGSI.bloom(cfg, seq_to_kmers(sq, k)) for sq in seqs]
    bigsi = BIGSI.build(cfg, blooms, samples)
    records = [base, seqs[2][:100], rand_seq(rng, 90), mutate(rng, base, 2), base[20:140]]
    out = {"k": k, "m": m, "h": h, "samples": samples, "sample_seqs": seqs, "records": records, "cases": []}
    for threshold, score in ((1.0, False), (0.5, False), (0.5, True), (0.0, False)):
        dd = [search_bigsi(bigsi, seq, threshold, score) for seq in records]
        out["cases"].append({
            "threshold": threshold, "score": score,
            "responses": _plain(dd),
            "csv": [[d_to_csv(d, wh, cr) for d in dd] for wh, cr in ((True, True), (True, False), (False, True), (False, False))],
            # the non-streaming bulk_search bodies (bigsi/__main__.py:285-288)
            "bulk_csv": "\n".join([d_to_csv(d, False, False) for d in dd]),
            "bulk_json": json.dumps(_plain(dd), indent=4),
            "search_json": json.dumps(_plain(dd[0]), indent=4),
        })
    bigsi.delete()
    dump("service.json", out)


def golden_merge():
    """BIGSI.merge (graph/bigsi.py:252-260): two indexes with odd column counts, a shared sample name and a
    deleted sample in the second one; the complete store of the merged index and queries on it."""
    from bigsi.storage import get_storage
    from oracle.ref_harness import _DICT_STORES

    rng = random.Random(8642)
    k, m, h = 11, 509, 3
    base = rand_seq(rng, 120)
    out = []
    for ci, (n1, n2) in enumerate([(5, 9), (8, 3), (1, 16)]):
        cfgs = [dict_config("golden_merge%d_%d" % (ci, j), k, m, h) for j in range(2)]
        for c in cfgs:
            get_storage(c).delete_all()
        seqs1 = [mutate(rng, base, rng.randrange(0, 5)) for _ in range(n1)]
        seqs2 = [mutate(rng, base, rng.randrange(0, 5)) for _ in range(n2)]
        names1 = ["a%d" % i for i in range(n1)]
        names2 = ["b%d" % i for i in range(n2)]
        names2[0] = names1[0]  # duplicate name across the two indexes
        b1 = BIGSI.build(cfgs[0], [BIGSI.bloom(cfgs[0], seq_to_kmers(s_, k)) for s_ in seqs1], names1)
        b2 = BIGSI.build(cfgs[1], [BIGSI.bloom(cfgs[1], seq_to_kmers(s_, k)) for s_ in seqs2], names2)
        if n2 > 2:
            b2.delete_sample(names2[2])  # its colour carries the tombstone name into the merge
        b1.merge(b2)
        store = _DICT_STORES[cfgs[0]["storage-config"]["filename"]]
        kv = [[base64.b64encode(bytes(key)).decode("ascii"), base64.b64encode(bytes(val)).decode("ascii")]
              for key, val in sorted(store.items())]
        merged = BIGSI(cfgs[0])
        queries = run_queries(merged, [{"seq": s_, "threshold": t} for s_ in (base, seqs1[0], seqs2[-1]) for t in (1.0, 0.7, 0.0)])
        out.append({"k": k, "m": m, "h": h, "seqs1": seqs1, "seqs2": seqs2, "names1": names1, "names2": names2,
                    "deleted_in_2": names2[2] if n2 > 2 else None, "kv_b64": kv, "queries": queries,
                    "num_samples": merged.num_samples})
        b1.delete()
        b2.delete()
    dump("merge.json", out)


def golden_cortex():
    """.ctx ingest (utils/cortex.py:23-27, 170-264): small version-6 graphs (written by bigsi_b200.cortex.write_ctx,
    several colours, several k) and the k-mers the reference's reader extracts from them."""
    import tempfile

    from bigsi_b200.cortex import write_ctx

    rng = random.Random(97531)
    out = []
    with tempfile.TemporaryDirectory() as d:
        for k, ncols, n in ((31, 1, 120), (31, 3, 40), (21, 2, 60), (13, 1, 33), (3, 1, 10)):
            kmers = [rand_seq(rng, k) for _ in range(n)]
            if k == 31:
                kmers.append("A" * 31)          # its own reverse complement's complement: T...T > A...A
                kmers.append("ACGT" * 7 + "ACG")
            path = os.path.join(d, "g.ctx")
            write_ctx(path, kmers, sample_names=["sample %d" % i for i in range(ncols)])
            with open(path, "rb") as f:
                raw = f.read()
            case = {"k": k, "ncols": ncols, "ctx_b64": base64.b64encode(raw).decode("ascii"), "extract": {}}
            for kk in sorted({k, max(1, k - 4)}):
                case["extract"][str(kk)] = list(extract_kmers_from_ctx(path, kk))
            out.append(case)
    dump("cortex.json", out)


if __name__ == "__main__":
    golden_hashes()
    golden_search()
    golden_config1()
    golden_scores()
    golden_kv_store()
    golden_service()
    golden_merge()
    golden_cortex()
