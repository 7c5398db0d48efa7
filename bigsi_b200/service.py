"""Resident service layer: search / bulk_search over an index that stays in HBM.

Mirrors the request handlers of the reference's CLI/REST front end (bigsi/__main__.py:41-80,
183-314) as plain functions: the same response dictionaries, the same JSON and CSV renderings.
The reference re-opens its KV store per request and forks a worker pool for bulk_search
(`__main__.py:270-284`), every worker with its own store handle; here the index is resident
(bigsi_b200.bigsi._STORES) and one process drives the GPU, so records are searched back to back.
No web framework: wiring these functions to HTTP routes is the deployment's business.
"""
import csv
import io
import json
import os

import numpy as np

from . import bits as _bits
from .bigsi import BIGSI
from .cortex import extract_kmers_from_ctx

CITATION = "http://dx.doi.org/10.1038/s41587-018-0010-1"  # bigsi/__main__.py:71


def search_bigsi(bigsi, seq, threshold, score):
    """bigsi/__main__.py:66-72."""
    return {"query": seq, "threshold": threshold, "results": bigsi.search(seq, threshold, score), "citation": CITATION}


def d_to_csv(d, with_header=True, carriage_return=True):
    """bigsi/__main__.py:41-63: one row per hit -- the query, then the hit's values in sorted key order;
    non-numeric fields quoted; rows end in CRLF and without `carriage_return` only the final LF is dropped
    (the CR stays, as in the reference)."""
    results = d["results"]
    keys = sorted(results[0].keys()) if results else []
    out = io.StringIO()
    w = csv.writer(out, quoting=csv.QUOTE_NONNUMERIC)
    if results and with_header:
        w.writerow(["query"] + keys)
    for res in results:
        w.writerow([d["query"]] + [res[k] for k in keys])
    text = out.getvalue()
    return text if carriage_return else text[:-1]


def read_fasta(path):
    """[(record name, sequence)] in file order; sequence lines of a record are concatenated (what
    `pyfasta.Fasta(path).values()` yields, bigsi/__main__.py:273-279)."""
    records, name, chunks = [], None, []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line.startswith(">"):
                if name is not None:
                    records.append((name, "".join(chunks)))
                name, chunks = line[1:], []
            elif line and name is not None:
                chunks.append(line)
    if name is not None:
        records.append((name, "".join(chunks)))
    return records


def search(config, seq, threshold=1.0, score=False, format="json"):
    """bigsi/__main__.py:195-209 (GET/POST /search)."""
    d = search_bigsi(BIGSI(config), seq, threshold, score)
    return d_to_csv(d) if format == "csv" else json.dumps(d, indent=4)


def bulk_search(config, fasta, threshold=1.0, score=False, format="json", stream=False, write=print):
    """bigsi/__main__.py:261-314 (GET/POST /bulk_search): every record of a FASTA file (a path, or an iterable
    of sequences) against the resident index.  Not streaming: one JSON list / one CSV body without header
    lines is returned.  Streaming: every record's result goes to `write` as soon as it is there (JSON object
    per line; CSV with the header on the first record only) and nothing is returned, like the reference."""
    seqs = [s for _, s in read_fasta(fasta)] if isinstance(fasta, str) else [str(s) for s in fasta]
    bigsi = BIGSI(config)
    if not stream:
        dd = [search_bigsi(bigsi, seq, threshold, score) for seq in seqs]
        if format == "csv":
            return "\n".join(d_to_csv(d, False, False) for d in dd)
        return json.dumps(dd, indent=4)
    with_header, carriage_return = True, False  # __main__.py:300-306: the flags carry over between records
    for i, seq in enumerate(seqs):
        d = search_bigsi(bigsi, seq, threshold, score)
        if format == "csv":
            if i == 0:
                with_header, carriage_return = True, False
            elif i == len(seqs) - 1:
                carriage_return = True
            else:
                with_header, carriage_return = False, False
            write(d_to_csv(d, with_header, carriage_return))
        else:
            write(json.dumps(d))
    return None


# -- ingest and build commands (bigsi/__main__.py:118-180, bigsi/cmds/bloom.py, bigsi/cmds/build.py) -------------
def bloom(config, ctx, outfile):
    """`bigsi bloom` (__main__.py:118-131 -> cmds/bloom.py:19-27): the k-mers of a McCortex graph -> Bloom filter
    (hashed and set on the GPU) -> a .bloom file, the raw MSB-first bytes of the filter."""
    outfile = os.path.realpath(outfile)
    bf = BIGSI.bloom(config, extract_kmers_from_ctx(ctx, config["k"]))
    directory = os.path.dirname(outfile)
    if not os.path.exists(directory):
        os.makedirs(directory)
    with open(outfile, "wb") as of:
        of.write(bf.tobytes())
    return outfile


def load_bloomfilter(path, m):
    """cmds/build.py:22-28: a .bloom file -> packed uint8 [ceil(m/8)] (what BIGSI.build takes as one filter)."""
    data = np.fromfile(path, dtype=np.uint8)
    if data.size * 8 < m:
        raise ValueError("%s holds %d bits, the index needs m=%d" % (path, data.size * 8, m))
    return _bits.from_packed(data[: (m + 7) // 8], m)


def build(config, bloomfilters=(), samples=(), from_file=None):
    """`bigsi build` (__main__.py:133-174 -> cmds/build.py:43-66): .bloom files (listed directly or in a TSV of
    `bloom_path<TAB>sample_name`) become the columns of a new resident index.  The reference chunks the build to
    bound host memory (9/8 m bytes per filter) and merges the chunks; here the filters are staged to the GPU in
    groups by BIGSI.build and transposed there."""
    bloomfilters, samples = list(bloomfilters), list(samples)
    if from_file and bloomfilters:
        raise ValueError("You can only specify blooms via from_file or bloomfilters, but not both")
    if from_file:
        with open(from_file, "r") as tsv:
            for row in csv.reader(tsv, delimiter="\t"):
                bloomfilters.append(row[0])
                samples.append(row[1])
    if samples:
        assert len(samples) == len(bloomfilters)
    else:
        samples = list(bloomfilters)
    m = config["m"]
    BIGSI.build(config, [load_bloomfilter(p, m) for p in bloomfilters], samples)
    return {"result": "success"}


def insert(config, bloomfilter, sample):
    """`bigsi insert` (__main__.py:105-116 -> cmds/insert.py): one .bloom file as a new column."""
    index = BIGSI(config)
    index.insert(load_bloomfilter(bloomfilter, index.bloomfilter_size), sample)
    return {"result": "success"}


def merge(config, merge_config):
    """`bigsi merge` (__main__.py:176-183)."""
    BIGSI(config).merge(BIGSI(merge_config))
    return {"result": "merged %s into %s." % (merge_config, config)}
