"""CPU: host-side logic that needs no device -- k-mer helpers, metadata, bit container, sharding
maths -- against the oracle / golden vectors from the unmodified reference."""
import numpy as np
import pytest

from bigsi_b200 import bits, utils
from bigsi_b200.metadata import DELETION_SPECIAL_SAMPLE_NAME, SampleMetadata
from bigsi_b200.sharded import merge_shard_hits, shard_columns, unpack_hits
from oracle import oracle as O
from tests.golden_util import load


def test_canonical_matches_reference_golden():
    for kmer, can in load("hashes.json")["canonical"]:
        assert utils.canonical(kmer) == can
        assert utils.convert_query_kmer(kmer) == O.canonical(kmer)
    assert utils.reverse_comp("ACGTN") == "NACGT"


def test_seq_to_kmers_and_unique():
    assert list(utils.seq_to_kmers("ATACACAAT", 3)) == ["ATA", "TAC", "ACA", "CAC", "ACA", "CAA", "AAT"]
    assert list(utils.seq_to_kmers("AT", 3)) == []
    assert utils.unique_kmers(["ATA", "TAC", "ATA"]) == ["ATA", "TAC"]


def test_metadata_semantics():
    # /root/reference/bigsi/tests/graph/test_metadata.py restated
    sm = SampleMetadata({})
    assert sm.num_samples == 0
    assert sm.add_sample("a") == 1 and sm.add_sample("b") == 2
    assert sm.sample_to_colour("a") == 0 and sm.colour_to_sample(1) == "b"
    with pytest.raises(ValueError):
        sm.add_sample("a")
    with pytest.raises(ValueError):
        sm.add_sample(DELETION_SPECIAL_SAMPLE_NAME)
    sm.delete_sample("a")
    assert sm.colour_to_sample(0) == DELETION_SPECIAL_SAMPLE_NAME and sm.sample_to_colour("a") is None
    assert sm.num_samples == 2
    assert sm.colours_to_samples([0, 1])[1] == "b" and sm.samples_to_colours(["a", "b"]) == {"b": 1}


def test_bit_container_msb_first():
    b = bits.from_packed(np.array([0b10100000, 0b00000001], dtype=np.uint8), 16)
    assert b.to01() == "1010000000000001" and len(b) == 16 and b.count() == 3
    assert b.tobytes() == bytes([0xA0, 0x01])
    assert (b & bits.bitarray("1111000000000000")).to01() == "1010000000000000"
    assert b[:3] == bits.bitarray("101") and b[0] is True and b[1] is False
    assert np.array_equal(bits.to_packed(b), [0xA0, 0x01])
    assert utils.non_zero_bitarrary_positions(b) == [0, 2, 15]
    with pytest.raises(TypeError):
        utils.bitwise_and([])


def test_shard_columns_and_merge():
    assert shard_columns(400_000, 8) == [(i * 50_000, (i + 1) * 50_000) for i in range(8)]
    sh = shard_columns(1001, 4)
    assert sh[0][0] == 0 and sh[-1][1] == 1001 and all(a % 8 == 0 for a, _ in sh)
    assert sum(b - a for a, b in sh) == 1001
    cols, cnts = merge_shard_hits([2, 0, 1], [[5, 1, 9], [0, 0, 0], [3, 0, 0]], [[7, 8, 0], [0, 0, 0], [4, 0, 0]], [0, 100, 200])
    assert cols.tolist() == [1, 5, 203] and cnts.tolist() == [8, 7, 4]
    Q, cap = 2, 3
    buf = np.zeros((2, Q * (2 + 2 * cap)), dtype=np.int32)
    buf[1, 0:2] = [2, 0]
    buf[1, 2 * Q : 2 * Q + 2] = [11, 12]
    buf[1, 2 * Q + Q * cap : 2 * Q + Q * cap + 2] = [5, 6]
    n, c, v = unpack_hits(buf, Q, cap)
    assert n.tolist() == [[0, 0], [2, 0]] and c[1, 0].tolist() == [11, 12, 0] and v[1, 0].tolist() == [5, 6, 0]


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU arm the driver runs beside the GPU arm) on a small workload: one JSON line with
    the keys of the bench contract; under a multi-rank launch only rank 0 prints."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--m", "100003",
           "--cols", "1000", "--kmers", "300", "--gpus", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="0", WORLD_SIZE="2"))
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config",
                "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["config"]["cols_per_gpu"] == 1000
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""
