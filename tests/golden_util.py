"""Helpers to read the committed golden fixtures (tests/golden/*.json)."""
import base64
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def rows_from_b64(b64, m, row_bytes):
    a = np.frombuffer(base64.b64decode(b64), dtype=np.uint8)
    assert a.size == m * row_bytes
    return a.reshape(m, row_bytes).copy()


def bloom_from_b64(b64):
    return np.frombuffer(base64.b64decode(b64), dtype=np.uint8).copy()
