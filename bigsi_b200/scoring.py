"""Alignment-style score of a per-window presence string (score=True of BIGSI.search).

Host-side mirror of bigsi/scoring/score.py:7-160.  The presence strings come from the GPU
(`bigsi_b200_sequence_presence`); the arithmetic below is a handful of float operations per run
of the string and has to follow the reference operation by operation (same order, same
intermediate `round(…, 2)` calls, numpy's exp/log10) so that the result dictionaries compare
equal -- the golden vectors in tests/golden/scores.json were produced by the unmodified reference.
"""
import math

import numpy as np

KMER_LEN_IN_SCORE = 31  # hard-wired in the reference (scoring/score.py:62,98)


def _as_bits(s):
    if isinstance(s, str):
        return np.frombuffer(s.encode("ascii"), dtype=np.uint8) == ord("1")
    a = np.asarray(s)
    return (a == ord("1")) if a.dtype == np.uint8 else a.astype(bool)


def remove_short_ones(s):
    """scoring/score.py:7-16: a window survives only if the next two windows are present as well
    (positions past the end count as present); strings shorter than 3 are returned unchanged."""
    b = _as_bits(s)
    if b.size >= 3:
        nxt = np.append(b[1:], True)
        nxt2 = np.append(nxt[1:], True)
        b = b & nxt & nxt2
    return "".join("1" if x else "0" for x in b)


def tabulate_score(ss):
    """scoring/score.py:19-32: lengths of the runs of '0' and '1' -- with the reference's counting,
    in which every run except the last one is reported one longer than it is."""
    table = {"0": [], "1": []}
    n = len(ss)
    if n == 0:
        return table
    b = np.frombuffer(ss.encode("ascii"), dtype=np.uint8)
    starts = np.concatenate(([0], np.nonzero(b[1:] != b[:-1])[0] + 1))
    ends = np.concatenate((starts[1:], [n]))
    last = len(starts) - 1
    for j, (a, e) in enumerate(zip(starts.tolist(), ends.tolist())):
        table[ss[a]].append(e - a + (0 if j == last else 1))
    return table


class Scorer:
    """scoring/score.py:35-160 (same constructor arguments and defaults)."""

    def __init__(self, DB_SIZE, MATCH=1, MISMATCH=2, LAMBDA_UNGAPPED=1.330, K_UNGAPPED=0.621, LAMBDA_GAPPED=1.28,
                 K_GAPPED=0.46):
        self.DB_SIZE = DB_SIZE
        self.MATCH = MATCH
        self.MISMATCH = MISMATCH
        self.LAMBDA_UNGAPPED = LAMBDA_UNGAPPED
        self.K_UNGAPPED = K_UNGAPPED
        self.LAMBDA_GAPPED = LAMBDA_GAPPED
        self.K_GAPPED = K_GAPPED
        self.kmer_adjust = 3

    def calculate_score(self, score_counter, convert):
        """scoring/score.py:57-94: every run of absent windows is explained by between run/34 and
        run-33 mismatches; best, worst and expected score are updated run by run."""
        best = worst = expected = self.MATCH * sum(score_counter["1"])
        span = KMER_LEN_IN_SCORE + self.kmer_adjust
        most_total = 0
        fewest_total = 0
        for run in score_counter["0"]:
            fewest = float(run) / span
            most = (run - span) + 1
            if most < fewest:
                most = fewest
            most_total += most
            fewest_total += fewest
            likely = fewest + 0.05 * most
            pen_most, pen_fewest, pen_likely = self.MISMATCH * most, self.MISMATCH * fewest, self.MISMATCH * likely
            best = round(best - pen_fewest + self.MATCH * (run - pen_fewest), 2)
            worst = round(worst - pen_most + self.MATCH * (run - pen_most), 2)
            expected = round(expected - pen_likely + self.MATCH * (run - pen_likely), 2)
        return {
            "score": round(expected * convert, 2),
            "min_score": round(worst * convert, 2),
            "max_score": round(best * convert, 2),
            "max_mismatches": math.ceil(most_total),
            "min_mismatches": math.floor(fewest_total),
            "mismatches": math.ceil(math.ceil(fewest_total) + (0.05 * math.floor(most_total))),
        }

    def score(self, s):
        """scoring/score.py:96-116."""
        ss = remove_short_ones(s)
        n_windows = len(ss)
        seq_len = n_windows + KMER_LEN_IN_SCORE - 1
        d = self.calculate_score(tabulate_score(ss), seq_len / n_windows)
        d["max_nident"] = seq_len - d["min_mismatches"]
        d["nident"] = seq_len - d["mismatches"]
        d["min_nident"] = seq_len - d["max_mismatches"]
        for key in ("pident", "max_pident", "min_pident"):
            d[key] = 100 * float(d[key.replace("pident", "nident")]) / seq_len
        d["length"] = seq_len
        d["evalue"] = self.evalue(d["score"], seq_len)
        d["pvalue"] = self.pvalue(d["evalue"])
        d["log_evalue"] = round(self.log_evalue(d["score"], seq_len), 2)
        d["log_pvalue"] = round(self.log_pvalue(d["log_evalue"]), 2)
        return d

    def bitscore(self, s):
        return (self.LAMBDA_UNGAPPED * self.score(s).get("score") - np.log(self.K_UNGAPPED)) / np.log(2)

    def evalue(self, score, n):
        return self.K_UNGAPPED * self.DB_SIZE * n * np.exp(-self.LAMBDA_UNGAPPED * score)

    def pvalue(self, evalue):
        return 1 - np.exp(-evalue)

    def log_evalue(self, score, n):
        m = self.DB_SIZE if self.DB_SIZE != 0 else 1
        return round(np.log10(self.K_UNGAPPED * m * n) - self.LAMBDA_UNGAPPED * score, 2)

    def log_pvalue(self, log_evalue):
        p = 1 - np.exp(-(10 ** log_evalue))
        logp = np.log10(p) if p > 0 else -np.inf
        return round(log_evalue, 2) if logp == -np.inf else round(logp, 2)
