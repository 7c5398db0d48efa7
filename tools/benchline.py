"""Prints the key numbers of one or more bench.py JSON lines (files given on the command line)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        with open(path) as f:
            d = json.loads(f.read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(path, "unreadable:", e)
        continue
    e2e, rf = d.get("e2e", {}), d.get("roofline", {})
    print("%s: n_gpus=%s value=%.1f M/s step=%.2f us | e2e=%.1f M/s (%.2f us) | kernel=%.2f us frac=%.3f (isolated %.2f us, %.3f) | launches=%s"
          % (path, d.get("n_gpus"), d["value"] / 1e6, 1e3 * d["ms_per_step"], e2e.get("value", 0) / 1e6,
             1e3 * e2e.get("ms_per_step", 0), 1e3 * rf.get("kernel_ms", 0), rf.get("frac", 0), 1e3 * rf.get("kernel_ms_isolated", 0),
             rf.get("frac_isolated", 0), d.get("gpu_launches")))
