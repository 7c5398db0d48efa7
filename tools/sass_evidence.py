"""SASS instruction census per kernel of the built library (cuobjdump -sass), for profiles/: bulk async copies (UBLKCP),
mbarrier operations (SYNCS), the AND / carry-save logic (LOP3), system-scope stores / loads (peer and host memory of the
in-kernel exchange), programmatic-dependent-launch control (ACQBULK / PDL instructions), no tensor-core instructions.
Usage: python tools/sass_evidence.py [lib.so] > profiles/rN_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bigsi_b200", "libbigsi_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
names = dict(zip(re.findall(r"Function : (\S+)", out), demangle))
cur, counts = None, collections.OrderedDict()
pat = {"UBLKCP": r"\bUBLKCP", "SYNCS (mbarrier)": r"\bSYNCS", "LOP3": r"\bLOP3", "LDS.128": r"\bLDS\.128", "STG.*SYS / ST.*SYS": r"\bST[G]?\.E.*\.SYS|\bST\.E.*SYS",
       "LD*.SYS": r"\bLD[G]?\.E.*\.SYS", "BAR": r"\bBAR\.", "SHFL": r"\bSHFL", "ATOM/RED": r"\b(ATOMG?|REDG?|RED)\.", "HMMA/UTC*MMA (tensor)": r"\b(HMMA|IMMA|UTCHMMA|UTCIMMA|UTCQMMA)"}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = names.get(m.group(1), m.group(1))
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if not m:
        continue
    ins = m.group(1)
    counts[cur]["instructions"] += 1
    for key, rx in pat.items():
        if re.search(rx, ins):
            counts[cur][key] += 1
cols = ["instructions"] + list(pat)
print("kernel".ljust(70) + "".join(c[:16].rjust(18) for c in cols))
for k, c in counts.items():
    short = re.sub(r"\(.*", "", k).replace("bigsi::", "").replace("void ", "")
    print(short[:69].ljust(70) + "".join(str(c.get(col, 0)).rjust(18) for col in cols))
