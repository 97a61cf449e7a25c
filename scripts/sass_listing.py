#!/usr/bin/env python
"""profiles/r2_sass_tcgen05.txt: which kernels of libttam.so hold tcgen05 / TMEM / TMA / cp.async instructions (cuobjdump -sass)."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OPS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "LDGSTS", "SYNCS", "UTMAPF")


def main():
    lib = ROOT / "two_tower_augmented_with_adaptive_mimic_mechanism_b200" / "libttam.so"
    out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for op in OPS:
            if re.search(r"\b" + op + r"(\b|\.)", line):
                counts[cur][op] += 1
    dem = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
    lines = ["# Blackwell-native instructions per kernel of libttam.so (cuobjdump -sass, sm_100a)",
             "# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG = cp.async.bulk.tensor (TMA load),",
             "# UTCBAR = tcgen05.commit, LDGSTS = cp.async, SYNCS = mbarrier ops.  Kernels without any of them are omitted.",
             "# regenerate: python scripts/sass_listing.py", ""]
    for (name, c), d in zip(counts.items(), dem):
        if any(c[k] for k in ("UTCHMMA", "LDTM", "UTMALDG", "LDGSTS")):
            lines.append(f"{d[:110]:110s} " + " ".join(f"{k}={v}" for k, v in sorted(c.items())))
    target = ROOT / "profiles" / "r2_sass_tcgen05.txt"
    target.write_text("\n".join(lines) + "\n")
    print(target)


if __name__ == "__main__":
    sys.exit(main())
