#!/usr/bin/env python
"""Per-kernel SASS census of libmreserve_b200.so: which kernels use the Blackwell tensor / TMA path.

    python tools/sass_census.py > profiles/sass_census_r02.md

Counts, per kernel (template instantiations of one kernel are merged), the SASS mnemonics that prove the path
(/opt/skills/guides/B200_PROFILING.md): UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA
load / store, HMMA = legacy mma.sync, MUFU = special-function unit ops.  Needs cuobjdump (no GPU)."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "merlot_reserve_b200" / "lib" / "libmreserve_b200.so"
PATTERNS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA", "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    insts = collections.Counter()
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            raw = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            raw = raw.replace("(anonymous namespace)::", "")
            base = re.sub(r"^(void|int|float)\s+", "", raw)
            name = re.split(r"[<(]", base)[0].split("::")[-1]
            kernels.setdefault(name, [collections.Counter(), 0])
            kernels[name][1] += 1
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[name][0]["_total"] += 1
            for p in PATTERNS:
                if op.startswith(p):
                    kernels[name][0][p] += 1
    print("# SASS census of libmreserve_b200.so (cuobjdump -sass; `tools/sass_census.py`)\n")
    print("UTCHMMA = tcgen05.mma; LDTM / STTM = tcgen05.ld / st (TMEM); UTMALDG / UTMASTG / UTMAREDG = TMA load / store / "
          "reduce; HMMA = mma.sync (legacy path); MUFU = SFU ops.  Instantiations of one template are summed.\n")
    print("| kernel | instantiations | SASS instr | " + " | ".join(PATTERNS) + " |")
    print("|---|---:|---:|" + "---:|" * len(PATTERNS))
    tot = collections.Counter()
    for k, (c, n) in sorted(kernels.items(), key=lambda kv: -kv[1][0]["UTCHMMA"] * 1000000 - kv[1][0]["_total"]):
        print(f"| `{k}` | {n} | {c['_total']} | " + " | ".join(str(c[p]) for p in PATTERNS) + " |")
        tot.update(c)
    print(f"| **total** | {sum(n for _, n in kernels.values())} | {tot['_total']} | " + " | ".join(str(tot[p]) for p in PATTERNS) + " |")
    tc = [k for k, (c, _) in kernels.items() if c["UTCHMMA"]]
    hm = [k for k, (c, _) in kernels.items() if c["HMMA"]]
    print(f"\ntcgen05 kernels: {', '.join('`%s`' % k for k in tc)}.")
    print(f"\nmma.sync kernels (short sequences, T <= 64: audio T = 31, span T = 16; profiles/attention_r01_notes.md): "
          f"{', '.join('`%s`' % k for k in hm)}.")


if __name__ == "__main__":
    sys.exit(main())
