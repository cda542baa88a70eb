"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (B200_PROFILING.md): UTCHMMA (tcgen05.mma), LDTM
(tcgen05.ld), UBLKCP (cp.async.bulk — the TMA 1-D copy), UTCBAR (tcgen05.commit), SYNCS (mbarrier), and FFMA / MUFU / BAR for the
CUDA-core kernels.  No GPU needed:   python tools/sass_counts.py > profiles/sass_counts.txt"""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

lib = sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).resolve().parent.parent / "spi_active_b200" / "libspi_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cols = ["UTCHMMA", "LDTM", "UBLKCP", "UTCBAR", "SYNCS", "FFMA", "FMUL", "FADD", "MUFU", "BAR", "LDS", "STS"]
rows, name, c = [], None, Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if name:
            rows.append((name, c))
        name, c = m.group(1), Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        c[m.group(1)] += 1
        c["total"] += 1
if name:
    rows.append((name, c))
names = subprocess.run(["c++filt", "-p"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {Path(lib).name} (sm_100a): static instruction counts per kernel")
print(f"{'kernel':58s}" + "".join(f"{k:>8s}" for k in cols) + f"{'total':>8s}")
for (raw, c), dn in zip(rows, names):
    dn = dn.replace("(anonymous namespace)::", "")
    print(f"{dn[:58]:58s}" + "".join(f"{c[k]:8d}" for k in cols) + f"{c['total']:8d}")
