#!/usr/bin/env python
"""Per-kernel census of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md): runs
`cuobjdump -sass` on the in-tree libavld.so (no GPU needed) and counts, per kernel, UTCHMMA (tcgen05.mma, with its
.2CTA form), LDTM (tcgen05.ld), UTMALDG (TMA tensor loads, with .MULTICAST / .2CTA), UTCBAR (tcgen05.commit),
UBLKCP (cp.async.bulk), SYNCS (mbarrier), plus FFMA / FFMA2 / HMMA for the CUDA-core kernels.

    python tools/sass_census.py [path/to/libavld.so] > profiles/sass_census.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(sys.argv[1]) if len(sys.argv) > 1 else Path(__file__).resolve().parent.parent / "amphibian_vae_latent_detector_b200" / "libavld.so"
out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMALDG.MULTICAST", "UTMASTG", "UTCBAR", "UTCBAR.MULTICAST", "UBLKCP",
         "SYNCS", "UTCATOMSWS", "FFMA", "FFMA2", "FADD2", "FMUL2", "HMMA", "DFMA", "ATOMG", "REDG", "RED"]
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = demangle(m.group(1))
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not (m and kern):
        continue
    op = m.group(1)
    total[kern] += 1
    base = op.split(".")[0]
    counts[kern][base] += 1
    if base == "UTCHMMA" and ".2CTA" in op:
        counts[kern]["UTCHMMA.2CTA"] += 1
    if base == "UTMALDG" and "MULTICAST" in op:
        counts[kern]["UTMALDG.MULTICAST"] += 1
    if base == "UTCBAR" and "MULTICAST" in op:
        counts[kern]["UTCBAR.MULTICAST"] += 1
print(f"# SASS census of {lib.name} ({lib.stat().st_size} bytes), cuobjdump -sass, sm_100a")
print(f"# columns: instructions, then the watched mnemonics that occur\n")
for k, c in counts.items():
    hits = "  ".join(f"{w}={c[w]}" for w in WATCH if c[w])
    name = re.sub(r"\s+", " ", k)
    print(f"{name[:150]}\n    instr={total[k]}  {hits}\n")
