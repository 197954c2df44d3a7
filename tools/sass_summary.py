"""Per-kernel SASS evidence for libvdn.so: counts of the Blackwell-native mnemonics (tcgen05.mma -> UTC*MMA,
TMA -> UTMALDG / UTMASTG, tcgen05.ld/st -> LDTM / STTM, tcgen05 barriers -> UTCBAR) and of the legacy tensor path
(mma.sync -> HMMA) in every kernel of the built library. Runs here (no GPU needed):
    python tools/sass_summary.py > profiles/sass_summary.txt"""
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "video_diffusion_nnx_b200", "libvdn.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "HMMA", "RED", "ATOM"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = dict.fromkeys(KEYS, 0)
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                counts[cur][k] += 1
    names = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: instruction counts per kernel")
    print(f"{'kernel':78s} " + " ".join(f"{k:>8s}" for k in KEYS))
    tot = dict.fromkeys(KEYS, 0)
    for (mangled, c), name in zip(counts.items(), names):
        name = re.sub(r"\((int|bool)\)", "", name)
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        print(f"{name[:78]:78s} " + " ".join(f"{c[k]:8d}" for k in KEYS))
        for k in KEYS:
            tot[k] += c[k]
    print(f"{'TOTAL':78s} " + " ".join(f"{tot[k]:8d}" for k in KEYS))


if __name__ == "__main__":
    sys.exit(main())
