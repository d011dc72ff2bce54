"""Counts the Blackwell-native SASS mnemonics (B200_PROFILING.md "What proves a Blackwell-native kernel") per kernel of
liblas_b200.so:  python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "las_pytorch_b200", "liblas_b200.so")
PAT = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "LDGSTS", "UCGABAR", "ELECT",
       "RED", "LDS", "STS", "LDG", "STG", "MUFU.TANH", "MUFU.EX2"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kern, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(anonymous namespace\)::|las::|^void ", "", kern)
            kern = re.sub(r"^void ", "", kern).split("(")[0]
            counts[kern] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and kern:
            op = m.group(1)
            for p in PAT:
                if op == p or op.startswith(p + "."):
                    counts[kern][p] += 1
            counts[kern]["(instructions)"] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: instruction counts per kernel (sm_100a)")
    for k, c in counts.items():
        if not (c["UTCHMMA"] or c["UTMALDG"] or c["UBLKCP"] or c["LDTM"]):
            continue
        print(f"{k}\n   " + "  ".join(f"{p}={c[p]}" for p in ["(instructions)"] + PAT if c[p]))
    legacy = [k for k, c in counts.items() if c["HMMA"]]
    print(f"# kernels with legacy HMMA (mma.sync / wmma): {legacy or 'none'}")


if __name__ == "__main__":
    main()
