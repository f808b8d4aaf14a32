"""Per-kernel SASS evidence for profiles/sass_summary.md: which Blackwell-native instructions each shipped kernel of
libmvsf_b200.so contains (B200_PROFILING.md "What proves a Blackwell-native kernel").
  python tools/sass_summary.py > profiles/sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mvsformerplusplus_b200", "libmvsf_b200.so")
PATS = [("UTCHMMA", r"\bUTCHMMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP"),
        ("SYNCS", r"\bSYNCS"), ("LDGSTS", r"\bLDGSTS"), ("FFMA2", r"\bFFMA2"), ("HMMA", r"\bHMMA"), ("LDS.128", r"\bLDS\.128"),
        ("MUFU.EX2", r"MUFU\.EX2")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        if re.search(r"/\*[0-9a-f]{4}\*/", line):
            counts[cur]["instructions"] += 1
            for name, pat in PATS:
                if re.search(pat, line):
                    counts[cur][name] += 1
    names = list(counts)
    dm = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    for n, d in zip(names, dm):
        d = re.sub(r"\(.*", "", d)          # drop the parameter list
        d = d.replace("mvsf::", "")
        demangle[n] = d
    print("# SASS summary of libmvsf_b200.so (cuobjdump -sass, sm_100a)\n")
    print("Counts of Blackwell-native instructions per kernel (`tools/sass_summary.py`).  UTCHMMA = tcgen05.mma, LDTM / STTM = "
          "tcgen05.ld / st (tensor memory), UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk, SYNCS = "
          "mbarrier, LDGSTS = cp.async, FFMA2 = packed fp32x2 FMA, HMMA = legacy mma.sync (none expected).\n")
    hdr = ["kernel", "instructions"] + [n for n, _ in PATS]
    print("| " + " | ".join(hdr) + " |")
    print("|" + "---|" * len(hdr))
    for n in names:
        c = counts[n]
        if c["instructions"] < 40:
            continue
        print("| `" + demangle[n] + "` | " + " | ".join(str(c[k]) if c[k] else "" for k in hdr[1:]) + " |")


if __name__ == "__main__":
    sys.exit(main())
