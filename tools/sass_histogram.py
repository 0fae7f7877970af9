#!/usr/bin/env python
"""Developer tool (runs here, no GPU): per-kernel SASS opcode histograms of the shipped library, committed under
profiles/ each round so that reviewers do not have to disassemble the binary to see what the kernels are made of
(tensor-core UTCHMMA, TMEM LDTM/STTM, TMA UBLKCP, packed FFMA2/FMUL2/FADD2, REDG atomics, ...).

    python tools/sass_histogram.py [tag]      # writes profiles/<tag>_sass_<kernel>.txt and profiles/<tag>_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "saro_gs_b200", "libsaro_gs_b200.so")
NOTABLE = ["UTCHMMA", "UTCMMA", "LDTM", "STTM", "UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "REDG", "RED", "ATOMG",
           "ATOMS", "MUFU", "SHFL", "VOTE", "MATCH", "BAR", "LDS", "STS", "LDG", "STG", "HMMA", "CREDUX", "REDUX"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", line)
        if m and cur:
            kernels[cur][m.group(1).split(".")[0]] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    out_dir = os.path.join(ROOT, "profiles")
    summary = ["kernel | instructions | " + " | ".join(NOTABLE)]
    for (mangled, hist), name in zip(kernels.items(), demangle):
        if not any(k in name for k in ("sgs::", "sgs_deform::", "sgs_plane::", "sgs_loss::")) or "cub::" in name:
            continue
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("sgs::", "").replace("::", "_")
        short = re.sub(r"[^A-Za-z0-9_<>,]", "", short).replace("<", "_").replace(">", "").replace(",", "_")
        total = sum(hist.values())
        with open(os.path.join(out_dir, f"{tag}_sass_{short}.txt"), "w") as f:
            f.write(f"# {name}\n# {total} SASS instructions (static), opcode histogram of libsaro_gs_b200.so\n")
            for op, n in hist.most_common():
                f.write(f"{n:6d} {op}\n")
        summary.append(f"{short} | {total} | " + " | ".join(str(hist.get(k, 0)) for k in NOTABLE))
    with open(os.path.join(out_dir, f"{tag}_sass_summary.txt"), "w") as f:
        f.write("\n".join(summary) + "\n")
    print("\n".join(summary))


if __name__ == "__main__":
    main()
