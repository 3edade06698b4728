"""Per-kernel counts of the bulk-copy (UBLKCP), mbarrier (SYNCS), cp.async (LDGSTS) and FP64 tensor-core (DMMA) SASS
instructions of the built library -> profiles/r2_sass_tma_excerpt.txt.  Runs without a GPU (cuobjdump only)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "itensornetworks.jl_b200", "lib", "libitn_b200.so")
PATS = (r"\bUBLKCP", r"\bSYNCS", r"\bLDGSTS", r"\bDMMA")


def demangle(name):
    return subprocess.run(["cu++filt", name.strip()], capture_output=True, text=True).stdout.strip()[:150]


def main(out_path):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    out = ["# SASS summary of itensornetworks.jl_b200/lib/libitn_b200.so (cuobjdump -sass, sm_100a): bulk-copy engine (UBLKCP),",
           "# mbarrier (SYNCS), cp.async (LDGSTS) and FP64 tensor-core (DMMA) instructions per kernel.  tools/sass_summary.py",
           "# kernel | UBLKCP | SYNCS | LDGSTS | DMMA", ""]
    tot = [0] * 4
    for f in funcs:
        c = [len(re.findall(p, f)) for p in PATS]
        if sum(c) == 0:
            continue
        out.append("%s | %d | %d | %d | %d" % ((demangle(f.split("\n", 1)[0]),) + tuple(c)))
        tot = [a + b for a, b in zip(tot, c)]
    out += ["", "total | %d | %d | %d | %d" % tuple(tot), "", "# first UBLKCP / SYNCS / DMMA lines of one k_fast and one k_block instance:"]
    for key in ("k_fast", "k_block"):
        for f in funcs:
            name = f.split("\n", 1)[0]
            if key in name and "UBLKCP" in f:
                out.append("## " + demangle(name))
                n = 0
                for line in f.split("\n"):
                    if re.search(r"UBLKCP|SYNCS|DMMA", line) and n < 14:
                        out.append(re.sub(r"\s+", " ", line).strip()[:140])
                        n += 1
                break
    with open(out_path, "w") as fh:
        fh.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_tma_excerpt.txt"))
