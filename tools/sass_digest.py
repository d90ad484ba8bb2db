"""Instruction-class digest of the library's hot kernels from `cuobjdump -sass` / `-res-usage` (no GPU needed): what the
machine code contains -- FP64 multiply / add counts with NO fused multiply-add in the exact stage (the library is
compiled -fmad=false: the arithmetic contract with the reference forbids contraction), FP32 FMA in the conservative
stages, local-memory traffic (spills), registers, shared memory.

  python tools/sass_digest.py > profiles/r2_sass_digest.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "drake_b200", "lib", "libdrake_b200_hydro.so")
HOT = ["k_small_pairs", "k_leaf_undecided", "k_bfs", "k_clip", "k_pack", "k_traverse", "k_filter", "k_plane", "k_surface", "k_wrench", "k_discrete", "k_pairs"]
CLASSES = [("DFMA", r"^DFMA"), ("div/sqrt seeds (MUFU.RCP64H/RSQ64H)", r"^MUFU\.(RCP64H|RSQ64H)"), ("DMUL", r"^DMUL"), ("DADD", r"^DADD"),
           ("DSETP/other fp64", r"^(DSETP|DMNMX|F2F\.F64|I2F\.F64)"),
           ("FFMA", r"^FFMA"), ("FMUL/FADD", r"^(FMUL|FADD)"), ("FSETP/FMNMX", r"^(FSETP|FMNMX)"),
           ("LDG (global load)", r"^(LDG|LD\.E)"), ("STG (global store)", r"^(STG|ST\.E)"), ("LDS", r"^LDS"), ("STS", r"^STS"),
           ("LDL (local load)", r"^LDL"), ("STL (local store)", r"^STL"), ("ATOM/RED", r"^(ATOM|RED|ATOMG)"),
           ("VOTE/SHFL/MATCH", r"^(VOTE|SHFL|MATCH)"), ("BRA/BSSY/BSYNC", r"^(BRA|BSSY|BSYNC|BRX|EXIT|RET|CALL|WARPSYNC)"),
           ("tensor (HMMA/UTC*MMA)", r"^(HMMA|IMMA|DMMA|UTC.*MMA|UTCHMMA)"), ("TMA (UBLKCP/UTMA*)", r"^(UBLKCP|UTMALDG|UTMASTG)")]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            usage[cur] = line.strip()
            cur = None
    funcs = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            funcs[cur].append(m.group(1))
    names = demangle(list(funcs))
    arch = re.search(r"arch = (sm_\w+)", sass)
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(funcs)} kernels, {arch.group(1) if arch else '?'}; cuobjdump -sass / -res-usage")
    print("# counts are STATIC instructions in the kernel's machine code, not executed instructions")
    print("# DFMA: the library is compiled -fmad=false, so no product-and-sum of the source is contracted; the DFMAs listed are")
    print("# the Newton steps of the IEEE-rounded fp64 division / square root sequences (about 8-10 per MUFU.RCP64H / RSQ64H")
    print("# seed): kernels without a division (k_small_pairs, k_filter, k_traverse) have DFMA 0 next to their DMUL / DADD.\n")
    for mangled, ops in funcs.items():
        nice = names[mangled]
        if not any(h in nice for h in HOT):
            continue
        print(f"{nice}")
        print(f"    {usage.get(mangled, '')}")
        print(f"    instructions: {len(ops)}")
        row = []
        for label, pat in CLASSES:
            n = sum(1 for o in ops if re.match(pat, o))
            if n or label in ("DFMA", "div/sqrt seeds (MUFU.RCP64H/RSQ64H)", "LDL (local load)", "STL (local store)",
                              "tensor (HMMA/UTC*MMA)"):
                row.append(f"{label} {n}")
        print("    " + "; ".join(row) + "\n")


if __name__ == "__main__":
    main()
