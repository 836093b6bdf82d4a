"""SASS opcode histogram per kernel of the built library (evidence for profiles/: which Blackwell instructions the hot
kernels really use -- FFMA2/FADD2/FMUL2 packed fp32, UTCHMMA/UTCMMA tcgen05, LDTM, UBLKCP/UTMALDG, HMMA legacy, ...).

  python tools/sass_hist.py [pyft8_b200/libft8_b200.so] > profiles/r02_sass_histogram.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GROUPS = [
    ("packed fp32 (FFMA2/FADD2/FMUL2)", r"^(FFMA2|FADD2|FMUL2)"),
    ("scalar fp32 (FFMA/FADD/FMUL)", r"^(FFMA|FADD|FMUL)(\.|$)"),
    ("MUFU", r"^MUFU"),
    ("int ALU (LOP3/IADD3/SHF/IMAD/PRMT/POPC/...)", r"^(LOP3|IADD3|IADD|SHF|IMAD|PRMT|POPC|LEA|ISETP|SEL|FLO|BREV|VIADD|IABS|IMNMX|VIMNMX)"),
    ("shared ld/st (LDS/STS/LDSM)", r"^(LDS|STS|LDSM)"),
    ("global ld/st (LDG/STG/LD/ST)", r"^(LDG|STG|LD|ST)(\.|$)"),
    ("local (LDL/STL = spills)", r"^(LDL|STL)"),
    ("shuffle/vote (SHFL/VOTE/MATCH/REDUX)", r"^(SHFL|VOTE|MATCH|REDUX)"),
    ("atomics (ATOM/ATOMS/ATOMG/RED)", r"^(ATOM|RED)"),
    ("barriers (BAR/SYNCS/MEMBAR/FENCE)", r"^(BAR|SYNCS|MEMBAR|FENCE|WARPSYNC|ERRBAR)"),
    ("tcgen05 MMA (UTC*MMA)", r"^UTC.*MMA"),
    ("TMEM (LDTM/STTM/UTCCP/UTCBAR/UTCATOM)", r"^(LDTM|STTM|UTCCP|UTCBAR|UTCATOM)"),
    ("TMA / bulk copy (UTMALDG/UTMASTG/UBLKCP/UTMAPF)", r"^(UTMA|UBLKCP)"),
    ("cp.async (LDGSTS)", r"^LDGSTS"),
    ("legacy tensor (HMMA/IMMA/DMMA)", r"^(HMMA|IMMA|DMMA)"),
]


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pyft8_b200", "libft8_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    kern, hist = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(.*", "", name).replace("ft8::", "").replace("void ", "")
            hist[kern] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", line)
        if m and kern:
            hist[kern][m.group(1)] += 1
    print(f"# SASS opcode histogram per kernel — `cuobjdump -sass {os.path.relpath(so, ROOT)}` (static instruction counts, sm_100a)\n")
    print("| kernel | instr | " + " | ".join(g for g, _ in GROUPS) + " |")
    print("|---|---|" + "---|" * len(GROUPS))
    for k, c in hist.items():
        tot = sum(c.values())
        row = []
        for _, pat in GROUPS:
            row.append(sum(v for op, v in c.items() if re.match(pat, op)))
        print(f"| `{k}` | {tot} | " + " | ".join(str(x) if x else "·" for x in row) + " |")
    print("\n## Top 12 opcodes (with modifiers) of the six largest kernels\n")
    for k, c in sorted(hist.items(), key=lambda kv: -sum(kv[1].values()))[:6]:
        print(f"* `{k}`: " + ", ".join(f"{op} {n}" for op, n in c.most_common(12)))


if __name__ == "__main__":
    main()
