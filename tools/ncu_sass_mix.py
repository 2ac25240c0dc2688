#!/usr/bin/env python
"""Instruction mix of one `ncu --set full --import-source on` report from its SASS source page: executed warp instructions and
stall samples per opcode class. Usage: python tools/ncu_sass_mix.py rep.ncu-rep [more.ncu-rep ...] (markdown on stdout)."""
import collections
import csv
import io
import re
import subprocess
import sys

CLASSES = [("mma / tcgen05", r"^(HMMA|IMMA|MMA|UTCMMA|UTCHMMA|UTCQMMA|UTC|TCGEN|QGMMA)"), ("LDS/STS (shared)", r"^(LDS|STS|LDSM|ATOMS)"),
           ("LDG/STG (global)", r"^(LDG|STG|LD\b|ST\b|RED|ATOMG|ATOM\b|UBLKCP|UTMALDG|UTMASTG|LDGSTS)"), ("FFMA/FMUL/FADD", r"^(FFMA|FMUL|FADD|FFMA2|FMNMX|FSEL|FSET|FSETP|FCHK)"),
           ("MUFU", r"^MUFU"), ("integer / logic", r"^(IMAD|IADD|LOP|SHF|LEA|ISETP|SEL|PRMT|I2F|F2I|IABS|POPC|FLO|VOTE|SGXT|BMSK|MOV|UMOV|S2R|CS2R|R2UR|UIADD|ULOP|USHF|UIMAD|ULEA|USEL|UISETP|R2P|P2R|PLOP|UPLOP)"),
           ("shuffle", r"^(SHFL|MATCH|REDUX)"), ("barrier / sync", r"^(BAR|SYNCS|WARPSYNC|BSYNC|BSSY|DEPBAR|MEMBAR|FENCE|NANOSLEEP|ERRBAR|CCTL|UTCBAR|ELECT)"),
           ("branch", r"^(BRA|EXIT|RET|CALL|BRX|JMP|YIELD|NOP|BREAK|BPT)")]


def mix(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    name = lines[0].split('","')[1][:90] if lines[0].startswith('"Kernel Name"') else rep
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    inst, stall = collections.Counter(), collections.Counter()
    for r in rows:
        op = re.sub(r"^@!?U?P\d+\s+", "", r["Source"].strip()).split(".")[0].split(" ")[0]
        cls = next((c for c, pat in CLASSES if re.match(pat, op)), "other")
        try:
            inst[cls] += int(r["Instructions Executed"])
            stall[cls] += int(r["Warp Stall Sampling (All Samples)"])
        except (ValueError, KeyError):
            pass
    ti, ts = sum(inst.values()) or 1, sum(stall.values()) or 1
    print(f"\n### `{name}`\n\n| class | warp instructions | share | stall samples | share |\n|---|---|---|---|---|")
    for c, n in inst.most_common():
        print(f"| {c} | {n} | {100.0 * n / ti:.1f} % | {stall[c]} | {100.0 * stall[c] / ts:.1f} % |")


for rep in sys.argv[1:]:
    mix(rep)
