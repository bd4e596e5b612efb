#!/usr/bin/env python
"""Executed warp-instructions by SASS opcode class for every kernel of an .ncu-rep (source page):
how much of a kernel is address generation + load/store issue (what TMA / bulk copies would remove)
and how much is integer work.  usage: ncu_opmix.py rep [kernel-regex]"""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "k_"
names = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True,
                       errors="replace").stdout
rows = list(csv.reader(names.splitlines()))
kcol = rows[0].index("Kernel Name")
kernels = []          # (full name, index of its first launch in the report)
for i, r in enumerate(rows[2:]):
    if re.search(kre, r[kcol]) and r[kcol] not in [k for k, _ in kernels]:
        kernels.append((r[kcol], i))
CLASS = [("global ld/st", r"^(LDG|STG|LD\b|ST\b|LDGSTS|RED|ATOMG|ATOM\b|UBLKCP|UTMALDG|UTMASTG|LDGDEPBAR|DEPBAR)"),
         ("shared ld/st", r"^(LDS|STS|ATOMS|LDSM)"),
         ("local ld/st", r"^(LDL|STL)"),
         ("shuffle/vote/match", r"^(SHFL|VOTE|MATCH|REDUX)"),
         ("branch/sync", r"^(BRA|BSSY|BSYNC|EXIT|CALL|RET|WARPSYNC|BAR|JMP|BRX|YIELD|NANOSLEEP|BREAK|BMOV)"),
         ("integer alu", r"^(IADD|IADD3|IMAD|LOP|LOP3|SHF|SHL|SHR|PRMT|LEA|ISETP|SEL|POPC|FLO|BREV|IABS|IMNMX|VIMNMX|VIADD|SGXT|BMSK|MOV|PLOP3|P2R|R2P|CS2R|S2R|UMOV|UIADD3|ULOP3|USHF|ULEA|UISETP|UIMAD|USEL|UPRMT|R2UR|S2UR|ULDC|LDC|I2F|F2I|I2I|FMUL|FADD|FFMA|FSETP|MUFU|I2FP|F2FP|FSEL|HFMA2|UFLO|UPOPC|UBREV|VABSDIFF|IDP|UP2UR)"),
         ]
for kn, first in kernels:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(first),
                          "--launch-count", "1"], capture_output=True, text=True, errors="replace").stdout
    rr = list(csv.reader(raw.splitlines()))
    if len(rr) < 3:
        continue
    hdr = rr[1]; ix = {h: i for i, h in enumerate(hdr)}
    cnt = collections.Counter(); ops = collections.Counter()
    for r in rr[2:]:
        if len(r) < len(hdr) or r[ix["# Samples"]] == "# Samples":
            continue
        src = r[ix["Source"]].strip()
        src = re.sub(r"^@!?U?P\d+\s+", "", src)
        op = src.split()[0].split(".")[0] if src else "?"
        try:
            n = int(r[ix["Instructions Executed"]] or 0)
        except ValueError:
            continue
        ops[op] += n
        for cname, pat in CLASS:
            if re.match(pat, op):
                cnt[cname] += n; break
        else:
            cnt["other"] += n
    tot = sum(cnt.values()) or 1
    print(kn[:70])
    print("   " + "  ".join(f"{k} {100*v/tot:.1f}%" for k, v in cnt.most_common()))
    print("   top opcodes: " + " ".join(f"{o}:{100*v/tot:.1f}" for o, v in ops.most_common(14)))
