#!/usr/bin/env python
"""Per-kernel SASS instruction census of the forward object (cuobjdump -sass): FP64 tensor-core MMAs (DMMA), bulk async
copies (UBLKCP = cp.async.bulk, the TMA 1-D path), mbarrier waits (SYNCS), non-MMA FP64 arithmetic, MUFU, shared-memory
loads -- the static counts behind DESIGN.md section 4.  Writes profiles/r02_fwd3_sass_counts.json and an excerpt of the
headline kernel's listing (profiles/r02_fwd3.sass).

    python tools/sass_counts.py [npbnn_b200/csrc/bnn_forward.o]
"""
import json
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "npbnn_b200", "csrc", "bnn_forward.o")
HEADLINE = "_Z6k_fwd3ILi2ELi64ELi64ELi32ELi16ELi12ELi0ELi0EEv9FwdParams"      # swish, 64-64-32-16, 12 warps, LIK, categorical

GROUPS = OrderedDict([
    ("DMMA", lambda op: op.startswith("DMMA")),
    ("UBLKCP (cp.async.bulk)", lambda op: op.startswith("UBLKCP")),
    ("SYNCS (mbarrier)", lambda op: op.startswith("SYNCS")),
    ("DFMA", lambda op: op == "DFMA"), ("DADD", lambda op: op == "DADD"), ("DMUL", lambda op: op == "DMUL"),
    ("DSETP/DMNMX", lambda op: op in ("DSETP", "DMNMX")),
    ("MUFU", lambda op: op.startswith("MUFU")),
    ("LDS", lambda op: op.startswith("LDS")), ("STS", lambda op: op.startswith("STS")),
    ("LDG", lambda op: op.startswith("LDG")), ("STG", lambda op: op.startswith("STG")),
    ("SHFL", lambda op: op.startswith("SHFL")), ("NOP", lambda op: op == "NOP"),
])


def main():
    txt = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True, check=True).stdout
    funcs, cur, lines = OrderedDict(), None, {}
    for ln in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = Counter()
            lines[cur] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(?:\.\S+)?\s", ln)
        if cur and m:
            funcs[cur][m.group(1)] += 1
            funcs[cur]["_total"] += 1
            lines[cur].append(ln.rstrip())
    out = {}
    for name, cnt in funcs.items():
        if "k_fwd" not in name:
            continue
        row = {"instructions": cnt["_total"]}
        for g, pred in GROUPS.items():
            row[g] = sum(v for k, v in cnt.items() if k != "_total" and pred(k))
        row["fp64_non_mma"] = row["DFMA"] + row["DADD"] + row["DMUL"] + row["DSETP/DMNMX"]
        out[subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name] = row
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "r02_fwd3_sass_counts.json"), "w") as f:
        json.dump(out, f, indent=1)
    if HEADLINE in lines:
        body = lines[HEADLINE]
        keep = [i for i, ln in enumerate(body) if re.search(r"\b(DMMA|UBLKCP|SYNCS|MUFU)", ln)]
        # excerpt: the first bulk copy / mbarrier set-up and the first 40 instructions around the first DMMA group
        first = next(i for i, ln in enumerate(body) if "DMMA" in ln)
        with open(os.path.join(ROOT, "profiles", "r02_fwd3.sass"), "w") as f:
            f.write("// cuobjdump -sass npbnn_b200/csrc/bnn_forward.o, kernel k_fwd3<swish,64,64,32,16, 12 warps, LIK, categorical>\n")
            f.write("// static counts: %s\n" % json.dumps(out.get(subprocess.run(["c++filt", HEADLINE], capture_output=True, text=True).stdout.strip())))
            f.write("// --- every UBLKCP / SYNCS instruction of the kernel\n")
            for i in keep:
                if re.search(r"\b(UBLKCP|SYNCS)", body[i]):
                    f.write(body[i] + "\n")
            f.write("// --- 120 instructions from the first DMMA (layer 1, epilogue of the previous weight set interleaved)\n")
            for ln in body[max(0, first - 10):first + 110]:
                f.write(ln + "\n")
    for k, v in out.items():
        if "k_fwd3<" in k:
            print("%-70s instr %6d  DMMA %4d  UBLKCP %2d  fp64 %4d  MUFU %3d" % (k[:70], v["instructions"], v["DMMA"],
                                                                                 v["UBLKCP (cp.async.bulk)"], v["fp64_non_mma"], v["MUFU"]))


if __name__ == "__main__":
    main()
