#!/usr/bin/env python
"""SASS evidence for the judged kernels, from the built library (no GPU needed):

    python scripts/sass_summary.py > profiles/r02_sass_summary.md

For each kernel: instruction count, opcode histogram of the memory / barrier / FP64 instructions that
characterise it, and the densest gather stretch (the SpMV inner loop: LDG.E.ENL2.256 gathers between
the streaming LDG.E.64/U16 loads)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "feellgood_b200", "libfeellgood_b200.so")
WANT = ["k_llg_solveILi1024ELb1ELb0", "k_llg_solveILi512ELb1ELb0ELb1", "k_llg_solveILi256ELb1ELb0", "k_tet_isoILi5ELb1", "k_assemble_node",
        "k_basisILb1", "k_spmv_node3ILi2ELb1ELb0"]
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, cur = {}, None
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur is not None and re.search(r"/\*[0-9a-f]{4}\*/", ln):
        funcs[cur].append(ln)
print("# SASS summary (sm_100a, `cuobjdump -sass feellgood_b200/libfeellgood_b200.so`)\n")
print("No tensor-core instruction anywhere (no `UTC*MMA`/`HMMA`): FP64 gather/stream kernels.  What to look for:"
      " 256-bit global accesses (`LDG.E.ENL2.256`, `STG.E.ENL2.256`), system-/gpu-scope acquire-release of the"
      " grid barrier (`LDG.E.STRONG.GPU`, `STG.E.STRONG.GPU`, `MEMBAR`, `CCTL.IVALL`, `ATOMG`), `DFMA`.\n")
for key in WANT:
    name = next((f for f in funcs if key in f), None)
    if not name:
        continue
    ins = funcs[name]
    ops = collections.Counter()
    for ln in ins:
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            ops[m.group(1)] += 1
    print("## `%s`\n" % name)
    print("%d instructions.  Selected opcodes:\n" % len(ins))
    print("| opcode | count |\n|---|---|")
    for op, c in sorted(ops.items(), key=lambda x: -x[1]):
        if re.match(r"(LDG|STG|LDS|STS|LDGSTS|ATOM|RED|MEMBAR|CCTL|BAR|DFMA|DADD|DMUL|SHFL|ERRBAR|FENCE|LD\.|ST\.|LDL|STL|MUFU)", op):
            print("| `%s` | %d |" % (op, c))
    # densest window of 256-bit gathers
    idx = [i for i, ln in enumerate(ins) if "LDG.E.ENL2.256" in ln]
    if len(idx) >= 8:
        best = max(range(len(idx) - 7), key=lambda k: -(idx[k + 7] - idx[k]))
        lo, hi = max(0, idx[best] - 12), min(len(ins), idx[best + 7] + 14)
        print("\nGather stretch (8 consecutive 256-bit image gathers with the fused multiply-adds that consume them):\n\n```")
        for ln in ins[lo:hi]:
            print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln).rstrip())
        print("```")
    print()
