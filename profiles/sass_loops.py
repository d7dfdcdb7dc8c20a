"""Opcode histogram of the loop bodies of one kernel in a cuobjdump -sass listing (loops = backward branches).
usage: python profiles/sass_loops.py <lib.so> <mangled-name-substring> [min_body]"""
import collections
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
min_body = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s+Function : ", txt)
ALU = ("VIADDMNMX", "VIMNMX", "VIMNMX3", "LOP3", "IADD3", "VIADD", "SHF", "SEL", "ISETP", "PRMT", "LEA", "FMNMX", "POPC", "FLO",
       "BREV", "IABS", "PLOP3", "FSETP", "FSEL", "MOV", "P2R", "R2P", "SGXT", "BMSK", "LOP", "IMNMX")
FMA = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "I2FP", "FSUB")
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print(name, len(ins), "instructions")
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, s) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:\w+,\s*)?`\(\.L_x_\d+\)|\bBRA(?:\.\w+)*\s.*0x([0-9a-f]+)", s)
        if "BRA" in s:
            t = re.search(r"0x([0-9a-f]+)", s)
            if t:
                ta = int(t.group(1), 16)
                if ta <= a and ta in addr_index:
                    loops.append((addr_index[ta], i))
    for lo, hi in loops:
        if hi - lo < min_body:
            continue
        c = collections.Counter()
        for _, s in ins[lo:hi + 1]:
            op = s.split()[0]
            if op.startswith("@"):
                op = s.split()[1]
            c[op.split(".")[0]] += 1
        n = hi - lo + 1
        alu = sum(v for k, v in c.items() if k in ALU)
        fma = sum(v for k, v in c.items() if k in FMA)
        print(f"  loop [{lo},{hi}] {n} instr: ALU-pipe {alu}, FMA-pipe {fma}, other {n - alu - fma}")
        print("   ", ", ".join(f"{k} {v}" for k, v in c.most_common()))
