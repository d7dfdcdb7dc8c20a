"""profiles/r02_sass_excerpts.txt: register / spill figures, loop opcode histograms and excerpts of the hot loops of
libdtfill.so as built from the working tree (cuobjdump, no GPU needed).   python profiles/make_sass_excerpts.py > profiles/r02_sass_excerpts.txt"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "distancetransform_depthcompletion_b200", "libdtfill.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
print("SASS excerpts of the hot loops of libdtfill.so as built from this commit (cuobjdump -sass, sm_100a; opcode histograms by")
print("profiles/sass_loops.py).  ALU pipe = VIADDMNMX/VIMNMX3/LOP3/SEL/SHF/ISETP/IADD3/LEA/PRMT...; FMA pipe = IMAD/FADD/FFMA.\n")
funcs = {f.split("\n", 1)[0].strip(): f for f in re.split(r"\n\s+Function : ", sass)[1:]}
def usage(name):
    m = re.search(r"Function %s:\s*\n\s*(REG:[^\n]*)" % re.escape(name), res)
    return m.group(1).strip() if m else "?"
def ins(f):
    out = []
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m: out.append((m.group(1), m.group(2).strip()))
    return out
want = [("k2_chamferILi20ELb0ELb0ELb1E", "k2_chamfer<20,false,false,true> (narrow tiles, the dominant kernel)"),
        ("k1_mask_rows_v16IfLb1E", "k1_mask_rows_v16<float,true>"),
        ("k7_edt_columnsILi2E", "k7_edt_columns<2> (Euclidean transform, envelope pass)"),
        ("k7_edt_rowsILb1E", "k7_edt_rows<true> (Euclidean transform, row pass)")]
for pat, title in want:
    for name, f in funcs.items():
        if pat not in name: continue
        I = ins(f)
        ops = [s.split()[1] if s.startswith("@") else s.split()[0] for _, s in I]
        cnt = lambda p: sum(1 for o in ops if o.startswith(p))
        print("== " + title); print("   " + name); print("   " + usage(name))
        print("   %d instructions; VIADDMNMX %d, VIMNMX3 %d, UBLKCP (cp.async.bulk) %d, SYNCS (mbarrier) %d, FENCE.VIEW.ASYNC %d, "
              "LDGSTS %d, spills: STL %d / LDL %d" % (len(I), cnt("VIADDMNMX"), cnt("VIMNMX3"), cnt("UBLKCP"), cnt("SYNCS"),
                                                    sum(1 for _, s in I if "FENCE.VIEW.ASYNC" in s), cnt("LDGSTS"), cnt("STL"), cnt("LDL")))
        h = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "sass_loops.py"), lib, pat], capture_output=True, text=True).stdout
        for line in h.split("\n")[1:]:
            if line.strip(): print("  " + line)
        # excerpt: the first 36 instructions after the first VIADDMNMX (or the first 30 of the first loop)
        first = next((i for i, o in enumerate(ops) if o.startswith("VIADDMNMX")), 0)
        print("   excerpt, instructions %d..%d:" % (first, first + 35))
        for a, s_ in I[first:first + 36]: print("      /*%s*/ %s" % (a, s_))
        for key in ("UBLKCP", "FENCE.VIEW.ASYNC"):
            for i, (a, s_) in enumerate(I):
                if key in s_:
                    print("   around %s (instruction %d):" % (key, i))
                    for a2, s2 in I[max(0, i - 3):i + 3]: print("      /*%s*/ %s" % (a2, s2))
                    break
        print()
