"""SASS opcode histograms of the tensor-core kernels in the built objects (cuobjdump -sass), as markdown:
python scripts/sass_histogram.py > profiles/rXX_sass_histogram.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "digs_b200", "csrc", "build")
COLS = ["UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "UTCATOMSWS", "SYNCS", "MUFU", "F2FP", "FFMA", "FMUL", "FADD", "STS", "STG", "LDG",
        "RED", "ATOMS", "ATOMG", "LDL", "STL", "BAR", "PREEXIT", "MEMBAR"]
WANT = [("step_tc.o", r"k_tc_step<"), ("jet_tc.o", r"k_tc_net<"), ("jet_tc.o", r"k_tc_prep"), ("wgrad_tc.o", r"k_tc_wgrad")]

print("# SASS opcode histograms of the tensor-core kernels (`cuobjdump -sass` of the shipped objects, sm_100a)\n")
print("Counts are static instructions per kernel instantiation.  `UTCHMMA` = tcgen05.mma (kind::f16), `LDTM` = tcgen05.ld, `UTMALDG` = "
      "TMA tensor load,\n`UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops, `UTCATOMSWS` = TMEM alloc / dealloc, `MUFU` = sin / cos / rsqrt, "
      "`PREEXIT` = griddepcontrol.launch_dependents.\n")
for obj, pat in WANT:
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
    name, hist = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.split("(")[0].replace("void ", "").replace("digs::tcpath::", "").replace("(int)", "")
            hist[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and name:
            hist[name][m.group(1)] += 1
    for k in sorted(hist):
        if not re.search(pat, k):
            continue
        h = hist[k]
        print("\n## `%s` (%d instructions)\n" % (k, sum(h.values())))
        print("| " + " | ".join(COLS) + " |")
        print("|" + "---:|" * len(COLS))
        print("| " + " | ".join(str(h.get(c, 0)) for c in COLS) + " |")
