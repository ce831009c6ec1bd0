"""SASS evidence of what the shipped library is made of: per kernel, counts of the mnemonics that prove the async-copy / TMA /
mbarrier machinery (and the absence of tensor-core ops: nothing on this path is a contraction).

    python profiles/tools/sass_summary.py > profiles/r2_sass_summary.md        (no GPU needed: cuobjdump on the in-tree .so)
"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
so = os.path.join(ROOT, "puffer_phc_b200", "lib", "libphc_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
pats = {"UBLKCP (TMA bulk)": r"UBLKCP", "LDGSTS (cp.async)": r"LDGSTS", "SYNCS (mbarrier)": r"SYNCS", "BAR.SYNC": r"BAR\.SYNC",
        "NANOSLEEP": r"NANOSLEEP", "ld/st .SYS (peer memory)": r"(LDG|STG)\.E[.\w]*\.SYS", "MEMBAR": r"MEMBAR", "fp64 (DFMA/DADD)": r"DFMA|DADD",
        "tensor-core (HMMA/UTCMMA/IMMA/QMMA)": r"HMMA|UTCMMA|IMMA|QMMA", "SHFL": r"SHFL", "ATOM/RED": r"ATOMG|ATOMS|RED\."}
print(f"# SASS summary of `puffer_phc_b200/lib/libphc_b200.so` ({', '.join(arch)}; `cuobjdump -sass`, counts of static instructions)\n")
print("| kernel | instructions | " + " | ".join(pats) + " |")
print("|---|---|" + "---|" * len(pats))
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    dem = re.sub(r"\(.*", "", dem).replace("phc::", "")
    n = len(re.findall(r"/\*[0-9a-f]{4,}\*/", f))
    print(f"| `{dem[:60]}` | {n} | " + " | ".join(str(len(re.findall(p, f))) for p in pats.values()) + " |")
