"""SASS evidence per translation unit (runs without a GPU): cuobjdump -sass over the sm_100a objects of libb200sqp.so, counting the
mnemonics that prove which hardware paths the kernels use.
    python profiles/sass_summary.py > profiles/r2_sass_summary.txt
DFMA/DMUL/DADD = fp64 pipe; DMMA = fp64 tensor-core tiles (mma.sync m8n8k4.f64); UBLKCP = TMA bulk copies (cp.async.bulk);
SYNCS = mbarrier arrive/try_wait; MUFU.RSQ64H / MUFU.RCP64H = the pivot reciprocal square roots / reciprocals; SHFL = warp shuffles;
BAR = block barriers; LDL/STL = local-memory (spill) traffic; ST.E.*.SYS / ATOM*.SYS / RED*.SYS = system-scope peer stores and arrivals.
No UTCHMMA / LDTM is expected: tcgen05.mma has no fp64 kind."""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATTERNS = [("DFMA", r"\bDFMA\b"), ("DMUL", r"\bDMUL\b"), ("DADD", r"\bDADD\b"), ("DMMA", r"\bDMMA"), ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"),
            ("MUFU.RSQ64H", r"MUFU\.RSQ64H"), ("MUFU.RCP64H", r"MUFU\.RCP64H"), ("SHFL", r"\bSHFL"), ("BAR", r"\bBAR\."), ("LDL", r"\bLDL"),
            ("STL", r"\bSTL"), ("*.SYS", r"\.SYS\b"), ("UTCHMMA", r"UTC\w*MMA"), ("LDTM", r"\bLDTM")]


def main():
    objs = sorted(glob.glob(os.path.join(ROOT, "control_box_rst_b200", "_build", "*.o")))
    print("# cuobjdump -sass per translation unit of libb200sqp.so (sm_100a); counts of instruction mnemonics over all kernels of the unit")
    print(f"{'unit':28s} {'kernels':>7s} " + " ".join(f"{n:>11s}" for n, _ in PATTERNS))
    hot = {}
    for o in objs:
        r = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True)
        if r.returncode != 0 or "Function :" not in r.stdout:
            continue
        text = r.stdout
        counts = [len(re.findall(p, text)) for _, p in PATTERNS]
        print(f"{os.path.basename(o):28s} {text.count('Function :'):7d} " + " ".join(f"{c:11d}" for c in counts))
        for fn in re.split(r"\n\s*Function : ", text)[1:]:
            name = fn.split("\n", 1)[0]
            if ("lmSolveKernel" in name and "VanDerPol" in name and "Li3ELi0ELi8E" in name) or "pipeFactorKernel" in name or "pipeLinearizeKernel" in name:
                hot[name] = [len(re.findall(p, fn)) for _, p in PATTERNS]
    print("\n# the hot kernels individually (mangled names)")
    for name, counts in hot.items():
        print(name)
        print("    " + "  ".join(f"{n}={c}" for (n, _), c in zip(PATTERNS, counts) if c))
    res = subprocess.run(["cuobjdump", "-res-usage", os.path.join(ROOT, "control_box_rst_b200", "_build", "kernels_vdp_cn.o")], capture_output=True, text=True).stdout
    print("\n# cuobjdump -res-usage kernels_vdp_cn.o (registers, stack, shared memory of the Van der Pol Crank-Nicolson kernels)")
    lines = res.splitlines()
    for i, ln in enumerate(lines):
        if "Function" in ln and "lmSolveKernel" in ln:
            print(ln.strip()[:160])
            print("   ", lines[i + 1].strip())


if __name__ == "__main__":
    sys.exit(main())
