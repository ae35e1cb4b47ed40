"""Per-source-line stall samples / executed instructions of one kernel: joins `ncu --page source --csv` (SASS order) with
`nvdisasm -g` line info of the same cubin (needs -lineinfo).  Works without a GPU.
usage: python profiles/source_hotspots.py REPORT.ncu-rep OBJECT.o KERNEL_SUBSTRING [top_n]"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, obj, kern = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    body = rows[2:]
    c_samp, c_inst, c_src = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=td, check=True, capture_output=True)
        dis = ""
        for cubin in glob.glob(os.path.join(td, "*.cubin")):
            dis += subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
    # locate the kernel's text section
    lines = dis.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
    insts, cur = [], ("?", 0)
    for l in lines[start + 1:]:
        if l.startswith(".text.") or l.startswith("//--------------------- .") and ".text." not in l and insts:
            if l.startswith(".text."):
                break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            insts.append((cur, l.split("*/", 1)[1].strip()))
    if len(insts) != len(body):
        print(f"warning: {len(insts)} disassembled instructions vs {len(body)} profiled rows", file=sys.stderr)
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot_s = tot_i = 0
    for (loc, _), r in zip(insts, body):
        s, n = int(r[c_samp] or 0), int(r[c_inst] or 0)
        a = agg[loc]
        a[0] += s
        a[1] += n
        for c in stall_cols:
            v = int(r[c] or 0)
            if v:
                a[2][hdr[c]] += v
        tot_s += s
        tot_i += n
    print(f"# {kern}: {tot_s} stall samples, {tot_i} warp instructions executed")
    print(f"{'file:line':34s} {'samples':>8s} {'%':>6s} {'warp-inst':>11s} {'%':>6s}  top stall reasons")
    for loc, (s, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        reasons = ", ".join(f"{k[6:]}={v}" for k, v in st.most_common(3))
        print(f"{loc[0] + ':' + str(loc[1]):34s} {s:8d} {100 * s / max(tot_s, 1):6.2f} {n:11d} {100 * n / max(tot_i, 1):6.2f}  {reasons}")


if __name__ == "__main__":
    main()
