"""Run ON THE GPU BOX (under gpurun): one `ncu --set full` capture of the dominant kernel of the bench workload, reduced to the
per-launch DRAM traffic and the headline counters, stamped with the hash of the kernel sources the library was built from.

    python profiles/capture_traffic.py [--cfg 1] [--batch 4096] [--kernel lmSolve] [--tag r2]

writes gpurun_out/traffic_<workload>.json (merge into profiles/traffic.json), gpurun_out/<tag>_<kernel>_cfg<cfg>_b<batch>.ncu-rep
and the text summary gpurun_out/<tag>_<kernel>_cfg<cfg>_b<batch>.txt (copy to profiles/).  bench.py only reports `roofline.traffic`
when the stamp equals the hash of the tree it runs from."""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from control_box_rst_b200 import problems  # noqa: E402
from profiles.summarize import KEYS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", type=int, default=1)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--kernel", default="lmSolve")
ap.add_argument("--tag", default="r2")
ap.add_argument("--skip", type=int, default=2, help="matching launches to skip (warm-up solves)")
ap.add_argument("--precision", default="f64")
a = ap.parse_args()

out_dir = os.path.join(ROOT, "gpurun_out")
os.makedirs(out_dir, exist_ok=True)
base = os.path.join(out_dir, f"{a.tag}_{a.kernel}_cfg{a.cfg}_b{a.batch}" + ("" if a.precision == "f64" else "_" + a.precision))
cmd = ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", f"regex:{a.kernel}", "-s", str(a.skip), "-c", "1", "-f", "-o", base,
       sys.executable, os.path.join(ROOT, "profiles", "prof_solve.py"), "--cfg", str(a.cfg), "--batch", str(a.batch), "--solves", str(a.skip + 1), "--precision", a.precision]
subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
raw = subprocess.run(["ncu", "-i", base + ".ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
get = lambda k: r[hdr.index(k)]  # noqa: E731
lines = [f"# ncu --set full --clock-control none, one launch after {a.skip} warm-up solves: python profiles/prof_solve.py --cfg {a.cfg} --batch {a.batch}",
         f"# kernel sources {bench.csrc_hash()}", f"kernel: {get('Kernel Name')}"]
for h, u, v in zip(hdr, units, r):
    if h in KEYS:
        lines.append(f"  {h:90s} {v} {u}")
open(base + ".txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))


def num(k):
    v, u = float(get(k).replace(",", "")), units[hdr.index(k)]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


ocp, kw, _ = problems.config(a.cfg)
key = bench.workload_name(a.cfg, ocp, a.batch, kw["iterations"]).split("_per_gpu")[0] + ("" if a.precision == "f64" else "_" + a.precision)
entry = {key: {"kernel": get("Kernel Name"), "dram_bytes_read": int(num("dram__bytes_read.sum")), "dram_bytes_write": int(num("dram__bytes_write.sum")),
               "csrc_sha256": bench.csrc_hash(),
               "source": f"profiles/{os.path.basename(base)}.txt (ncu --set full, one launch, B={a.batch}, {kw['iterations']} LM iterations, kernel sources {bench.csrc_hash()})"}}
json.dump(entry, open(os.path.join(out_dir, f"traffic_{key}.json"), "w"), indent=1)
print(json.dumps(entry))
