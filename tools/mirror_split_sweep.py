"""Explicit part sizes of the host mirror's software pipeline (PGB_PIPELINE_SPLIT, per cent) on config 2 through
tools/cpp/bench_mirror alignReads, 16 and 8 host threads.  Usage: python tools/mirror_split_sweep.py [steps]"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from paragraph_b200 import synth  # noqa: E402
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
exe = os.path.join(ROOT, "tools", "cpp", "bench_mirror")
nodes, edges, reads = bench.workload(0)
with tempfile.TemporaryDirectory() as tmp:
    f2 = os.path.join(tmp, "config2.txt")
    synth.write_workload_file(f2, [("DEL", nodes, edges, reads)])
    for split in ("", "50,50", "40,60", "30,40,30", "20,60,20", "15,70,15", "25,50,25", "10,40,40,10", "20,30,30,20", "15,35,35,15"):
        for th in (16, 8):
            env = dict(os.environ)
            if split:
                env["PGB_PIPELINE_SPLIT"] = split
            r = subprocess.run([exe, f2, "alignReads", str(steps), "3", str(th), "0"], capture_output=True, text=True, timeout=600, env=env)
            if r.returncode != 0:
                print("FAILED", split, th, r.stderr[-300:])
                continue
            d = json.loads(r.stdout)
            print("split %-12s threads %2d: %.2f Mreads/s (%.3f ms/step) | %s" % (split or "(default)", th, d["reads_per_s"] / 1e6,
                  1e3 * d["seconds"] / d["steps"], r.stderr.strip().split("\n")[-1][40:]), flush=True)
