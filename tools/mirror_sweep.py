"""Sweep the host mirror's pipeline knobs on config 2 (tools/cpp/bench_mirror alignReads): parts x threads, with the
per-phase host times bench_mirror prints on stderr.  Usage: python tools/mirror_sweep.py [steps]"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from paragraph_b200 import synth  # noqa: E402
import bench  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    exe = os.path.join(ROOT, "tools", "cpp", "bench_mirror")
    nodes, edges, reads = bench.workload(0)
    with tempfile.TemporaryDirectory() as tmp:
        f2 = os.path.join(tmp, "config2.txt")
        synth.write_workload_file(f2, [("DEL", nodes, edges, reads)])
        for minr, parts, first in ((0, 2, 50), (4096, 2, 50), (4096, 2, 60), (4096, 3, 50)):
            for th in (1, 8, 16):
                env = dict(os.environ, PGB_PIPELINE_PARTS=str(parts), PGB_PIPELINE_FIRST_PCT=str(first),
                           PGB_PIPELINE_MIN_READS=str(minr if minr else 1 << 30))
                r = subprocess.run([exe, f2, "alignReads", str(steps), "3", str(th), "0"], capture_output=True, text=True,
                                   timeout=600, env=env)
                if r.returncode != 0:
                    print("FAILED", minr, parts, th, r.stderr[-300:])
                    continue
                d = json.loads(r.stdout)
                print("min_reads %5d parts %d first %d%% threads %2d: %.2f Mreads/s (%.3f ms/step) | %s" % (
                    minr, parts, first, th, d["reads_per_s"] / 1e6, 1e3 * d["seconds"] / d["steps"], r.stderr.strip().split("\n")[-1]))
                sys.stdout.flush()


if __name__ == "__main__":
    main()
