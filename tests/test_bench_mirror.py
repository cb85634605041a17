"""tools/cpp/bench_mirror.cpp (the C++ bench of the drop-in surface behind bench.py's e2e_mirror / e2e_pipeline /
sweep legs) over the emulated ABI: all three modes run, keep the same reads, and the sharded mode equals the
pipeline mode."""
import json
import os
import subprocess

from conftest import ROOT
from paragraph_b200 import synth


def test_bench_mirror_modes_over_the_emulated_abi(tmp_path):
    shim = os.path.join(str(tmp_path), "libpgshim.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", shim,
                           os.path.join(ROOT, "tests", "emu", "pg_abi_shim.cpp")])
    exe = os.path.join(str(tmp_path), "bench_mirror_shim")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-pthread", "-include",
                           os.path.join(ROOT, "tests", "emu", "pg_shim_names.h"), "-o", exe,
                           os.path.join(ROOT, "tools", "cpp", "bench_mirror.cpp"), "-L" + str(tmp_path), "-lpgshim",
                           "-Wl,-rpath," + str(tmp_path)])
    sw = synth.packed_sweep(11, 6, coverage=4)
    wl = os.path.join(str(tmp_path), "wl.txt")
    synth.write_workload_file(wl, synth.sweep_as_site_list(sw))
    out = {}
    for mode, dev in (("alignReads", "0"), ("pipeline", "0"), ("sharded", "0,0")):
        r = subprocess.run([exe, wl, mode, "1", "0", "2", dev], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        out[mode] = json.loads(r.stdout)
        assert out[mode]["reads"] == len(sw["site"]) and out[mode]["sites"] == 6
        assert 0 < out[mode]["kept"] <= out[mode]["reads"]
    assert out["sharded"]["kept"] == out["pipeline"]["kept"]
    assert out["sharded"]["node_rows"] == out["pipeline"]["node_rows"] > 0
    assert out["sharded"]["devices"] == 2


def test_bench_reads_the_config2_step_profile():
    """bench.py's roofline takes its instruction counts from the newest committed ncu capture of the BENCHMARK workload
    (profiles/r<round><state>_step_ncu.json), never from a capture of another shape (r..._config4_step_ncu.json)."""
    import bench
    prof = bench.step_profile()
    assert prof is not None and "config4" not in prof["source"]
    assert "config 2" in (prof["what"] or "")
    assert prof["fill_launches"] == 3 and prof["fill_alu_inst"] > 1e8 and prof["fill_dram_bytes"] > 1e8
