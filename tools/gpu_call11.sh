#!/bin/bash
set -u
mkdir -p gpurun_out
cat > /tmp/mini.py <<'PY'
import sys; sys.path.insert(0, '.')
from paragraph_b200 import capi, synth
nodes, edges, reads = synth.config2(seed=1, n_reads=64)
ctx = capi.Context(0); ctx.add_graph(nodes, edges)
print(len(ctx.align(reads)), "reads aligned")
PY
for lib in paragraph_b200/libpgalign.so ab_build/libpg_static_bar.so; do
  echo "== synccheck $lib" >> gpurun_out/r02e_synccheck_variants.txt
  PG_LIB=$lib timeout 300 compute-sanitizer --tool synccheck --print-limit 3 python /tmp/mini.py 2>&1 | grep -E "Barrier error|ERROR SUMMARY|reads aligned|Device Frame" | head -12 >> gpurun_out/r02e_synccheck_variants.txt
done
echo "== synccheck default build, PG_NO_TMA=1" >> gpurun_out/r02e_synccheck_variants.txt
PG_NO_TMA=1 timeout 300 compute-sanitizer --tool synccheck --print-limit 3 python tools/sanitize_check.py 2>&1 | grep -E "Barrier error|ERROR SUMMARY|sanitize|Device Frame" | head -12 >> gpurun_out/r02e_synccheck_variants.txt
cat gpurun_out/r02e_synccheck_variants.txt
