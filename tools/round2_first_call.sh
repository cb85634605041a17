#!/bin/bash
# First GPU call of the next round: measure and verify the prepared fill-kernel variant (DESIGN.md section 10,
# -DPG_SPEC_DEAD=1: speculative "no gap alive" blocks).  Build the variants HERE first (no GPU needed):
#     python tools/ab_variants.py build
# then:
#     gpurun --timeout 600 -- 'bash tools/round2_first_call.sh'
# Outputs land in gpurun_out/: ab.txt (fill / trace ms + result digest per variant -- the digests must agree),
# spec_tests.txt (the GPU suite on the variant), spec_fuzz.txt (differential fuzz vs the oracle on the variant).
set -u
mkdir -p gpurun_out
python tools/ab_variants.py run > gpurun_out/ab.txt 2>&1
cat gpurun_out/ab.txt
PG_LIB=ab_build/libpg_spec_dead.so timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/spec_tests.txt 2>&1
tail -3 gpurun_out/spec_tests.txt
PG_LIB=ab_build/libpg_spec_dead.so timeout 200 python tools/gpu_fuzz.py 60 300 11 > gpurun_out/spec_fuzz.txt 2>&1
tail -2 gpurun_out/spec_fuzz.txt
# the pruning experiment on top (-DPG_SPEC_PRUNE=1): same checks
PG_LIB=ab_build/libpg_spec_prune.so timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/prune_tests.txt 2>&1
tail -3 gpurun_out/prune_tests.txt
PG_LIB=ab_build/libpg_spec_prune.so timeout 200 python tools/gpu_fuzz.py 60 300 12 > gpurun_out/prune_fuzz.txt 2>&1
tail -2 gpurun_out/prune_fuzz.txt
# the other workload shapes (configs 3 / 4 share / 5) on the default build and on the speculative one
timeout 400 python tools/bench_configs.py > gpurun_out/configs_base.txt 2>&1
tail -8 gpurun_out/configs_base.txt
PG_LIB=ab_build/libpg_spec_dead.so timeout 400 python tools/bench_configs.py > gpurun_out/configs_spec.txt 2>&1
tail -8 gpurun_out/configs_spec.txt
