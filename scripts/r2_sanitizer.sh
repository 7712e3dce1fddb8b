#!/bin/bash
# compute-sanitizer over smoke() (tiny scene through every kernel of the training step): memcheck, then racecheck
mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|smoke ok|Error|Invalid" gpurun_out/r2s_memcheck.log | head -8 | cut -c1-200
timeout 150 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|smoke ok|hazard|Error" gpurun_out/r2s_racecheck.log | head -8 | cut -c1-200
