#!/bin/bash
# compute-sanitizer record of smoke() (4000 atoms, 25 MD steps, one rebuild): memcheck + racecheck + initcheck-free summary
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok" gpurun_out/r2_sanitizer_$tool.log
done
