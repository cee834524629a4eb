#!/bin/bash
# round-2 evidence on one GPU: API conformance, sanitizer, normals A/B, C3 full size, cdf status, ncu captures
mkdir -p gpurun_out
timeout 900 python tools/ref_conformance.py > gpurun_out/r02_ref_conformance.log 2>&1; echo "conformance rc=$?"; head -12 gpurun_out/ref_conformance.txt
timeout 200 python tools/ab_normals.py 65536 6 > gpurun_out/r02_normals_ab.txt 2>&1; echo "ab rc=$?"; tail -3 gpurun_out/r02_normals_ab.txt
timeout 200 python tools/c3_run.py > gpurun_out/r02_c3_run.txt 2>&1; echo "c3 rc=$?"; tail -1 gpurun_out/r02_c3_run.txt
timeout 200 python tools/cdf_bench.py > gpurun_out/r02_cdf_bench.txt 2>&1; echo "cdf rc=$?"; tail -2 gpurun_out/r02_cdf_bench.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mcmc_run_fast --launch-skip 24 --launch-count 1 -o gpurun_out/r02_mcmc_run_fast -f python tools/short_run.py 1048576 27 > gpurun_out/r02_ncu_mcmc.log 2>&1; echo "ncu mcmc rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"next_beta_kernel|weights_kernel|mom_cov_small|mahal_small|binade_hist|med_hist" --launch-skip 160 --launch-count 8 -o gpurun_out/r02_hbm_kernels -f python tools/short_run.py 1048576 27 > gpurun_out/r02_ncu_hbm.log 2>&1; echo "ncu hbm rc=$?"
for tool in memcheck racecheck initcheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_driver.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: exit $?" > gpurun_out/sanitizer_$tool.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize driver ok|Error:|Invalid|hazard" gpurun_out/sanitizer_$tool.log | sort | uniq -c | head -30 >> gpurun_out/sanitizer_$tool.txt
  cat gpurun_out/sanitizer_$tool.txt
done
