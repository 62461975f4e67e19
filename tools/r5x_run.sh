set -x
ncu --set full --clock-control none --import-source on --kernel-name regex:'^(sort_kernel|match_kernel|path_kernel|knn_kernel)$' -c 4 -f -o gpurun_out/r5w python tools/profile_target.py 10240 1 stage > gpurun_out/r5w_ncu.log 2>&1
grep PROF gpurun_out/r5w_ncu.log
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:'^skid_fixup_kernel$' -c 1 -f -o gpurun_out/r5w_skidfix python tools/config_bench.py > gpurun_out/r5w_skidfix_ncu.log 2>&1
grep PROF gpurun_out/r5w_skidfix_ncu.log
for tool in memcheck racecheck; do
  echo "== $tool" >> gpurun_out/r5w_sanitize.txt
  timeout 250 compute-sanitizer --tool $tool --print-limit 3 python tools/sanitize_target.py 192 2>&1 | grep -E "outputs identical|ERROR SUMMARY|RACECHECK SUMMARY" | head -4 >> gpurun_out/r5w_sanitize.txt
done
cat gpurun_out/r5w_sanitize.txt
