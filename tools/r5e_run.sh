# session-4 A/B (GPU box): PathMachine in shared memory (default build) vs on the stack (ab_pmstack.so)
# (as run at build r2_zb, where -DFSD_PM_SHARED=0 put the machine back on the stack; since r2_zd it is a member of the frame slot)
set -x
C=$PWD/ft_fsd_path_planning_b200/csrc
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r5e_tests.txt 2>&1; tail -3 gpurun_out/r5e_tests.txt
for kind in color mixed; do
  n=10240; [ $kind = mixed ] && n=8192
  timeout 200 python tools/mode_ab.py --frames $n --kind $kind shared= stack=FSD_LIBFSDPLAN=$C/ab_pmstack.so shared2= stack2=FSD_LIBFSDPLAN=$C/ab_pmstack.so >> gpurun_out/r5e_ab.txt 2>&1
done
cut -c1-150 gpurun_out/r5e_ab.txt
timeout 120 ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,sass__inst_executed_local_loads,sass__inst_executed_local_stores,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,l1tex__m_l1tex2xbar_write_bytes.sum --clock-control none --kernel-name regex:'path_kernel' --launch-skip 2 -c 1 python tools/profile_target.py 10240 3 stage > gpurun_out/r5e_ncu_shared.txt 2>&1
FSD_LIBFSDPLAN=$C/ab_pmstack.so timeout 120 ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,sass__inst_executed_local_loads,sass__inst_executed_local_stores,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,l1tex__m_l1tex2xbar_write_bytes.sum --clock-control none --kernel-name regex:'path_kernel' --launch-skip 2 -c 1 python tools/profile_target.py 10240 3 stage > gpurun_out/r5e_ncu_stack.txt 2>&1
grep -h "dram__\|local\|issue_active\|duration\|xbar" gpurun_out/r5e_ncu_shared.txt gpurun_out/r5e_ncu_stack.txt
for tool in racecheck memcheck; do
  echo "== $tool" >> gpurun_out/r5e_sanitize.txt
  timeout 200 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_target.py 192 2>&1 | grep -E "outputs identical|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|Hazard" | head -12 >> gpurun_out/r5e_sanitize.txt
done
cat gpurun_out/r5e_sanitize.txt
