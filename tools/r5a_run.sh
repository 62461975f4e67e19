# session-4 A/B (GPU box): matching in the tail of the sort kernel + bins instead of the key / order kernels
set -x
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r5a_tests.txt 2>&1; tail -3 gpurun_out/r5a_tests.txt
for kind in color colorless mixed; do
  n=10240; [ $kind = colorless ] && n=10000; [ $kind = mixed ] && n=8192
  timeout 200 python tools/mode_ab.py --frames $n --kind $kind fused= legacy=FSD_PLAN_MODE=285 fused2= legacy2=FSD_PLAN_MODE=285 >> gpurun_out/r5a_ab.txt 2>&1
done
cat gpurun_out/r5a_ab.txt | cut -c1-150
