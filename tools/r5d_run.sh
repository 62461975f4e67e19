set -x
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r5d_tests.txt 2>&1; tail -3 gpurun_out/r5d_tests.txt
for kind in color colorless mixed; do
  n=10240; [ $kind = colorless ] && n=10000; [ $kind = mixed ] && n=8192
  timeout 200 python tools/mode_ab.py --frames $n --kind $kind lists= legacy=FSD_PLAN_MODE=285 lists2= legacy2=FSD_PLAN_MODE=285 >> gpurun_out/r5d_ab.txt 2>&1
done
cut -c1-150 gpurun_out/r5d_ab.txt
