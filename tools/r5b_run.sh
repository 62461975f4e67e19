# session-4 A/B (GPU box): path keys + bins filed by the match kernel (no key / order kernels) vs FSD_PLAN_MODE bit 8
set -x
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r5b_tests.txt 2>&1; tail -3 gpurun_out/r5b_tests.txt
for kind in color colorless mixed; do
  n=10240; [ $kind = colorless ] && n=10000; [ $kind = mixed ] && n=8192
  timeout 200 python tools/mode_ab.py --frames $n --kind $kind keyed= legacy=FSD_PLAN_MODE=285 keyed2= legacy2=FSD_PLAN_MODE=285 >> gpurun_out/r5b_ab.txt 2>&1
done
cut -c1-150 gpurun_out/r5b_ab.txt
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name regex:'skid_' -c 6 -f -o gpurun_out/r5b_skid python tools/config_bench.py > gpurun_out/r5b_skid_ncu.log 2>&1
ls -la gpurun_out/r5b*
