set -x
python -m pytest tests -m gpu -q > gpurun_out/r5w_tests.txt 2>&1; tail -2 gpurun_out/r5w_tests.txt
python bench.py --impl reference > gpurun_out/r5w_bench_ref.log 2> gpurun_out/r5w_bench_ref.err
python bench.py > gpurun_out/r5w_bench.log 2> gpurun_out/r5w_bench.err
python tools/config_bench.py > gpurun_out/r5w_configs.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r5w_launches.csv python bench.py --steps 2 --warmup 3 --no-reference-numba > gpurun_out/r5w_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:'sort_kernel|match_kernel|path_kernel|knn_kernel' -c 4 -f -o gpurun_out/r5w python tools/profile_target.py 10240 1 stage > gpurun_out/r5w_ncu.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:'skid_(track|step|fixup)_kernel' -c 3 -f -o gpurun_out/r5w_skid python tools/config_bench.py > gpurun_out/r5w_skid_ncu.log 2>&1
ls -la gpurun_out/r5w*
