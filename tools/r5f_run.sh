set -x
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r5f_tests.txt 2>&1; tail -2 gpurun_out/r5f_tests.txt
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name regex:'path_kernel' --launch-skip 2 -c 1 -f -o gpurun_out/r5f python tools/profile_target.py 10240 3 stage > gpurun_out/r5f_ncu.log 2>&1
tail -3 gpurun_out/r5f_ncu.log
timeout 100 python tools/mode_ab.py --frames 10240 --kind color cur= >> gpurun_out/r5f_ab.txt 2>&1; cut -c1-150 gpurun_out/r5f_ab.txt
