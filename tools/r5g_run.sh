set -x
C=$PWD/ft_fsd_path_planning_b200/csrc
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r5g_tests.txt 2>&1; tail -2 gpurun_out/r5g_tests.txt
timeout 100 python tools/resume_probe.py 8192 1e-5 1e-6 > gpurun_out/r5g_probe.txt 2>&1
timeout 100 python tools/resume_probe.py 8192 1e-6 1e-7 >> gpurun_out/r5g_probe.txt 2>&1
echo "== slot 0 only (before the fix)" >> gpurun_out/r5g_probe.txt
FSD_LIBFSDPLAN=$C/ab_slot0only.so timeout 100 python tools/resume_probe.py 8192 1e-5 1e-6 >> gpurun_out/r5g_probe.txt 2>&1
cat gpurun_out/r5g_probe.txt | cut -c1-400 | tail -12
