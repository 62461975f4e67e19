set -x
C=$PWD/ft_fsd_path_planning_b200/csrc
timeout 110 python -m pytest tests -m gpu -x -q > gpurun_out/r5h_tests.txt 2>&1; tail -2 gpurun_out/r5h_tests.txt
timeout 60 python tools/mode_ab.py --frames 10240 --kind color slotM= prev=FSD_LIBFSDPLAN=$C/ab_prev.so > gpurun_out/r5h_ab.txt 2>&1
cut -c1-150 gpurun_out/r5h_ab.txt
