# developer A/B helper: run gpu_check for each library variant given on the command line ("cur" = the built library)
for v in "$@"; do echo "== $v"; if [ "$v" = cur ]; then unset FSD_LIBFSDPLAN; else export FSD_LIBFSDPLAN=$PWD/ft_fsd_path_planning_b200/csrc/ab_$v.so; fi
python tools/gpu_check.py --frames 4096 --seed 3 --colorless --iters 20 2>&1 | tail -2; python tools/gpu_check.py --frames 10240 --iters 20 2>&1 | tail -2; done
