# developer A/B (GPU box): upper bound of work-ordered scheduling -- frames ordered by their MEASURED path time
C=$PWD/ft_fsd_path_planning_b200/csrc
FSD_LIBFSDPLAN=$C/ab_probe.so FSD_AB_DUMP_GRID=/tmp/order_asc.npy python tools/mode_ab.py probe=
python - <<'P'
import numpy as np
o=np.load('/tmp/order_asc.npy'); np.save('/tmp/order_desc.npy', o[::-1].copy())
rng=np.random.default_rng(0); np.save('/tmp/order_rand.npy', rng.permutation(len(o)))
# heavy first only for the heaviest 10 %, the rest in generator order
k=len(o)//10; heavy=o[::-1][:k]; rest=np.setdiff1d(np.arange(len(o)), heavy, assume_unique=False)
np.save('/tmp/order_top10.npy', np.concatenate([heavy, rest]))
P
python tools/mode_ab.py base= desc=FSD_AB_ORDER=/tmp/order_desc.npy asc=FSD_AB_ORDER=/tmp/order_asc.npy rand=FSD_AB_ORDER=/tmp/order_rand.npy top10=FSD_AB_ORDER=/tmp/order_top10.npy
