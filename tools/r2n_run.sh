python -m pytest tests/test_gpu_parity.py -x -q -k "fused_gather or plan_pinned" > gpurun_out/r2n_tests.txt 2>&1; tail -3 gpurun_out/r2n_tests.txt
nvidia-smi topo -m > gpurun_out/r2n_topo.txt 2>&1
for g in p2p p2p-unicast nccl; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --gather $g > gpurun_out/r2n_bench2_$g.log 2> gpurun_out/r2n_bench2_$g.err
  echo "== $g rc=$?"; tail -c 600 gpurun_out/r2n_bench2_$g.err; python - <<'P' $g
import json,sys
try:
    l=[x for x in open(f'gpurun_out/r2n_bench2_{sys.argv[1]}.log') if x.startswith('{')][-1]; d=json.loads(l)
    print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], json.dumps(d['comm'])[:900])
except Exception as e: print('no json', e)
P
done
