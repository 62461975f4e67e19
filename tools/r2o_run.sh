for g in p2p nccl; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 --gather $g > gpurun_out/r2o_bench8_$g.log 2> gpurun_out/r2o_bench8_$g.err
  echo "== $g rc=$?"; tail -c 400 gpurun_out/r2o_bench8_$g.err; python - <<'P' $g
import json,sys
try:
    l=[x for x in open(f'gpurun_out/r2o_bench8_{sys.argv[1]}.log') if x.startswith('{')][-1]; d=json.loads(l)
    print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], json.dumps(d['comm'])[:1800])
except Exception as e: print('no json', e)
P
done
