# developer helper (GPU box with N GPUs): bench.py at N GPUs the way the driver launches it;  tools/scale_run.sh N tag [--gather ...]
N=$1; tag=$2; shift 2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 "$@" > gpurun_out/${tag}_bench${N}.log 2> gpurun_out/${tag}_bench${N}.err
echo "rc=$?"; python - <<P
import json
l=[x for x in open('gpurun_out/${tag}_bench${N}.log') if x.startswith('{')][-1]; d=json.loads(l)
c=d['comm']
print(d['n_gpus'], round(d['value']), d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'c5', d['config5_shard']['ms_per_step'], d['config5_shard']['parity']['flagged'] if d['config5_shard']['parity'] else None)
print(' path', [round(v,3) for v in c['per_rank']['path_kernel_ms']], 'sort', [round(v,3) for v in c['per_rank']['sort_kernel_ms']], 'gather', [round(v,3) for v in c['per_rank']['gather_ms']], c['gathered_buffer_identical_on_all_ranks'])
print(' parity', d['parity'])
P
