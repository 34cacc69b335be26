cd $GRAFT_REPO_ROOT
for comm in nccl peer nccl peer; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 100 --warmup 5 --comm $comm > gpurun_out/bench_n2_$comm.json 2> gpurun_out/bench_n2_$comm.err; echo "rc=$?"
tail -1 gpurun_out/bench_n2_$comm.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$comm', d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['launches_per_step'])"
done
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('single', d['ms_per_step'], d['value'])"
