cd $GRAFT_REPO_ROOT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2c.json 2> gpurun_out/bench_n2c.err; echo "rc=$?"
tail -1 gpurun_out/bench_n2c.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload wide --samples 4 --steps 3 --warmup 3 > gpurun_out/bench_n2w.json 2> gpurun_out/bench_n2w.err; echo "rc=$?"
tail -1 gpurun_out/bench_n2w.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['step_roofline']['achieved_tflops'], d['roofline'])"
timeout 300 python bench.py --workload wide --samples 4 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['step_roofline']['achieved_tflops'], d['roofline'])"
