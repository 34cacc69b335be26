cd $GRAFT_REPO_ROOT
timeout 180 python -m pytest tests/test_gpu_tf32.py -q -m gpu -x -k "batch_resident or large_batch" 2>&1 | tail -3
timeout 180 python bench.py --workload wide --samples 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_big4.json 2> gpurun_out/bench_big4.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_big4.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['step_roofline']['achieved_tflops'], d['roofline']['frac']); [print(k, round(v['us'],1), v['launches_per_step']) for k,v in d['kernels'].items()]"
