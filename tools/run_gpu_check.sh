cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_tf32.py tests/test_gpu_head.py -q -m gpu 2>&1 | tail -2
timeout 180 python bench.py --workload wide --samples 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_big5.json 2> gpurun_out/bench_big5.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_big5.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['step_roofline']['achieved_tflops'], d['roofline']['frac'], d['e2e'])"
timeout 300 python bench.py --workload wide --samples 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_big6.json 2> gpurun_out/bench_big6.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_big6.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['step_roofline']['achieved_tflops'], d['roofline']['frac'], d['e2e'])"
