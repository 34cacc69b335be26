cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tf32.py -q -m gpu -x 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r1_regression.csv python bench.py --workload regression --steps 3 --warmup 3 --eager --no-cpu-baseline > /dev/null 2>&1; echo "ncu rc=$?"
python profiles/launch_summary.py gpurun_out/launches_r1_regression.csv | grep -v "Fill"
timeout 200 python bench.py --workload regression --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('regression', d['dtype'], round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))"
