cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "all rc=$?"
tail -3 gpurun_out/t_all.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_h3.json 2> gpurun_out/bench_h3.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_h3.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_h3.csv python bench.py --steps 3 --warmup 3 --eager --no-cpu-baseline > gpurun_out/ncu_h1.log 2>&1; echo "ncu rc=$?"
python profiles/launch_summary.py gpurun_out/launches_h3.csv | grep -v Fill
