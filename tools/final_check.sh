cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_r1_final3.json 2> gpurun_out/bench_r1_final3.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_r1_final3.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['config']['launches_per_step'])"
