cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r1_final2.json 2> gpurun_out/bench_r1_final2.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_r1_final2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches'], d['clocks'])"
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 2>/dev/null | tail -1 | cut -c1-300
