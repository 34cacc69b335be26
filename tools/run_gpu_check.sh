cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_gpu_step.py -q -m gpu -k "peer" 2>&1 | tail -2
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/check_peer_adam.py 2>&1 | grep -E "^rank|Error" | head -4
for i in 1 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 100 --warmup 5 --comm peer 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('peer', d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'])"
done
