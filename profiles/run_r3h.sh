mkdir -p gpurun_out/r3h
O=gpurun_out/r3h
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py > $O/dist_check.log 2>&1; echo "dist_check rc=$?"; grep -E "rank|Error|error|OK|ok" $O/dist_check.log | tail -8 | cut -c1-250
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 --job-latents 48 --out $O/bench256_n2.jsonl > $O/bench_n2.log 2> $O/bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r3h/bench256_n2.jsonl').read().splitlines()[-1]); j=b['job']
print('N=2', round(b['value']), 'e2e', round(b['e2e']['value']), 'job', round(j['value']), 'wall', round(j['wall_s'],1), 'sweep', round(j['sweep_ms']), 'gather_ms', round(j['gather_ms'],2), 'verify', round(j['verify_ms']), 'agree', j['picks_agree_across_ranks'], j['verify'])
PY
tail -3 $O/bench_n2.err | cut -c1-300
