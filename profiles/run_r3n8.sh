mkdir -p gpurun_out/r3n8
O=gpurun_out/r3n8
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 5 --warmup 3 --out $O/bench256_n8.jsonl > $O/bench_n8.log 2> $O/bench_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r3n8/bench256_n8.jsonl').read().splitlines()[-1]); j=b['job']
print('N=8', round(b['value']), 'e2e', round(b['e2e']['value']), 'job', round(j['value']), 'latents', j['latents_total'], 'wall', round(j['wall_s'],1), 'sweep', round(j['sweep_ms']), 'gather_ms', round(j['gather_ms'],2), 'verify', round(j['verify_ms']), 'agree', j['picks_agree_across_ranks'], j['verify'], j['picks'], j['merged'])
PY
tail -3 $O/bench_n8.err | cut -c1-300
