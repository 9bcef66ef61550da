mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
timeout 600 python -m pytest tests -m gpu -q -x -k "generator_full or suffix or sweep_properties or selftest or conv2dmod_bf16 or topk or counterfactual" > $O/tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests.log
for v in "par1:SX_HALO_PAR=1" "par9_8stages:SX_HALO_PAR=9"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python profiles/exp_layers.py --batch 256 --iters 5 --tag $tag 2>&1 | tail -1 | tee -a $O/exp_layers.txt
done
python bench.py --steps 10 --warmup 3 --no-job --no-cpu-baseline --out $O/bench256.jsonl > /dev/null 2> $O/b256.err; echo "bench rc=$?"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2m/bench256.jsonl').read().splitlines()[-1]); r=b['roofline']
print(round(b['value']), 'e2e', round(b['e2e']['value']), 'frac', round(r['frac'],4), {k:round(v) for k,v in r['per_layer_tflops'].items()})
PY
