mkdir -p gpurun_out/r2u
O=gpurun_out/r2u
python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $O/gpu_tests.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
python profiles/bench_generator_only.py --batch 64 --iters 10 --precisions fp32 2>&1 | cut -c1-1300 | tee $O/config4_fp32_ffma2.jsonl
python bench.py --steps 5 --warmup 3 --job-latents 32 --out $O/bench256.jsonl > /dev/null 2> $O/b.err; echo "bench rc=$?"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2u/bench256.jsonl').read().splitlines()[-1]); j=b['job']
print(round(b['value']), 'e2e', round(b['e2e']['value']), 'frac', round(b['roofline']['frac'],4), 'job', round(j['value']), 'verify_ms', round(j['verify_ms']), j['verify']['verified'], 'launches', b['gpu_launches'])
PY
