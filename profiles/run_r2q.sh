mkdir -p gpurun_out/r2q
O=gpurun_out/r2q
python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $O/gpu_tests.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 $O/smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 --out $O/bench256.jsonl > $O/bench256.log 2> $O/bench256.err; echo "bench256 rc=$?"
python bench.py --image-size 64 --steps 20 --warmup 5 --out $O/bench64.jsonl > $O/bench64.log 2> $O/bench64.err; echo "bench64 rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2> $O/bench_ref.err; echo "ref rc=$?"; cut -c1-400 $O/bench_ref.log
python bench.py --precision fp32 --classifier-dtype fp32 --classifier-mode eager --preprocess torch --steps 2 --warmup 1 --no-job --no-cpu-baseline --out $O/bench256_parity_mode.jsonl > /dev/null 2> $O/bpm.err; echo "parity-mode bench rc=$?"
python - <<'PY'
import json
for f in ('bench256','bench64','bench256_parity_mode'):
    b=json.loads(open('gpurun_out/r2q/%s.jsonl'%f).read().splitlines()[-1]); j=b.get('job'); r=b['roofline']
    print(f, round(b['value']), 'e2e', round(b['e2e']['value']), 'frac', round(r['frac'],4), 'clf', round(r['classifier_share_of_step'],3), 'hbm', round(r['hbm_kernels']['frac'],3), round(r['hbm_kernels']['share_of_step'],4), 'cpu', b['cpu_baseline'] and round(b['cpu_baseline']['value'],1), b['clocks'])
    if j: print('   job', round(j['value']), 'wall', round(j['wall_s'],1), 'sweep', round(j['sweep_ms']), 'verify', round(j['verify_ms']), j['verify']['candidates'], j['verify']['verified'], 'fast==exact', j['throughput_mode_picks_equal_exact'])
    print('   ', {k:round(v) for k,v in r['per_layer_tflops'].items()})
PY
