mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $O/gpu_tests.log
SX_HALO_PAR=7 timeout 600 python -m pytest tests -m gpu -q -x -k "conv2dmod_bf16 or selftest or generator_full or suffix" > $O/par7_tests.log 2>&1; echo "par7 tests rc=$?"; tail -2 $O/par7_tests.log
python profiles/bench_generator_only.py --batch 64 --iters 10 --precisions fp32 2>&1 | cut -c1-1500 | tee $O/config4_fp32_new.jsonl
SX_SIMT_64=1 python profiles/bench_generator_only.py --batch 64 --iters 10 --precisions fp32 2>&1 | cut -c1-1500 | tee $O/config4_fp32_old.jsonl
python bench.py --steps 10 --warmup 3 --out $O/bench256.jsonl > $O/bench256.log 2> $O/bench256.err; echo "bench256 rc=$?"
python bench.py --image-size 64 --steps 10 --warmup 3 --out $O/bench64.jsonl > $O/bench64.log 2> $O/bench64.err; echo "bench64 rc=$?"
python - <<'PY'
import json
for f in ('bench256','bench64'):
    b=json.loads(open('gpurun_out/r2f/%s.jsonl'%f).read().splitlines()[-1]); j=b['job']; r=b['roofline']
    print(f, round(b['value']), 'frac', round(r['frac'],4), 'job', round(j['value']), 'wall', round(j['wall_s'],1), 'sweep_ms', round(j['sweep_ms']), 'verify_ms', round(j['verify_ms']), j.get('verify'))
    print('   per layer', {k:round(v) for k,v in r['per_layer_tflops'].items()}, r['kernel_ms_per_step'])
PY
