mkdir -p gpurun_out/r2k
O=gpurun_out/r2k
python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $O/gpu_tests.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 $O/smoke.log
for v in "tile8:" "tile16:SX_TORGB_TILE=16"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs python profiles/exp_layers.py --batch 256 --iters 5 --tag $tag 2>&1 | tail -1 | tee -a $O/exp_layers.txt
done
python bench.py --steps 10 --warmup 3 --no-job --no-cpu-baseline --max-batch 512 --out $O/bench256_b512.jsonl > /dev/null 2> $O/b512.err; echo "bench b512 rc=$?"
python bench.py --steps 10 --warmup 3 --no-job --no-cpu-baseline --out $O/bench256_b256.jsonl > /dev/null 2> $O/b256.err; echo "bench b256 rc=$?"
python - <<'PY'
import json
for f in ('bench256_b512','bench256_b256'):
    b=json.loads(open('gpurun_out/r2k/%s.jsonl'%f).read().splitlines()[-1]); r=b['roofline']
    print(f, round(b['value']), 'e2e', round(b['e2e']['value']), 'frac', round(r['frac'],4), 'clf share', round(r['classifier_share_of_step'],3), 'hbm', r['hbm_kernels'])
    print('   ', r['kernel_ms_per_step'])
PY
timeout 600 python profiles/topk_parity.py --image-size 64 --latents 32 --classifier mobilenet --arms fp32,bench,verify --out $O/topk_parity_config1_mobilenet.json > $O/topk_cfg1.log 2>&1; echo "config1 rc=$?"; grep -E "^\[" $O/topk_cfg1.log | cut -c1-400
