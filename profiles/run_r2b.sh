mkdir -p gpurun_out/r2b
python -m pytest tests -m gpu -x -q > gpurun_out/r2b/gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2b/gpu_tests.log
python -m pytest tests -m gpu -x -q -s -k "topk or bf16_sweep_close or discriminator_filter" 2>&1 | grep -E "throughput mode|bf16 vs fp32|passed|failed" | head
python bench.py --image-size 64 --steps 10 --warmup 3 --out gpurun_out/r2b/bench64.jsonl > gpurun_out/r2b/bench64.log 2> gpurun_out/r2b/bench64.err; echo "bench64 rc=$?"; python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2b/bench64.jsonl').read().splitlines()[-1]); print(b['value'], json.dumps(b['job'])[:1800])
PY
python bench.py --steps 10 --warmup 3 --out gpurun_out/r2b/bench256.jsonl > gpurun_out/r2b/bench256.log 2> gpurun_out/r2b/bench256.err; echo "bench256 rc=$?"; python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2b/bench256.jsonl').read().splitlines()[-1]); print(b['value'], json.dumps(b['job'])[:1800])
PY
timeout 1500 python profiles/topk_parity.py --image-size 64 --latents 256 --arms fp32,bench,verify,oracle --oracle-batch 256 --oracle-max-seconds 900 --out gpurun_out/r2b/topk_parity_64.json > gpurun_out/r2b/topk64.log 2>&1; echo "topk rc=$?"; grep -v "latents," gpurun_out/r2b/topk64.log | tail -12 | cut -c1-900
