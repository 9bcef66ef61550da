mkdir -p gpurun_out/r3f
{ python profiles/exp_stem.py; SX_STEM_VARIANT=1 python profiles/exp_stem.py; SX_STEM_VARIANT=2 python profiles/exp_stem.py; } > gpurun_out/r3f/exp_stem.txt 2>&1; cat gpurun_out/r3f/exp_stem.txt | cut -c1-160
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r3f/tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r3f/tests.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r3f/bench.jsonl 2> gpurun_out/r3f/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r3f/bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3f/bench.jsonl').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])
print(json.dumps(d.get('job'))[:600])
PY
