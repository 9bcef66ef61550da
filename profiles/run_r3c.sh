mkdir -p gpurun_out/r3c
timeout 300 python -m pytest tests -m gpu -q -x -k "native_stem or classifier or s2d" > gpurun_out/r3c/tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r3c/tests.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-job --no-cpu-baseline > gpurun_out/r3c/bench.jsonl 2> gpurun_out/r3c/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r3c/bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3c/bench.jsonl').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['config'].get('classifier_mode'))
print(d['roofline']['frac'], d.get('hbm_kernels'))
PY
SX_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3c/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-job > gpurun_out/r3c/ncu_bench.log 2>&1; echo "ncu rc=$?"
python profiles/summarize_launches.py gpurun_out/r3c/launches.csv 2>/dev/null | head -30 | cut -c1-200
