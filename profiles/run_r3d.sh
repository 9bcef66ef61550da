mkdir -p gpurun_out/r3d
timeout 600 python bench.py --steps 10 --warmup 3 --no-job --no-cpu-baseline > gpurun_out/r3d/bench.jsonl 2> gpurun_out/r3d/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r3d/bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3d/bench.jsonl').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['config'].get('classifier_mode'))
print(d['roofline']['frac'])
PY
SX_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3d/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-job > gpurun_out/r3d/ncu_bench.log 2>&1; echo "ncu rc=$?"
python profiles/summarize_launches.py gpurun_out/r3d/launches.csv 2>/dev/null | head -24 | cut -c1-160
