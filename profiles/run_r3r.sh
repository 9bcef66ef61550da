O=gpurun_out/r3r; mkdir -p $O
timeout 200 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; echo "rc=$?"; tail -2 $O/tests.log | cut -c1-300
timeout 100 python bench.py --image-size 64 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench64.jsonl 2> $O/bench64.err; echo "bench64 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3r/bench64.jsonl').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
j=d['job']; print(j['wall_s'], j['value'], j['sweep_ms'], j['verify_ms'], j['verify']['verified'])
PY
