O=gpurun_out/r3m; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; echo "rc=$?"; tail -3 $O/tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.jsonl 2> $O/bench.err; echo "bench rc=$?"; tail -2 $O/bench.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3m/bench.jsonl').read().strip().splitlines()[-1])
print(d['metric'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['achieved'], d['clocks'])
j=d.get('job') or {}
print({k:(round(v,3) if isinstance(v,float) else v) for k,v in j.items() if k in ('wall_s','value','ratio_to_step_rate','sweep_ms','verify_ms')}, (j.get('verify') or {}).get('verified'))
print(d['cpu_baseline'])
PY
