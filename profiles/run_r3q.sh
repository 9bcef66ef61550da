O=gpurun_out/r3q; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; echo "rc=$?"; tail -2 $O/tests.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --job-latents 32 > $O/bench.jsonl 2> $O/bench.err; echo "bench rc=$?"; tail -1 $O/bench.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3q/bench.jsonl').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('classifier_share_of_step'), d['clocks']['sm_mhz'])
j=d['job']; print(j['wall_s'], j['value'], j['verify']['verified'], j['picks_agree_across_ranks'])
PY
