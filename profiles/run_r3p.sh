O=gpurun_out/r3p; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "classifier or stem or s2d or fused" > $O/tests.log 2>&1; echo "rc=$?"; tail -3 $O/tests.log | cut -c1-300
timeout 300 python bench.py --image-size 64 --steps 20 --warmup 5 --no-job --no-cpu-baseline > $O/bench64.jsonl 2> $O/bench64.err; echo "bench64 rc=$?"; tail -1 $O/bench64.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3p/bench64.jsonl').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('classifier_share_of_step'), d['clocks']['sm_mhz'], d['config']['classifier_mode'][:80])
PY
