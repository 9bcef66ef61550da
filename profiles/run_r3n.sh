O=gpurun_out/r3n; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log | cut -c1-300
for cb in 1024 2048; do
SX_CLASSIFY_BATCH=$cb timeout 300 python bench.py --steps 10 --warmup 3 --no-job --no-cpu-baseline > $O/bench_cb$cb.jsonl 2> $O/bench_cb$cb.err; echo "cb=$cb rc=$?"; tail -1 $O/bench_cb$cb.err | cut -c1-200
python - <<PY
import json
d=json.loads(open('$O/bench_cb$cb.jsonl').read().strip().splitlines()[-1])
print($cb, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('classifier_share_of_step'), d['clocks']['sm_mhz'])
PY
done
