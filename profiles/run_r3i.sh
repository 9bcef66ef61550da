O=gpurun_out/r3i; mkdir -p $O
for mb in 512 1024; do
timeout 300 python bench.py --image-size 64 --steps 10 --warmup 3 --no-job --no-cpu-baseline --max-batch $mb > $O/bench64_mb$mb.jsonl 2> $O/bench64_mb$mb.err; echo "mb=$mb rc=$?"; tail -1 $O/bench64_mb$mb.err | cut -c1-200
python - <<PY
import json
d=json.loads(open('$O/bench64_mb$mb.jsonl').read().strip().splitlines()[-1])
print($mb, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('classifier_share_of_step'))
PY
done
