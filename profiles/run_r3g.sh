O=gpurun_out/r3g; mkdir -p $O
timeout 600 python bench.py --image-size 64 --steps 20 --warmup 5 > $O/bench64.jsonl 2> $O/bench64.err; echo "bench64 rc=$?"; tail -2 $O/bench64.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3g/bench64.jsonl').read().strip().splitlines()[-1])
print(d['metric'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
j=d.get('job') or {}
print({k:(round(v,3) if isinstance(v,float) else v) for k,v in j.items() if k in ('wall_s','value','ratio_to_step_rate','sweep_ms','verify_ms','throughput_mode_picks_equal_exact')})
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_s2d --launch-skip 3 --launch-count 1 -o $O/stem_full python profiles/exp_stem.py > $O/ncu_stem.log 2>&1; echo "ncu rc=$?"
ncu -i $O/stem_full.ncu-rep --page raw --csv > $O/stem_raw.csv 2>/dev/null; python profiles/ncu_table.py $O/stem_raw.csv | cut -c1-250
rm -f $O/stem_full.ncu-rep
