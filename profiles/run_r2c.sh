mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 $O/gpu_tests.log
python -m pytest tests -m gpu -q -s -k "topk" 2>&1 | grep -E "throughput mode|passed|failed" | head -5
for v in "base:" "ring8:SX_HALO_VARIANT=8" "ring10:SX_HALO_VARIANT=16" "maxco64:SX_HALO_MAX_CO=64" "maxco32:SX_HALO_MAX_CO=32" "rgbquad:SX_RGB_PREV_QUAD=1"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs python profiles/exp_layers.py --batch 256 --iters 5 --tag $tag 2>&1 | tail -1 | tee -a $O/exp_layers.txt
done
python profiles/bench_generator_only.py --batch 64 --iters 20 --out $O/config4.jsonl > $O/config4.log 2>&1; echo "config4 rc=$?"; cut -c1-700 $O/config4.log
ncu --set full --clock-control none -k regex:"rgb_prev|upsample2x_modulate|torgb_kernel|demod_kernel|modulate_kernel" -c 45 -o $O/bw python profiles/exp_layers.py --batch 256 --iters 1 > $O/ncu_bw.log 2>&1; echo "ncu bw rc=$?"
ncu -i $O/bw.ncu-rep --page raw --csv > $O/bw_raw.csv 2>/dev/null; python profiles/ncu_table.py $O/bw_raw.csv | tail -20
ncu --set full --clock-control none -k regex:"conv_tc" -s 28 -c 14 -o $O/cfg4 python profiles/bench_generator_only.py --precisions bf16 --iters 1 --warmup 2 > $O/ncu_cfg4.log 2>&1; echo "ncu cfg4 rc=$?"
ncu -i $O/cfg4.ncu-rep --page raw --csv > $O/cfg4_raw.csv 2>/dev/null; python profiles/ncu_table.py $O/cfg4_raw.csv | tail -16
rm -f $O/bw.ncu-rep $O/cfg4.ncu-rep
python bench.py --image-size 64 --steps 10 --warmup 3 --out $O/bench64.jsonl > $O/bench64.log 2> $O/bench64.err; echo "bench64 rc=$?"
python bench.py --steps 10 --warmup 3 --out $O/bench256.jsonl > $O/bench256.log 2> $O/bench256.err; echo "bench256 rc=$?"
python - <<'PY'
import json
for f in ('bench64','bench256'):
    b=json.loads(open('gpurun_out/r2c/%s.jsonl'%f).read().splitlines()[-1]); j=b['job']
    print(f, round(b['value']), 'job', round(j['value']), 'wall', round(j['wall_s'],1), 'sweep_ms', round(j['sweep_ms']), 'verify_ms', round(j['verify_ms']), j.get('verify'))
PY
timeout 1200 python profiles/topk_parity.py --image-size 256 --latents 64 --arms fp32,bench,verify --out $O/topk_parity_256.json --dump $O/topk_parity_256.npz --dump-arms fp32,bench > $O/topk256.log 2>&1; echo "topk256 rc=$?"; grep -v "latents," $O/topk256.log | tail -8 | cut -c1-700
