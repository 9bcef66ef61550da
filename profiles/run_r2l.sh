mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_tc --csv --log-file $O/subpixel_proxy.csv python profiles/exp_subpixel_projection.py > $O/subpixel_proxy.log 2>&1; echo "proxy rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2l/subpixel_proxy.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows: print(r[4][:70], r[-1], r[-2])
PY
timeout 300 python profiles/topk_parity.py --image-size 64 --latents 32 --classifier mobilenet --arms fp32,bench,verify --out $O/topk_parity_config1_mobilenet.json > $O/topk_cfg1.log 2>&1; echo "config1 rc=$?"; grep -E "^\[" $O/topk_cfg1.log | cut -c1-300
ncu --set full --clock-control none -k regex:conv_tc -s 28 -c 14 -o $O/fwd python profiles/exp_layers.py --batch 256 --iters 1 > $O/ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"
ncu -i $O/fwd.ncu-rep --page raw --csv > $O/fwd_raw.csv 2>/dev/null; rm -f $O/fwd.ncu-rep; python profiles/ncu_table.py $O/fwd_raw.csv | tail -15
SX_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_step.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-job > $O/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"; wc -l $O/launches_step.csv
