O=gpurun_out/r3o; mkdir -p $O
SX_NCU_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches64.csv python bench.py --image-size 64 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-job > $O/ncu_bench.log 2>&1; echo "ncu rc=$?"
python profiles/summarize_launches.py $O/launches64.csv 2>/dev/null | head -40 | cut -c1-200
