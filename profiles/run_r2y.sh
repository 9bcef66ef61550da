mkdir -p gpurun_out/r2y
O=gpurun_out/r2y
# conv launch order of a forward: c0..c13; 2 warm-up forwards = 28 conv launches; c12 and c13 of the third forward = skip 40, count 2
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 40 -c 2 -o $O/c12c13 python profiles/exp_layers.py --batch 256 --iters 1 > $O/ncu.log 2>&1; echo "ncu rc=$?"
ncu -i $O/c12c13.ncu-rep --page source --csv > $O/c12c13_source.csv 2>/dev/null
ncu -i $O/c12c13.ncu-rep --page raw --csv > $O/c12c13_raw.csv 2>/dev/null
rm -f $O/c12c13.ncu-rep
python profiles/top_stalls.py $O/c12c13_source.csv 14 | cut -c1-230
python profiles/ncu_table.py $O/c12c13_raw.csv
