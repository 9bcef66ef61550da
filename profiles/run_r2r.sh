mkdir -p gpurun_out/r2r
O=gpurun_out/r2r
python -m pytest tests -m gpu -q -x -k "edge_cases or verify_topk_small or another_device" > $O/new_tests.log 2>&1; echo "new tests rc=$?"; tail -15 $O/new_tests.log | cut -c1-250
ncu --set full --clock-control none -k regex:conv_tc -s 28 -c 14 -o $O/fwd python profiles/exp_layers.py --batch 256 --iters 1 > $O/ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"
ncu -i $O/fwd.ncu-rep --page raw --csv > $O/fwd_raw.csv 2>/dev/null; rm -f $O/fwd.ncu-rep; python profiles/ncu_table.py $O/fwd_raw.csv | tail -7
