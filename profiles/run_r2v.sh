mkdir -p gpurun_out/r2v
O=gpurun_out/r2v
timeout 600 python -m pytest tests -m gpu -q -x -k "generator_full or suffix or selftest or conv2dmod_bf16 or sweep_properties" > $O/tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/tests.log
SX_HALO_PAR=7 timeout 600 python -m pytest tests -m gpu -q -x -k "generator_full or suffix or selftest or conv2dmod_bf16" > $O/tests7.log 2>&1; echo "par7 tests rc=$?"; tail -2 $O/tests7.log
for i in 1 2; do python profiles/exp_layers.py --batch 256 --iters 5 --tag hoisted$i 2>&1 | tail -1 | tee -a $O/exp_layers.txt; done
