mkdir -p gpurun_out/r2z
O=gpurun_out/r2z
timeout 600 python -m pytest tests -m gpu -q -x -k "generator_full or suffix or selftest or conv2dmod_bf16 or sweep_properties" > $O/tests.log 2>&1; echo "default tests rc=$?"; tail -2 $O/tests.log
SX_HALO_PAR=65 timeout 600 python -m pytest tests -m gpu -q -x -k "generator_full or suffix or selftest or conv2dmod_bf16 or sweep_properties" > $O/tests65.log 2>&1; echo "par65 tests rc=$?"; tail -4 $O/tests65.log | cut -c1-300
for v in "interior:SX_HALO_PAR=1" "c12par_k32:SX_HALO_PAR=65" "c12par_k64:SX_HALO_PAR=5"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python profiles/exp_layers.py --batch 256 --iters 5 --tag $tag 2>&1 | tail -1 | tee -a $O/exp_layers.txt
done
