mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 600 python -m pytest tests -m gpu -q -x -k "conv2dmod_bf16 or selftest or generator_full or suffix or sweep_properties" > $O/par_tests.log 2>&1; echo "default(par1) tests rc=$?"; tail -3 $O/par_tests.log
for v in "par0:SX_HALO_PAR=0" "par1:SX_HALO_PAR=1" "par3:SX_HALO_PAR=3" "par17:SX_HALO_PAR=17" "par5:SX_HALO_PAR=5" "ring4:SX_HALO_VARIANT=8"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python profiles/exp_layers.py --batch 256 --iters 5 --tag $tag 2>&1 | tail -1 | tee -a $O/exp_layers.txt
done
SX_HALO_PAR=23 timeout 600 python -m pytest tests -m gpu -q -x -k "conv2dmod_bf16 or selftest or generator_full or suffix" > $O/par23_tests.log 2>&1; echo "par23 tests rc=$?"; tail -3 $O/par23_tests.log
