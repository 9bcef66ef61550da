mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
SX_HALO_PAR=7 timeout 600 python -m pytest tests -m gpu -q -x -k "conv2dmod_bf16 or selftest or generator_full or suffix or sweep_properties or rgb_prefill" > $O/par_tests.log 2>&1; echo "par tests rc=$?"; tail -15 $O/par_tests.log
for v in "base:SX_HALO_PAR=0" "par1:SX_HALO_PAR=1" "par2:SX_HALO_PAR=2" "par4:SX_HALO_PAR=4" "par7:SX_HALO_PAR=7"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python profiles/exp_layers.py --batch 256 --iters 5 --tag $tag 2>&1 | tail -1 | tee -a $O/exp_layers.txt
done
