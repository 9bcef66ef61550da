mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
timeout 600 python -m pytest tests -m gpu -q -x -k "generator_full or suffix or sweep_properties or selftest or conv2dmod_bf16" > $O/tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests.log
SX_HALO_VARIANT=48 SX_HALO_PAR=7 timeout 600 python -m pytest tests -m gpu -q -x -k "generator_full or suffix or selftest or conv2dmod_bf16" > $O/tests_var.log 2>&1; echo "variant tests rc=$?"; tail -3 $O/tests_var.log
for v in "default:" "c10stream:SX_HALO_VARIANT=16" "c8src6:SX_HALO_VARIANT=32" "both:SX_HALO_VARIANT=48"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python profiles/exp_layers.py --batch 256 --iters 5 --tag $tag 2>&1 | tail -1 | tee -a $O/exp_layers.txt
done
