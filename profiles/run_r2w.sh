mkdir -p gpurun_out/r2w
O=gpurun_out/r2w
for v in "default:" "c12sets3:SX_HALO_VARIANT=16" "c13sets3:SX_HALO_VARIANT=32"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python profiles/exp_layers.py --batch 256 --iters 5 --tag $tag 2>&1 | tail -1 | tee -a $O/exp_layers.txt
done
