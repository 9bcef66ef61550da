mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
for v in "default:" "c11a6:SX_HALO_VARIANT=16" "c9a4b6:SX_HALO_VARIANT=32"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python profiles/exp_layers.py --batch 256 --iters 5 --tag $tag 2>&1 | tail -1 | tee -a $O/exp_layers.txt
done
