mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
timeout 1500 python profiles/topk_parity.py --image-size 256 --latents 128 --latent-seed 4242 --arms bench,verify,fp32 --out $O/topk_parity_256_job128.json > $O/topk.log 2>&1; echo "rc=$?"; grep -E "^\[" $O/topk.log | cut -c1-420
