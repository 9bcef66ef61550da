O=gpurun_out/r3k; mkdir -p $O
timeout 900 python profiles/topk_parity.py --image-size 64 --latents 256 --arms fp32,bench,verify --max-batch 1024 --out $O/topk_parity_64_native_stem.json > $O/topk.log 2>&1; echo "rc=$?"; tail -25 $O/topk.log | cut -c1-300
