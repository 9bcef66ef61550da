mkdir -p gpurun_out/r2a
python -m pytest tests -m gpu -x -q > gpurun_out/r2a/gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2a/gpu_tests.log
python __graft_entry__.py smoke > gpurun_out/r2a/smoke.log 2>&1; echo "smoke rc=$?"; tail -6 gpurun_out/r2a/smoke.log
timeout 1500 python profiles/topk_parity.py --image-size 64 --latents 256 --arms fp32,bench,bf16g_fp32c,oracle --oracle-max-seconds 500 --out gpurun_out/r2a/topk_parity_64.json --dump gpurun_out/r2a/topk_parity_64.npz > gpurun_out/r2a/topk64.log 2>&1; echo "topk rc=$?"; grep -v "latents," gpurun_out/r2a/topk64.log | tail -12
python bench.py --image-size 64 --steps 10 --warmup 3 --out gpurun_out/r2a/bench64.jsonl > gpurun_out/r2a/bench64.log 2> gpurun_out/r2a/bench64.err; echo "bench64 rc=$?"; cut -c1-600 gpurun_out/r2a/bench64.log
python bench.py --steps 10 --warmup 3 --out gpurun_out/r2a/bench256.jsonl > gpurun_out/r2a/bench256.log 2> gpurun_out/r2a/bench256.err; echo "bench256 rc=$?"; cut -c1-600 gpurun_out/r2a/bench256.log
