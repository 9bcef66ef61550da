mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
timeout 900 python profiles/bench_train_step.py --image-size 256 --batch 32 --steps 8 --warmup 6 --precision bf16 --out $O/train.jsonl 2>&1 | tail -1 | cut -c1-1200
timeout 900 python profiles/bench_train_step.py --image-size 256 --batch 32 --steps 8 --warmup 6 --precision bf16 --channels-last --out $O/train.jsonl 2>&1 | tail -1 | cut -c1-1200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 profiles/bench_train_step.py --gpus 2 --image-size 256 --batch 32 --steps 8 --warmup 6 --precision bf16 --out $O/train.jsonl 2>&1 | tail -1 | cut -c1-1200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tests/dist_train_check.py 2>&1 | tail -3 | cut -c1-300
