mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
timeout 900 python -m pytest tests -m gpu -q -x -k "training or train_step or backward or grad" > $O/train_tests.log 2>&1; echo "train tests rc=$?"; tail -25 $O/train_tests.log | cut -c1-300
timeout 600 python profiles/bench_train_step.py --image-size 64 --batch 32 --steps 4 --warmup 2 --precision fp32 --out $O/train.jsonl 2>&1 | tail -3 | cut -c1-900
timeout 600 python profiles/bench_train_step.py --image-size 256 --batch 16 --steps 4 --warmup 2 --precision bf16 --out $O/train.jsonl 2>&1 | tail -3 | cut -c1-900
