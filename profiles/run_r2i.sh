mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
timeout 900 python -m pytest tests -m gpu -q -x -k "training or train_step" > $O/train_tests.log 2>&1; echo "train tests rc=$?"; tail -25 $O/train_tests.log | cut -c1-300
timeout 600 python profiles/bench_train_step.py --image-size 256 --batch 16 --steps 2 --warmup 2 --precision bf16 --profile > $O/train_profile.txt 2>&1; grep -v "^-" $O/train_profile.txt | cut -c1-200 | head -45
