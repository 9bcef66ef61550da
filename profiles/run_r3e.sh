mkdir -p gpurun_out/r3e
{
CUDNN=1 python profiles/exp_stem.py
for v in 1 2; do SX_STEM_VARIANT=$v python profiles/exp_stem.py; done
for d in 1 2 4 3 5 6; do SX_STEM_DEBUG=$d python profiles/exp_stem.py; done
} > gpurun_out/r3e/exp_stem.txt 2>&1
cat gpurun_out/r3e/exp_stem.txt | cut -c1-200
