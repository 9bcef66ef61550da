mkdir -p gpurun_out/r3b
timeout 300 python -m pytest tests -m gpu -q -x -k "native_stem" > gpurun_out/r3b/tests.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r3b/tests.log | cut -c1-300
