mkdir -p gpurun_out/r3a
python -m pytest tests -m gpu -q -x -k "extraction" > gpurun_out/r3a/tests.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r3a/tests.log | cut -c1-250
