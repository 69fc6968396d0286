set +e
mkdir -p gpurun_out/r2p
timeout 600 python -m pytest tests/test_gpu_linear.py -m gpu -x -q -k "attempt" --timeout 120 > gpurun_out/r2p/pytest_attempt.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p/pytest_attempt.log
tail -25 gpurun_out/r2p/pytest_attempt.log
timeout 200 python scripts/_mb_attempt.py > gpurun_out/r2p/mb_attempt.log 2>&1; echo "rc=$?" >> gpurun_out/r2p/mb_attempt.log
cat gpurun_out/r2p/mb_attempt.log | tail -12
