set +e
mkdir -p gpurun_out/r2p
T="tests/test_gpu_linear.py::test_linear_attempt_equals_stage_sequence"
timeout 300 python -m pytest "$T" -m gpu -x -q -k "1000" --timeout 120 > gpurun_out/r2p/pytest_mn_probe.log 2>&1
rc=$?
echo "probe rc=$rc" | tee -a gpurun_out/r2p/pytest_mn_probe.log
if [ $rc -ne 0 ]; then
  tail -30 gpurun_out/r2p/pytest_mn_probe.log | grep -E "assert|Error|k'" | head
  export TDQ_ATTEMPT_SWAP=1
  timeout 300 python -m pytest "$T" -m gpu -x -q -k "1000" --timeout 120 > gpurun_out/r2p/pytest_mn_probe_swap.log 2>&1
  echo "swap probe rc=$?" | tee -a gpurun_out/r2p/pytest_mn_probe_swap.log
fi
timeout 900 python -m pytest tests/test_gpu_linear.py -m gpu -q -k "attempt" --timeout 300 > gpurun_out/r2p/pytest_attempt.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p/pytest_attempt.log
tail -15 gpurun_out/r2p/pytest_attempt.log
timeout 200 python scripts/_mb_attempt.py > gpurun_out/r2p/mb_attempt.log 2>&1; echo "rc=$?" >> gpurun_out/r2p/mb_attempt.log
tail -8 gpurun_out/r2p/mb_attempt.log
MB_METHOD=bosh3 timeout 200 python scripts/_mb_attempt.py > gpurun_out/r2p/mb_attempt_bosh3.log 2>&1
tail -3 gpurun_out/r2p/mb_attempt_bosh3.log
