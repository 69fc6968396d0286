set +e
mkdir -p gpurun_out/r2c
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2c/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c/pytest.log
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2c/pytest.log | tail -40
