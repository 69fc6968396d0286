set +e
mkdir -p gpurun_out/r2b
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/r2b/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b/pytest.log
tail -5 gpurun_out/r2b/pytest.log
