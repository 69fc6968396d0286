set +e
OUT=gpurun_out/r2t
mkdir -p $OUT
timeout 185 python -m pytest tests -m gpu -q -x --timeout 60 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
