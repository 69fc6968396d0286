set +e
OUT=gpurun_out/r2r
mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_linear.py -m gpu -q --timeout 100 > $OUT/pytest_linear.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_linear.log
tail -5 $OUT/pytest_linear.log
timeout 120 python scripts/_mb_attempt.py > $OUT/mb_attempt_accs.log 2>&1; echo "rc=$?" >> $OUT/mb_attempt_accs.log
tail -10 $OUT/mb_attempt_accs.log
