set +e
OUT=gpurun_out/r2p
mkdir -p $OUT
timeout 300 python scripts/_mb_attempt.py > $OUT/mb_attempt_groups2.log 2>&1; echo "rc=$?" >> $OUT/mb_attempt_groups2.log
tail -20 $OUT/mb_attempt_groups2.log
timeout 600 python -m pytest tests/test_gpu_linear.py -m gpu -q -k "attempt" --timeout 300 > $OUT/pytest_attempt2.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_attempt2.log
tail -5 $OUT/pytest_attempt2.log
