mkdir -p gpurun_out/r2d
for m in 4 2; do timeout 200 python tools/diag_pcg.py $m 120 160 > gpurun_out/r2d/pcg_mode$m.txt 2>&1; done
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -150 > gpurun_out/r2d/pytest.log
cat gpurun_out/r2d/pcg_mode4.txt
grep -E "passed|failed|^FAILED" gpurun_out/r2d/pytest.log
