mkdir -p gpurun_out/r2f
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -80 > gpurun_out/r2f/pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f/bench.json 2> gpurun_out/r2f/bench.err
timeout 400 python bench.py --profile gpurun_out/r2f/profile > gpurun_out/r2f/profile.log 2>&1
rm -rf gpurun_out/ncu_bench
grep -E "passed|failed|^FAILED" gpurun_out/r2f/pytest.log
head -c 600 gpurun_out/r2f/bench.json
