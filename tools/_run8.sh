mkdir -p gpurun_out/r2h
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2h/pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h/bench.json 2> gpurun_out/r2h/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/r2h/launches.csv python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2h/bench_under_ncu.log 2>&1
timeout 400 python bench.py --profile gpurun_out/r2h/profile > gpurun_out/r2h/profile.log 2>&1
rm -rf gpurun_out/ncu_bench
grep -E "passed|failed|^FAILED" gpurun_out/r2h/pytest.log
head -c 500 gpurun_out/r2h/bench.json
