# What the round-end validation on a B200 box ran (gpurun --timeout 2700 -- bash tools/gpu_round_end.sh): GPU tests, the bench
# line, the ncu launch list and --set full captures behind profiles/r02, the sanitizer logs.  tools/_run5.sh / _run7.sh: the
# 2- and 8-GPU lines.  Reports (.ncu-rep) are summarised on the box and deleted there (gpurun copies back <= 64 MiB).
mkdir -p gpurun_out/r2m
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2m/pytest.log
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m/bench.json 2> gpurun_out/r2m/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/r2m/launches.csv python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2m/bench_under_ncu.log 2>&1
timeout 400 python bench.py --profile gpurun_out/r2m/profile > gpurun_out/r2m/profile.log 2>&1
rm -rf gpurun_out/ncu_bench
find gpurun_out/r2m -name '*.ncu-rep' -delete
timeout 300 python tools/ncu_profile.py capture --out gpurun_out/r2m/ncu_bilateral --kernels 'bg_' --skip 200 --count 60 -- python bench_configs.py --which bilateral --iters 2
python tools/ncu_profile.py summarize gpurun_out/r2m/ncu_bilateral/prof.ncu-rep --out gpurun_out/r2m/ncu_bilateral > /dev/null 2>&1
rm -f gpurun_out/r2m/ncu_bilateral/prof.ncu-rep
timeout 300 compute-sanitizer --tool racecheck --print-limit 10 python tools/sanitize_small.py > gpurun_out/r2m/racecheck.log 2>&1
timeout 300 compute-sanitizer --tool initcheck --print-limit 10 python tools/sanitize_small.py > gpurun_out/r2m/initcheck.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/r2m/pytest.log
tail -2 gpurun_out/r2m/racecheck.log gpurun_out/r2m/initcheck.log
head -c 400 gpurun_out/r2m/bench.json
