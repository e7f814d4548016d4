mkdir -p gpurun_out/r2c
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -200 > gpurun_out/r2c/pytest.log
timeout 600 python tools/diag_bilateral.py > gpurun_out/r2c/diag_bilateral.txt 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c/bench.json 2> gpurun_out/r2c/bench.err
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_small.py > gpurun_out/r2c/racecheck.log 2>&1
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 600 compute-sanitizer --tool initcheck python tools/sanitize_small.py > gpurun_out/r2c/initcheck.log 2>&1
timeout 300 python tools/ncu_profile.py capture --out gpurun_out/r2c/ncu_unfused --kernels 'sg_env_group_kernel|render_env' --skip 0 --count 16 -- python bench_configs.py --which unfused --iters 1
python tools/ncu_profile.py summarize gpurun_out/r2c/ncu_unfused/prof.ncu-rep --out gpurun_out/r2c/ncu_unfused > /dev/null 2>&1
rm -f gpurun_out/r2c/ncu_unfused/prof.ncu-rep
tail -c 3000 gpurun_out/r2c/racecheck.log > gpurun_out/r2c/racecheck_tail.log
tail -c 3000 gpurun_out/r2c/initcheck.log > gpurun_out/r2c/initcheck_tail.log
du -sh gpurun_out
tail -3 gpurun_out/r2c/pytest.log
