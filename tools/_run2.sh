mkdir -p gpurun_out/r2b
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -150 > gpurun_out/r2b/pytest.log
timeout 200 python tools/diag_step.py > gpurun_out/r2b/diag16.txt 2>&1
DIAG_B=2 timeout 200 python tools/diag_step.py > gpurun_out/r2b/diag2.txt 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b/bench.json 2> gpurun_out/r2b/bench.err
IRB_LIB=$PWD/inverserenderingofindoorscene_b200/libirb_b200_base.so timeout 400 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2b/bench_base.json 2> gpurun_out/r2b/bench_base.err
timeout 300 python tools/ncu_profile.py capture --out gpurun_out/r2b/ncu_render --kernels 'render_sg_group_kernel|render_loss|lsreg' --skip 20 --count 6
timeout 300 python tools/ncu_profile.py capture --out gpurun_out/r2b/ncu_unfused --kernels 'sg_env_group_kernel|render_env' --skip 4 --count 4 -- python bench_configs.py --which unfused --iters 2
timeout 300 python tools/ncu_profile.py capture --out gpurun_out/r2b/ncu_bilateral --kernels 'bg_' --skip 200 --count 60 -- python bench_configs.py --which bilateral --iters 2
for d in ncu_render ncu_unfused ncu_bilateral; do
  python tools/ncu_profile.py summarize gpurun_out/r2b/$d/prof.ncu-rep --out gpurun_out/r2b/$d > /dev/null 2>&1
done
rm -f gpurun_out/r2b/ncu_unfused/prof.ncu-rep gpurun_out/r2b/ncu_bilateral/prof.ncu-rep
du -sh gpurun_out
tail -3 gpurun_out/r2b/pytest.log
