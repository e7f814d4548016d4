mkdir -p gpurun_out/r2o
python tools/ncu_profile.py capture --out gpurun_out/r2o --kernels 'render_env' --skip 2 --count 4 -- python bench_configs.py --which unfused --iters 2 > gpurun_out/r2o/cap.log 2>&1
python tools/ncu_profile.py summarize gpurun_out/r2o/prof.ncu-rep --out gpurun_out/r2o/sum > /dev/null 2>&1
ncu -i gpurun_out/r2o/prof.ncu-rep --page source --csv --kernel-name regex:'render_env_tma_kernel<0' 2>/dev/null | head -400 > gpurun_out/r2o/src_fwd.csv
rm -f gpurun_out/r2o/prof.ncu-rep
cat gpurun_out/r2o/sum/ncu_summary.txt
