mkdir -p gpurun_out/r2x
python tools/ncu_profile.py capture --out gpurun_out/r2x --kernels 'render_env|sg_env' --skip 0 --count 12 -- python bench_configs.py --which unfused --iters 1 > gpurun_out/r2x/cap.log 2>&1
python tools/ncu_profile.py summarize gpurun_out/r2x/prof.ncu-rep --out gpurun_out/r2x/sum > /dev/null 2>&1
rm -f gpurun_out/r2x/prof.ncu-rep
grep -v "^  stalls" gpurun_out/r2x/sum/ncu_summary.txt | cut -c1-100 | sort | uniq -c
timeout 300 python bench_configs.py --which unfused --iters 10 > gpurun_out/r2x/unfused_b4.jsonl 2>/dev/null
timeout 300 python bench_configs.py --which unfused --iters 10 --unfused-batch 16 > gpurun_out/r2x/unfused_b16.jsonl 2>/dev/null
timeout 300 python bench_configs.py --which trainlight --iters 10 > gpurun_out/r2x/trainlight.jsonl 2>/dev/null
cut -c1-1000 gpurun_out/r2x/*.jsonl
