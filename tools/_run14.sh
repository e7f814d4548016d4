mkdir -p gpurun_out/r2n
for z in 0 1 2; do
  IRB_RE_SWZ=$z timeout 300 python -m pytest tests/test_render_gpu.py -m gpu -q -k "render_env_tma_ring or golden_parity or other_env_tables" 2>&1 | tail -4 > gpurun_out/r2n/swz$z.log
  echo "swz=$z: $(tail -1 gpurun_out/r2n/swz$z.log)"
done
timeout 300 python bench_configs.py --which unfused --iters 10 > gpurun_out/r2n/unfused.jsonl 2>&1
cat gpurun_out/r2n/unfused.jsonl | cut -c1-700
