mkdir -p gpurun_out/r2w
M=gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 python -m pytest tests/test_render_gpu.py tests/test_lazy_env_gpu.py -m gpu -q 2>&1 | tail -3
timeout 600 ncu --metrics $M --clock-control none -k regex:'render_env' -s 2 -c 4 --csv --log-file gpurun_out/r2w/ncu.csv python bench_configs.py --which unfused --iters 2 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2w/ncu.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][:40]),{})[r[mi]]=r[vi]
for k,m in d.items():
    print(k, ' '.join('%s=%s'%(a.split('.')[0].split('__')[-1],b) for a,b in m.items()))
PY
timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_render_gpu.py -m gpu -q -k "sg_to_env_tile_ring or render_env_tma_ring" 2>&1 | tail -4
