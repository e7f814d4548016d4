mkdir -p gpurun_out/r2q
P=inverserenderingofindoorscene_b200
M=gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread
for v in default pk64 nopk pk16; do
  case $v in
    default) export IRB_LIB=; unset IRB_LIB; unset IRB_RENDER_ENV_VARIANT;;
    plain) unset IRB_LIB; export IRB_RENDER_ENV_VARIANT=plain;;
    *) export IRB_LIB=$PWD/$P/libirb_b200_$v.so; unset IRB_RENDER_ENV_VARIANT;;
  esac
  timeout 300 python -m pytest tests/test_render_gpu.py -m gpu -q -k "render_env_tma_ring or golden_parity or other_env_tables" 2>&1 | tail -1 > gpurun_out/r2q/test_$v.log
  timeout 600 ncu --metrics $M --clock-control none -k regex:render_env -s 2 -c 4 --csv --log-file gpurun_out/r2q/ncu_$v.csv python bench_configs.py --which unfused --iters 2 > /dev/null 2>&1
  echo "== $v: $(cat gpurun_out/r2q/test_$v.log)"
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2q/ncu_$v.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][:40]),{})[r[mi]]=r[vi]
for k,m in d.items():
    print(k, ' '.join('%s=%s'%(a.split('.')[0].split('__')[-1],b) for a,b in m.items()))
PY
done
