mkdir -p gpurun_out/r2y
P=inverserenderingofindoorscene_b200
M=gpu__time_duration.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 python -m pytest tests/test_bilateral_gpu.py -m gpu -q 2>&1 | tail -2
for v in default nohint; do
  if [ $v = default ]; then unset IRB_LIB; else export IRB_LIB=$PWD/$P/libirb_b200_$v.so; fi
  timeout 300 python bench_configs.py --which bilateral --iters 10 > gpurun_out/r2y/bil_$v.jsonl 2>/dev/null
  echo "== $v: $(cut -c1-400 gpurun_out/r2y/bil_$v.jsonl)"
  timeout 600 ncu --metrics $M --clock-control none -k regex:'bg_(q|p|xr)_kernel' -s 30 -c 9 --csv --log-file gpurun_out/r2y/ncu_$v.csv python bench_configs.py --which bilateral --iters 1 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2y/ncu_$v.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][:22]),{})[r[mi]]=r[vi]
for k,m in list(d.items())[:6]:
    print(k, ' '.join('%s=%s'%(a.split('.')[0].split('__')[-1],b) for a,b in m.items()))
PY
done
