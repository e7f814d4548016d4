mkdir -p gpurun_out/r2k
for v in "" _x1 _x2; do
  export IRB_LIB=$PWD/inverserenderingofindoorscene_b200/libirb_b200$v.so
  timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2k/bench$v.json 2> gpurun_out/r2k/bench$v.err
  timeout 300 python bench_configs.py --which unfused,trainlight --iters 10 > gpurun_out/r2k/configs$v.jsonl 2> gpurun_out/r2k/configs$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2k/bench$v.json')); print('$v', d['value'], d['kernel_ms'])
for l in open('gpurun_out/r2k/configs$v.jsonl'):
    r=json.loads(l)
    print('$v', {k:round(v,4) for k,v in r.items() if k.endswith('_ms')})
PY
done
unset IRB_LIB
timeout 200 python tools/diag_bilateral.py sweep1
