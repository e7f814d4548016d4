mkdir -p gpurun_out/r2l
for v in "" _b3; do
  export IRB_LIB=$PWD/inverserenderingofindoorscene_b200/libirb_b200$v.so
  timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2l/bench$v.json 2> gpurun_out/r2l/bench$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2l/bench$v.json')); print('$v', d['value'], d['kernel_ms'])
PY
done
