mkdir -p gpurun_out/r2j
for v in "" _bg6 _bg8; do
  IRB_LIB=$PWD/inverserenderingofindoorscene_b200/libirb_b200$v.so timeout 200 python tools/diag_bilateral.py sweep1 > gpurun_out/r2j/bilateral$v.txt 2>&1
done
timeout 300 python bench_configs.py --which unfused,trainlight --iters 10 > gpurun_out/r2j/configs.jsonl 2> gpurun_out/r2j/configs.err
cat gpurun_out/r2j/bilateral*.txt
cat gpurun_out/r2j/configs.jsonl | cut -c1-1200
