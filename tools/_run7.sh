mkdir -p gpurun_out/r2g
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2g/bench_n8.json 2> gpurun_out/r2g/bench_n8.err
tail -c 400 gpurun_out/r2g/bench_n8.err
head -c 400 gpurun_out/r2g/bench_n8.json
nvidia-smi topo -m > gpurun_out/r2g/topo.txt 2>&1
lscpu | head -30 > gpurun_out/r2g/lscpu.txt
