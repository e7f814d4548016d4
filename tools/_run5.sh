mkdir -p gpurun_out/r2e
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2e/bench_n2.json 2> gpurun_out/r2e/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2e/ref_n2.json 2> gpurun_out/r2e/ref_n2.err
tail -c 600 gpurun_out/r2e/bench_n2.err
head -c 1500 gpurun_out/r2e/bench_n2.json
