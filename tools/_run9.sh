mkdir -p gpurun_out/r2i
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2i/pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i/bench.json 2> gpurun_out/r2i/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2i/ref.json 2> gpurun_out/r2i/ref.err
grep -E "passed|failed|^FAILED" gpurun_out/r2i/pytest.log
head -c 300 gpurun_out/r2i/bench.json; echo; cat gpurun_out/r2i/ref.json | head -c 300
