python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/r2o_bench_default.json 2> gpurun_out/r2o_bench_default.err
tail -c 6000 gpurun_out/r2o_bench_default.json
