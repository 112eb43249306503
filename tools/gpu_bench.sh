set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err; echo rc=$?
cat gpurun_out/bench_plain.json
tail -5 gpurun_out/bench_plain.err
