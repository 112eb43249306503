# tests + bench + ncu launch list of the config-2 driver (run under gpurun)
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err; echo rc=$?
cat gpurun_out/bench_plain.json
python tools/profile_case.py 64 2 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cur.csv python tools/profile_case.py 64 2 > gpurun_out/ncu_launches.log 2>&1
cat gpurun_out/plain.log
