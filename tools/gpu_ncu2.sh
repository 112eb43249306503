set -x
nproc
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json
python tools/profile_case.py 64 2 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv python tools/profile_case.py 64 2 > gpurun_out/ncu_launches.log 2>&1
cat gpurun_out/plain.log
python tools/profile_case.py 64 2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:subtree_factor -c 1 -o gpurun_out/prof_subtree_r01 python tools/profile_case.py 64 2 > gpurun_out/ncu_subtree.log 2>&1
tail -n 3 gpurun_out/ncu_subtree.log
ls -la gpurun_out | tail -8
