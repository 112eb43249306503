set -x
timeout 300 python tools/big_case.py 8 2000 4 2000 > gpurun_out/plain_big.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:front_panel_cluster_oc -s 20 -c 1 -f -o gpurun_out/full_front_panel python tools/big_case.py 8 2000 4 2000 > gpurun_out/ncu_front_panel.log 2>&1
ncu -i gpurun_out/full_front_panel.ncu-rep --page raw --csv > gpurun_out/full_front_panel.csv 2>/dev/null
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo rc=$?
python -c "import json; d=json.load(open('gpurun_out/bench_final.json')); print(d['value'], d['e2e']['value'], d['secondary']['e2e_ms'], d['secondary']['kernels_ms_per_step'], d['secondary']['roofline']['frac'])"
