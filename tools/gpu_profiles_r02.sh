# Round-2 profile captures (run under gpurun): each profiled command first runs plain (exit 0), then the launch list of
# the bench command, then one --set full capture per hot kernel.  Raw CSV pages come back in gpurun_out/.
set -x
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/bench_short.json 2>/dev/null &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/ncu_bench.log 2>&1
python tools/profile_case.py 64 2 > gpurun_out/plain.log 2>&1 || exit 1
for k in subtree_factor subtree_forward subtree_backward subtree_leaf_kernel front_small; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/full_$k python tools/profile_case.py 64 2 > gpurun_out/ncu_$k.log 2>&1
  ncu -i gpurun_out/full_$k.ncu-rep --page raw --csv > gpurun_out/full_$k.csv 2>/dev/null
  rm -f gpurun_out/full_$k.ncu-rep
done
timeout 300 python tools/big_case.py 32 2000 4 2000 > gpurun_out/plain_big.log 2>&1 || exit 1
for k in front_update_strip front_panel_cluster_oc front_forward_cluster front_backward_cluster; do
  skip=20; case $k in front_forward_cluster|front_backward_cluster) skip=1;; esac
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/full_$k python tools/big_case.py 32 2000 4 2000 > gpurun_out/ncu_$k.log 2>&1
  ncu -i gpurun_out/full_$k.ncu-rep --page raw --csv > gpurun_out/full_$k.csv 2>/dev/null
  rm -f gpurun_out/full_$k.ncu-rep
done
python tools/ipm_vec_probe.py > gpurun_out/plain_ipm.log 2>&1 || exit 1
for k in ipm_ftb ipm_step; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/full_$k python tools/ipm_vec_probe.py > gpurun_out/ncu_$k.log 2>&1
  ncu -i gpurun_out/full_$k.ncu-rep --page raw --csv > gpurun_out/full_$k.csv 2>/dev/null
  rm -f gpurun_out/full_$k.ncu-rep
done
ls -la gpurun_out | tail -30
