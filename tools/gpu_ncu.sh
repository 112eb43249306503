set -x
python tools/profile_case.py 64 2 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python tools/profile_case.py 64 2 > gpurun_out/ncu_launches.log 2>&1
cat gpurun_out/plain.log
python tools/profile_case.py 64 2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:front_update -s 12 -c 2 -o gpurun_out/prof_update_r01 python tools/profile_case.py 64 2 > gpurun_out/ncu_update.log 2>&1
python tools/profile_case.py 64 2 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:front_panel -s 12 -c 2 -o gpurun_out/prof_panel_r01 python tools/profile_case.py 64 2 > gpurun_out/ncu_panel.log 2>&1
tail -n 3 gpurun_out/ncu_update.log; tail -n 3 gpurun_out/ncu_panel.log
ls -la gpurun_out
