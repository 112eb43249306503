# one ncu capture of the strip kernel in the probe harness (argument: launch index to capture)
ncu --set full --clock-control none --import-source on -k regex:front_update_strip -s ${1:-1} -c 1 -f -o gpurun_out/probe_strip ./tools/update_probe 32 > gpurun_out/probe_strip.log 2>&1
ncu -i gpurun_out/probe_strip.ncu-rep --page raw --csv > gpurun_out/probe_strip.csv
ncu -i gpurun_out/probe_strip.ncu-rep --page source --csv > gpurun_out/probe_strip_source.csv
