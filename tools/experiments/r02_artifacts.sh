set -x
python tools/measure_traffic.py tensor > gpurun_out/r02_traffic.log 2>&1
python tools/tc_profile.py 30 > gpurun_out/r02_tc_role_cycles.json 2>gpurun_out/r02_tc_profile.err
ncu --set full --import-source on --clock-control none -k regex:conv_tc2 -s 1 -c 1 -f -o gpurun_out/r02_conv_tc2_full python bench.py --seconds 6 --steps 1 --warmup 1 --no-e2e --no-cpu --no-legs > /dev/null 2>&1
ncu -i gpurun_out/r02_conv_tc2_full.ncu-rep --page details --csv > gpurun_out/r02_conv_tc2_full_details.csv 2>/dev/null
ncu -i gpurun_out/r02_conv_tc2_full.ncu-rep --page raw --csv > gpurun_out/r02_conv_tc2_full_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-legs > /dev/null 2>&1
python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>&1
tail -c 1500 gpurun_out/r02_bench_full.json
