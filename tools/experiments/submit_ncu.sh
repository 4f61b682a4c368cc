# one ncu capture of the thread-per-output submit kernel (4096 mono x 160 frames, 16 -> 48 kHz, 32 taps)
ncu --set full --import-source on --clock-control none -k regex:submit_fused_tp -s 30 -c 1 -f -o gpurun_out/r02_submit_tp_run python tools/stream_calls.py 4096 1 16000 48000 1 160 exact > /dev/null 2>&1
ncu -i gpurun_out/r02_submit_tp_run.ncu-rep --page raw --csv > gpurun_out/r02_submit_tp_run_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_submit_tp_run.ncu-rep --page source --csv > gpurun_out/r02_submit_tp_run_source.csv 2>/dev/null
ls -la gpurun_out/
