for hw in 0 1; do
  echo "== RSB_SUBMIT_HALFWARP=$hw"
  env $( [ "$hw" = 1 ] && echo RSB_SUBMIT_HALFWARP=1 || echo RSB_X=1) python tools/stream_calls.py 4096 1 16000 48000 1 160
  env $( [ "$hw" = 1 ] && echo RSB_SUBMIT_HALFWARP=1 || echo RSB_X=1) python tools/stream_calls.py 1024 2 44100 48000 3 512
  env $( [ "$hw" = 1 ] && echo RSB_SUBMIT_HALFWARP=1 || echo RSB_X=1) python tools/stream_calls.py 512 8 96000 48000 2 512
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv python tools/stream_calls.py 4096 1 16000 48000 1 160 2>/dev/null | grep submit_fused | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv python tools/stream_calls.py 1024 2 44100 48000 3 512 2>/dev/null | grep submit_fused | tail -3
ncu --set full --import-source on --clock-control none -k regex:submit_fused_tp -s 30 -c 1 -o gpurun_out/r02_submit_tp32 -f python tools/stream_calls.py 4096 1 16000 48000 1 160 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:submit_fused_tp -s 30 -c 1 -o gpurun_out/r02_submit_tp128 -f python tools/stream_calls.py 1024 2 44100 48000 3 512 > /dev/null 2>&1
