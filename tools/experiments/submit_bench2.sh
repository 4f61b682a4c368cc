python tools/stream_calls.py 4096 1 16000 48000 1 160 | cut -c1-170
python tools/stream_calls.py 1024 2 44100 48000 3 512 | cut -c1-170
python tools/stream_calls.py 512 8 96000 48000 2 512 | cut -c1-170
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv python tools/stream_calls.py 4096 1 16000 48000 1 160 2>/dev/null | grep "submit_" | tail -2 | cut -c40-90,300-330
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv python tools/stream_calls.py 1024 2 44100 48000 3 512 2>/dev/null | grep "submit_" | tail -2 | cut -c40-90,300-330
