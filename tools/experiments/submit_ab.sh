# A/B of the submit path: build/ab/libold.so against the in-tree library (exact kernel forced so that
# AUTO's tile-path crossover does not hide the thread-per-output kernel)
L=resampler_b200/lib/libresampler_b200.so
mkdir -p build/ab; cp $L build/ab/libnew.so
for v in old new old new; do
  cp build/ab/lib$v.so $L
  echo "== $v"
  python tools/stream_calls.py 4096 1 44100 48000 3 160 exact 2>&1 | tail -1
  python tools/stream_calls.py 1024 2 44100 48000 3 512 exact 2>&1 | tail -1
  python tools/stream_calls.py 4096 1 16000 48000 1 160 exact 2>&1 | tail -1
  python tools/stream_calls.py 512 8 96000 48000 2 512 exact 2>&1 | tail -1
done
cp build/ab/libnew.so $L
python -m pytest tests/test_gpu_parity.py -q -x -k "submit or fused or divergent or async or handoff" 2>&1 | tail -2
