run() { name="$1"; shift; timeout 200 python bench.py "$@" --no-e2e --no-cpu --steps 5 --warmup 3 > gpurun_out/tmp.log 2>gpurun_out/tmp.err; python - "$name" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/tmp.log").read().strip().splitlines()[-1]); print(sys.argv[1], d["value"], d["ms_per_step"], d["config"]["kernel"])
except Exception as e:
    print(sys.argv[1], "FAILED", open("gpurun_out/tmp.err").read()[-400:])
PY
}
for k in tensor fast; do
run "mono16->48k_32taps_4096x30s_$k" --kernel $k --channels 1 --in-hz 16000 --out-hz 48000 --latency 1 --streams 4096 --seconds 30 --call-frames 160
run "stereo48->44.1_128taps_$k" --kernel $k --in-hz 48000 --out-hz 44100
run "mono44.1->48_128taps_2048_$k" --kernel $k --channels 1 --streams 2048
run "8ch96->48_64taps_512x20s_$k" --kernel $k --channels 8 --in-hz 96000 --out-hz 48000 --latency 2 --streams 512 --seconds 20
done
