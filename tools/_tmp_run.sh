mkdir -p gpurun_out/fin1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/fin1/bench.json 2> gpurun_out/fin1/bench.err; tail -3 gpurun_out/fin1/bench.err; python -c "
import json; d=json.load(open('gpurun_out/fin1/bench.json')); print(d['value'], d['ms_per_step'], d['decode_ms_per_token'], d['first_chunk_ms'], d['roofline']['frac'], d.get('parity'))"
timeout 400 python tools/tune_sweep.py 1000,9032,1032,5016,9064,17032,1255,255255 3 0 2>&1 | tail -9
