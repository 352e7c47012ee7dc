mkdir -p gpurun_out/v5g
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/timeline.py --window 4 --out gpurun_out/v5g/tl.json 2>&1 | tail -17
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/v5g/bench.json 2> gpurun_out/v5g/bench.err; python -c "
import json; d=json.load(open('gpurun_out/v5g/bench.json')); print(d['value'], d['decode_ms_per_token'], d['first_chunk_ms'], d['roofline']['frac'])"
