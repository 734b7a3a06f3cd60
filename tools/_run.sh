timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v11.json 2> gpurun_out/bench_v11.err
python -c "
import json
d=json.load(open('gpurun_out/bench_v11.json'))
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'cpu',round(d['cpu_baseline']['value']),'lk us',round(d['roofline']['us_per_launch']))
"
tail -3 gpurun_out/bench_v11.err
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('K=50: value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']))
"
