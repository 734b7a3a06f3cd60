timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 120 python tools/ba_profile.py 10 1
timeout 120 python tools/ba_profile.py 20 1
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v9.json 2> gpurun_out/bench_v9.err
python -c "
import json
d=json.load(open('gpurun_out/bench_v9.json'))
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'cpu',round(d['cpu_baseline']['value']),'lk us',round(d['roofline']['us_per_launch']))
"
tail -3 gpurun_out/bench_v9.err
