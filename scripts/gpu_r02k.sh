# final sanity of the final build: full GPU suite, default bench line, reference arm
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -4
python bench.py > gpurun_out/r02k_bench_c2.json 2> gpurun_out/r02k_bench_c2.err; tail -c 200 gpurun_out/r02k_bench_c2.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02k_bench_reference.json 2> /dev/null
python -c "
import json
d=json.load(open('gpurun_out/r02k_bench_c2.json')); r=json.load(open('gpurun_out/r02k_bench_reference.json'))
print('c2', round(d['value'],1), d['ms_per_step'], 'e2e', round(d['e2e']['value'],1), 'frac', d['roofline']['frac'], 'ref', round(r['value'],1), 'ratio', round(d['value']/r['value'],2), 'e2e ratio', round(d['e2e']['value']/r['value'],2))"
python -c "import __graft_entry__ as g; g.smoke()"
