#!/bin/bash
# one GPU: the whole GPU suite, both bench arms as the driver runs them, smoke()
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02w_pytest.log 2>&1; tail -n 6 gpurun_out/r02w_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02w_smoke.log 2>&1; tail -n 2 gpurun_out/r02w_smoke.log
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r02w_ref.json 2> gpurun_out/r02w_ref.err; tail -c 300 gpurun_out/r02w_ref.json; tail -4 gpurun_out/r02w_ref.err
( time timeout 900 python bench.py ) > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err; python tools/benchrow.py default gpurun_out/r02w_bench.json; tail -6 gpurun_out/r02w_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02w_bench.json').read().strip().splitlines()[-1])
print('e2e', d['e2e']['value'], 'roundtrip', d['e2e_roundtrip']['value'], 'frac', d['roofline']['frac'])
for k,v in d['configs'].items(): print(k, f"{v['value']:.4e}", f"{v['ms_per_step']:.4f}", f"e2e {v['e2e']['value']:.3e}", v['roofline']['kernel'][:14], v.get('box_moves',{}).get('steps_per_s'))
PY
