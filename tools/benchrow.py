#!/usr/bin/env python
"""usage: benchrow.py <tag> <json file> : one line of the numbers that matter from a bench.py line"""
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    k = d["kernels"]
    print(f"{tag:26s} value={d['value']:.4e} warm={d['value_l2_warm']:.4e} ms/step={d['ms_per_step']:.4f} pass_ms={k['sweep_pass'].get('ms_per_pass', k['sweep_pass']['ms_per_launch']):.4f} "
          f"cells_ms={k['cell_build']['ms_per_build']:.4f} rng_ms={k['proposals']['ms_per_launch']:.4f} bonds_ms={k['bond_count'].get('ms_per_count', k['bond_count']['ms_per_launch']):.4f} "
          f"e2e={d['e2e']['value']:.3e}" + (f" halo_ms={k['halo']['ms_per_sweep']:.4f}" if "halo" in k else ""))
except Exception as e:
    print(tag, "FAILED", repr(e))
    try:
        print(open(path.replace(".json", ".err")).read()[-1500:])
    except Exception:
        pass
