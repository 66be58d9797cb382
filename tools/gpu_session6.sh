#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu3.log
tail -4 gpurun_out/pytest_gpu3.log
for wl in c2c_f64_2p20 c2c_f64_2p26 batch_f32 r2c_f64_2p24; do
  timeout 600 python bench.py --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$wl.json").read().strip().splitlines()[-1])
    print("$wl", "value", round(d["value"],2), d["unit"], "ms/step", round(d["ms_per_step"],4), "e2e", d["e2e"] and round(d["e2e"]["value"],3), "roof", d["roofline"] and (round(d["roofline"]["achieved"]), round(d["roofline"]["frac"],3), [round(x,4) for x in d["roofline"]["pass_ms"]]), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],4), d.get("plan"))
except Exception as e:
    print("$wl parse failed", e); print(open("gpurun_out/bench_$wl.err").read()[-2000:])
PY
done
timeout 300 python bench.py --impl reference > gpurun_out/bench_reference.json 2>gpurun_out/bench_reference.err; echo "ref exit $?"; cut -c1-400 gpurun_out/bench_reference.json
