#!/bin/bash
# Final validation of a build (run under gpurun, one GPU): every GPU test, smoke(), the default bench line with its CPU
# baseline, and the reference arm.  Outputs under gpurun_out/<tag>_*.
T=${1:-r02final}
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 | cut -c1-300 > $O/${T}_gputests.log
cat $O/${T}_gputests.log
python __graft_entry__.py --smoke > $O/${T}_smoke.log 2>&1
tail -2 $O/${T}_smoke.log
python bench.py > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err
tail -2 $O/${T}_bench_default.err | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
python - <<PY
import json
d = json.load(open("$O/${T}_bench_default.json"))
print(round(d["value"], 2), round(d["ms_per_step"], 2), d["e2e"]["value"], d["gpu_launches"], d["kernel_time_shares"], d["clocks"],
      d["cpu_baseline"]["value"], d["roofline"]["frac"])
r = json.load(open("$O/${T}_bench_reference.json"))
print(r["value"], r["impl"], r["cpu_baseline"]["kind"])
PY
