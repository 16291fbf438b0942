#!/bin/bash
# Round-2 GPU measurement suite (run under gpurun): per-config bench lines, launch list, full ncu capture.
TAG=${1:-r02f}
OUT=gpurun_out
python -m pytest tests/test_optim.py tests/test_edge_frames.py tests/test_drop_semantics.py -m gpu -q 2>&1 | tail -3 > $OUT/${TAG}_tests.log
python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_oc20.json 2> $OUT/${TAG}_bench.err
python bench.py --steps 10 --warmup 3 --optimizer torch --no-cpu-baseline > $OUT/${TAG}_bench_oc20_torchopt.json 2>> $OUT/${TAG}_bench.err
for c in qm9 matpes gatav2 global; do
  python bench.py --config $c --steps 5 --warmup 3 > $OUT/${TAG}_bench_$c.json 2>> $OUT/${TAG}_bench.err
done
cat $OUT/${TAG}_tests.log
tail -5 $OUT/${TAG}_bench.err
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/${TAG}_bench_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["value"],2), d["unit"], round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],2), "cpu", (d.get("cpu_baseline") or {}).get("value"), d["kernel_time_shares"])
    except Exception as e: print(f, "ERR", e)
PY
