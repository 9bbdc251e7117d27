#!/bin/bash
# isolated arg-max launch time and whole-step time for each way of cutting the tail of the work items
for t in 0 2 4 8; do
  python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-other-configs --e2e-steps 1 --tune argmax.tail_opt=$t > gpurun_out/tail_$t.json 2>> gpurun_out/tail.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/tail_$t.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("tail_opt=$t step %.1f us  K3 isolated %.1f us (frac %.3f)  K3 stream %.1f us  probe %.0f GB/s" % (d["ms_per_step"]*1e3, r["avg_launch_ms"]*1e3, r["frac"], r["stream_ms_per_launch"]*1e3, r["read_peak_gbs"]))
PY
done
