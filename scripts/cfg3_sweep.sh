#!/bin/bash
# dense crowds (cfg3): the three-kernel chain with persistent grids vs one CTA per image, decode+NMS CTA sizes, and the
# forced two-kernel chain
for opt in "parse.persist=0" "parse.persist=1" "parse.persist=2" "parse.persist=1 --tune parse.k12_threads=512" "parse.persist=2 --tune parse.k12_threads=512" "parse.persist=3 --tune parse.k12_threads=128" "parse.fused=1"; do
  python bench.py --config cfg3 --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs --e2e-steps 1 --tune $opt > gpurun_out/cfg3_tmp.json 2>> gpurun_out/cfg3.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/cfg3_tmp.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("$opt: step %.1f us (%.3f of copy peak)  stages %s  serial %.1f us  launches %d" % (d["ms_per_step"]*1e3, r["pipeline_frac"], {k: round(v*1e3,1) for k,v in r["stage_ms_per_step"].items()}, r["serial_ms_per_step"]*1e3, d["gpu_launches"]//d["steps"]))
PY
done
