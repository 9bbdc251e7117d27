for opt in "parse.stage_all=-1" "parse.stage_all=0" "parse.stage_all=1" "parse.k12_threads=128" "parse.threads=256" "parse.threads=1024"; do
  python bench.py --config cfg3 --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs --e2e-steps 1 --tune $opt > gpurun_out/cfg3_tmp.json 2>> gpurun_out/cfg3.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/cfg3_tmp.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("$opt: step %.1f us (%.3f of copy peak)  stages %s  serial %.1f us  repeats %s" % (d["ms_per_step"]*1e3, r["pipeline_frac"], {k: round(v*1e3,1) for k,v in r["stage_ms_per_step"].items()}, r["serial_ms_per_step"]*1e3, [round(x*1e3,1) for x in d["repeats"]["ms_per_step"]]))
PY
done
