#!/bin/bash
# N-GPU run: the 2-GPU parity tests, then the weak-scaling bench at 20 and 200 steps (every gather variant timed
# in the same process), then the cfg5 job.  usage: scripts/multi_gpu_run.sh N tag
N=${1:-2}; TAG=${2:-n$N}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5; fi
for K in 20 200; do
  timeout 600 $RUN bench.py --gpus $N --steps $K --warmup 5 > gpurun_out/scale_${TAG}_k$K.json 2> gpurun_out/scale_${TAG}_k$K.err || tail -20 gpurun_out/scale_${TAG}_k$K.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_${TAG}_k$K.json").read().strip().splitlines()[-1])
    print("N=$N K=$K: %.3f M img/s, %.1f us/step, repeats %s" % (d["value"]/1e6, d["ms_per_step"]*1e3, ["%.1f" % (x*1e3) for x in d["repeats"]["ms_per_step"]]))
    for k,v in d["config"].get("gather_variants",{}).items():
        if k!="note": print("   gather", k, "%.1f us first, %.1f us best" % (v["ms_per_step"]*1e3, v["min_ms_per_step"]*1e3))
    print("   e2e %.0f img/s, h2d %.1f GB/s per rank (copy only: %.1f)" % (d["e2e"]["value"], d["e2e"]["h2d_gbs_per_rank"], d["e2e"]["h2d_only_gbs_per_rank"]))
except Exception as e: print("no result", e)
PY
done
timeout 600 $RUN bench.py --gpus $N --config cfg5 --steps 10 --warmup 3 --no-gather-compare > gpurun_out/cfg5_${TAG}.json 2> gpurun_out/cfg5_${TAG}.err || tail -20 gpurun_out/cfg5_${TAG}.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/cfg5_${TAG}.json").read().strip().splitlines()[-1])
    print("cfg5 N=$N: %.3f M img/s, %.3f ms per pass, checksum %s" % (d["value"]/1e6, d["ms_per_step"], d["config"]["job"]["pose_checksum"]))
except Exception as e: print("no cfg5 result", e)
PY
