mkdir -p gpurun_out
run() { # tag, env...
  TAG=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 --e2e-steps 3 --extras strong > gpurun_out/r02j_$TAG.json 2> gpurun_out/r02j_$TAG.err
  echo "$TAG rc=$?"
  python - <<PY
import json
j=json.loads(open("gpurun_out/r02j_$TAG.json").read().strip().splitlines()[-1])
print("$TAG", j["ms_per_step"], j["e2e"]["ms_per_step"], "strong", j["strong"]["ms_per_step"])
for r in j["per_rank"]: print("   ", r)
PY
}
run default A=1
run threads4 EDLIB_B200_HOST_THREADS=4
run trace EDLIB_B200_TRACE=1
grep "r7\]\|r0\]" gpurun_out/r02j_trace.err | grep -v "filter seed\|plain sweep" | tail -60 | cut -c1-120
nvidia-smi --query-gpu=index,clocks.sm,power.draw,temperature.gpu --format=csv | head -10
