mkdir -p gpurun_out
nvidia-smi -L | head -8; nproc; cat /sys/fs/cgroup/cpu.max
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --extras strong > gpurun_out/r02o_scale_n$N.json 2> gpurun_out/r02o_scale_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 --extras strong > gpurun_out/r02o_scale_n$N.json 2> gpurun_out/r02o_scale_n$N.err
  fi
  echo "N=$N rc=$?"; tail -2 gpurun_out/r02o_scale_n$N.err | cut -c1-300
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02o_scale_n$N.json").read().strip().splitlines()[-1])
    print({k:j[k] for k in ["n_gpus","value","ms_per_step","kernel_ms_per_step"]}, j["e2e"]["value"], j["e2e"]["ms_per_step"], "strong", j["strong"]["ms_per_step"], j["strong"]["alignments_per_s"])
except Exception as e: print("parse failed", e)
PY
done
