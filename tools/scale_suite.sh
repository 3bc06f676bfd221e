# strong-scaling bench lines at N = 2 and N = 4 (BASELINE config 3: 262 144 Urchin worlds in total)
OUT=gpurun_out/r02s
mkdir -p $OUT
for N in 2 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 5 --warmup 3 --no_cpu --no_ncu --no_render 2> $OUT/bench_$N.err | grep "^{" > $OUT/bench_${N}gpu_strong.json
  python - <<PY
import json
d = json.loads(open("$OUT/bench_${N}gpu_strong.json").read())
print("N=$N", d["scaling"], round(d["value"] / 1e6, 2), "M env-steps/s; e2e", round(d["e2e"]["value"] / 1e6, 2), "worlds/GPU", d["config"]["worlds_per_gpu"], "block", d["config"]["scene"]["block"], "pipeline", d["config"]["scene"]["pipeline"])
PY
done
