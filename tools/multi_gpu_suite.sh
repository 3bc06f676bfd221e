# 8-GPU evidence for BASELINE configs 3 (strong scaling as stated, and weak), 5 (world-count sweep) and 4 (dataset collection)
N=${1:-8}
OUT=gpurun_out/r02n
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_${N}gpu_strong.json 2> $OUT/bench_strong.err; tail -2 $OUT/bench_strong.err
timeout 300 $TR --master-port 29522 bench.py --gpus $N --steps 3 --warmup 3 --scaling weak --no_cpu --no_ncu --no_render > $OUT/bench_${N}gpu_weak.json 2> $OUT/bench_weak.err; tail -2 $OUT/bench_weak.err
timeout 300 $TR --master-port 29523 bench.py --gpus $N --sweep --steps 2 --warmup 1 --no_cpu > $OUT/sweep_${N}gpu.json 2> $OUT/sweep.err; tail -2 $OUT/sweep.err
( time timeout 300 $TR --master-port 29524 -m boxlcd_b200.collect --env=LuxoCube --barrels=110 --logdir=/tmp/luxo ) > $OUT/collect_luxocube_${N}gpu.log 2>&1; ls /tmp/luxo/train | wc -l >> $OUT/collect_luxocube_${N}gpu.log; du -sh /tmp/luxo >> $OUT/collect_luxocube_${N}gpu.log
( time timeout 300 $TR --master-port 29525 -m boxlcd_b200.collect --env=Urchin --collect_n=16000 ) > $OUT/collect_npz_${N}gpu.log 2>&1; ls -la rollouts >> $OUT/collect_npz_${N}gpu.log
python - <<PY
import json
for f in ["bench_${N}gpu_strong", "bench_${N}gpu_weak"]:
  try:
    d = json.load(open("$OUT/" + f + ".json")); print(f, d["scaling"], d["n_gpus"], round(d["value"] / 1e6, 2), "M env-steps/s; e2e", d["e2e"] and round(d["e2e"]["value"] / 1e6, 2), d["config"]["worlds_per_gpu"], d["config"]["scene"]["block"], d["config"]["scene"].get("pipeline"))
  except Exception as e: print(f, "failed", e)
try:
  d = json.load(open("$OUT/sweep_${N}gpu.json")); print([(r["total_worlds"], round(r["env_steps_per_s"] / 1e6, 2)) for r in d["sweep"]])
except Exception as e: print("sweep failed", e)
PY
tail -4 $OUT/collect_luxocube_${N}gpu.log; tail -5 $OUT/collect_npz_${N}gpu.log
