# one-GPU evidence run: GPU test suite (auto path and forced pipeline), headline bench, per-env benches, ncu launch list
OUT=gpurun_out/r02p
mkdir -p $OUT
python -m pytest tests -m gpu -q -rA > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
BLCD_PIPELINE=1 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_pipeline.log 2>&1; tail -2 $OUT/pytest_gpu_pipeline.log
python bench.py --steps 3 --warmup 3 > $OUT/bench_1gpu.json 2> $OUT/bench.err; tail -2 $OUT/bench.err
for e in Dropbox Bounce2 UrchinBall LuxoCube; do w=262144; if [ $e = Bounce2 ]; then w=65536; fi; if [ $e = Dropbox ]; then w=10000; fi; python bench.py --env $e --worlds $w --steps 2 --warmup 3 --no_ncu --no_render --cpu_seconds 6 > $OUT/bench_$e.json 2>> $OUT/bench.err; done
python bench.py --env Bounce2 --worlds 262144 --steps 2 --warmup 3 --no_ncu --no_render --no_cpu > $OUT/bench_Bounce2_262144.json 2>> $OUT/bench.err
python bench.py --env CrabCube --worlds 65536 --steps 2 --warmup 2 --no_ncu --no_render --no_cpu --no_e2e > $OUT/bench_CrabCube.json 2>> $OUT/bench.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
  try:
    d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
    print(f.split("/")[-1], round(d["value"] / 1e6, 2), "M; e2e", d["e2e"] and round(d["e2e"]["value"] / 1e6, 2), "cpu", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"] / 1e6, 3), d["config"]["scene"]["block"], d["config"]["scene"].get("pipeline"), d.get("roofline_solver", {}).get("frac"))
  except Exception as e: print(f, "failed", e)
PY
