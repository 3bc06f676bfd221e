# one-GPU evidence run: GPU test suite (auto path and forced pipeline), headline bench, per-env benches, ncu launch list
OUT=gpurun_out/r02p
mkdir -p $OUT
python -m pytest tests -m gpu -q -rA > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
BLCD_PIPELINE=1 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_pipeline.log 2>&1; tail -2 $OUT/pytest_gpu_pipeline.log
python bench.py --steps 3 --warmup 3 > $OUT/bench_1gpu.json 2> $OUT/bench.err; tail -2 $OUT/bench.err
for e in Dropbox Bounce2 UrchinBall LuxoCube; do w=262144; if [ $e = Bounce2 ]; then w=65536; fi; if [ $e = Dropbox ]; then w=10000; fi; python bench.py --env $e --worlds $w --steps 2 --warmup 3 --no_ncu --no_render --cpu_seconds 6 > $OUT/bench_$e.json 2>> $OUT/bench.err; done
python bench.py --env Bounce2 --worlds 262144 --steps 2 --warmup 3 --no_ncu --no_render --no_cpu > $OUT/bench_Bounce2_262144.json 2>> $OUT/bench.err
python bench.py --env CrabCube --worlds 65536 --steps 2 --warmup 3 --no_ncu --no_render --no_cpu --no_e2e > $OUT/bench_CrabCube.json 2>> $OUT/bench.err
python bench.py --env CrabCube --worlds 131072 --steps 2 --warmup 3 --no_ncu --no_render --no_cpu --no_e2e > $OUT/bench_CrabCube_131072.json 2>> $OUT/bench.err
python bench.py --env SpiderCube --worlds 65536 --steps 2 --warmup 3 --no_ncu --no_render --no_cpu --no_e2e > $OUT/bench_SpiderCube.json 2>> $OUT/bench.err
python bench.py --env SpiderCube --worlds 131072 --steps 2 --warmup 3 --no_ncu --no_render --no_cpu --no_e2e > $OUT/bench_SpiderCube_131072.json 2>> $OUT/bench.err
M=gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,launch__registers_per_thread
BLCD_PIPELINE=1 BLCD_PIPE_RANGES=1 timeout 300 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $OUT/crabcube_pipeline_kernels_ncu.csv python tools/ncu_case.py CrabCube 131072 1 range > $OUT/ncu_crab.log 2>&1
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
  try:
    d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
    print(f.split("/")[-1], round(d["value"] / 1e6, 2), "M; e2e", d["e2e"] and round(d["e2e"]["value"] / 1e6, 2), "cpu", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"] / 1e6, 3), d["config"]["scene"]["block"], d["config"]["scene"].get("pipeline"), d.get("roofline_solver", {}).get("frac"))
  except Exception as e: print(f, "failed", e)
PY
