python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
for e in UrchinBall LuxoCube; do python bench.py --env $e --no_cpu --steps 2 --warmup 3 2>&1 | tail -1 > gpurun_out/b_env_$e.json; done
for e in Crab CrabCube SpiderCube; do python bench.py --env $e --worlds 65536 --no_cpu --steps 2 --warmup 3 2>&1 | tail -1 > gpurun_out/b_env_$e.json; done
python - <<'PY'
import json
for f in ["bench_final","b_env_UrchinBall","b_env_LuxoCube","b_env_Crab","b_env_CrabCube","b_env_SpiderCube"]:
    j=json.loads(open("gpurun_out/"+f+".json").read().strip().split("\n")[-1]); print(f, round(j["value"]/1e6,3), round(j["e2e"]["value"]/1e6,3), j["config"]["manifold_slot_overflows"])
PY
