set -x
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 50 --warmup 3 > gpurun_out/r02_bench_n8_dom.json 2> gpurun_out/r02_bench_n8_dom.err
tail -n 5 gpurun_out/r02_bench_n8_dom.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02_bench_n8_dom.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"])
for k,v in d["config"]["modes"].items():
    print(" ", k, v if not isinstance(v, dict) else (v.get("ms_per_step"), v.get("stage_ms_per_step"), v.get("unavailable")))
for k,v in d["scaling_configs"].items():
    print(k, v.get("ms_per_step"), v.get("mode"), v.get("error"))
    for kk,vv in (v.get("modes") or {}).items():
        print("   ", kk, vv if not isinstance(vv, dict) else (vv.get("ms_per_step"), vv.get("stage_ms_per_step"), vv.get("unavailable"), vv.get("domain")))
P
