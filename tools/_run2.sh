set -x
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/t_multi.log 2>&1
tail -n 15 gpurun_out/t_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/r02_bench_n2_dom.json 2> gpurun_out/r02_bench_n2_dom.err
tail -n 5 gpurun_out/r02_bench_n2_dom.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02_bench_n2_dom.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], json.dumps(d["config"]["modes"])[:1500])
for k,v in d["scaling_configs"].items():
    print(k, v.get("ms_per_step"), v.get("mode"), json.dumps(v.get("modes"))[:1200], v.get("error"))
P
