set -x
python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "persistent" > gpurun_out/t_new.log 2>&1
tail -n 3 gpurun_out/t_new.log
python tools/stage_breakdown.py 50_50 > gpurun_out/sb_50.log 2>&1
python tools/stage_breakdown.py 400 9 > gpurun_out/sb_400.log 2>&1
python tools/stage_breakdown.py 200 5 > gpurun_out/sb_200.log 2>&1
grep woodbury gpurun_out/sb_50.log gpurun_out/sb_400.log gpurun_out/sb_200.log
