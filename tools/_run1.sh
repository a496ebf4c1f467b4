set -x
python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "assembly" > gpurun_out/t_asm.log 2>&1
python tools/stage_breakdown.py 400 9 > gpurun_out/sb_400.log 2>&1
python tools/stage_breakdown.py 50_50 > gpurun_out/sb_50.log 2>&1
CEIT_TILE_CLOCKS=1 python -m ceit_b200.build --force > gpurun_out/clk_build.log 2>&1
CEIT_TILE_CLOCKS=1 python tools/wb_clocks.py > gpurun_out/wb_clocks_50_50.log 2>&1
tail -5 gpurun_out/t_asm.log gpurun_out/wb_clocks_50_50.log gpurun_out/sb_400.log
