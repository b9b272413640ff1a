mkdir -p gpurun_out/r02g
for cf in 1.05 1.5 2.0 2.5 3.0; do LIDAL_CELL_FACTOR=$cf timeout 300 python tools/time_score.py 2>&1 | tail -1; done > gpurun_out/r02g/cell_sweep.txt
cat gpurun_out/r02g/cell_sweep.txt
ncu --set full --clock-control none --import-source on -k regex:"nn_search|interframe_kernel" -c 2 -o gpurun_out/r02g/score python tools/ncu_score.py > gpurun_out/r02g/ncu_score.log 2>&1
ls -la gpurun_out/r02g
