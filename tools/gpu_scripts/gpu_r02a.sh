set -x
mkdir -p gpurun_out/r02a
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a/bench_base.json 2> gpurun_out/r02a/bench_base.err
python tools/kernel_profile.py > gpurun_out/r02a/kernel_profile.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:"nn_search|interframe_kernel|region_reduce" -c 3 -o gpurun_out/r02a/score python tools/ncu_score.py > gpurun_out/r02a/ncu_score.log 2>&1
$NCU -k regex:"voxelize_ex_vec4|devoxelize16|tta_kernel|point_corner" -c 6 -o gpurun_out/r02a/pv python tools/ncu_engine_step.py 1 > gpurun_out/r02a/ncu_pv.log 2>&1
$NCU -k regex:"kmap_query_sym|ks_permute|ks_count|table_build|dsm_|gb_" -c 10 -o gpurun_out/r02a/map python tools/ncu_engine_step.py 1 > gpurun_out/r02a/ncu_map.log 2>&1
$NCU -k regex:"rs_sweep|rs_digit" -c 4 -o gpurun_out/r02a/sort python tools/ncu_engine_step.py 1 > gpurun_out/r02a/ncu_sort.log 2>&1
ls -la gpurun_out/r02a
