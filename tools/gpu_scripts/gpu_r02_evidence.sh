set -x
OUT=gpurun_out/r02_ev; mkdir -p $OUT
# (1) launch list of the bench command (cold-cache, serialised: shares only) with DRAM bytes per launch
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-lidal > $OUT/bench_under_ncu.json 2> $OUT/bench_under_ncu.err
NCU="ncu --set full --clock-control none --import-source on"
# (2) conv kernel: warm launches of three shapes (tools/ncu_layers.py: 12 launches per shape)
$NCU -k regex:conv_tc -s 10 -c 1 -o $OUT/conv_l0_96 python tools/ncu_layers.py --lex > $OUT/ncu_conv1.log 2>&1
$NCU -k regex:conv_tc -s 46 -c 1 -o $OUT/conv_l2_128 python tools/ncu_layers.py --lex > $OUT/ncu_conv2.log 2>&1
$NCU -k regex:conv_tc -s 58 -c 1 -o $OUT/conv_l3_256 python tools/ncu_layers.py --lex > $OUT/ncu_conv3.log 2>&1
# (3) scoring kernels
$NCU -k regex:"nn_search|interframe_kernel|region_reduce|grid_fill" -s 30 -c 4 -o $OUT/score python tools/ncu_score.py > $OUT/ncu_score.log 2>&1
# (4) engine non-conv kernels of one step
$NCU -k regex:"voxelize_segments|devoxelize16|point_corner|segment_scatter" -c 8 -o $OUT/pv python tools/ncu_engine_step.py 1 > $OUT/ncu_pv.log 2>&1
$NCU -k regex:"kmap_query_sym|ks_permute|ks_count|table_build|dsm_emit|gb_flags" -c 8 -o $OUT/map python tools/ncu_engine_step.py 1 > $OUT/ncu_map.log 2>&1
$NCU -k regex:"rs_sweep|rs_digit" -c 4 -o $OUT/sort python tools/ncu_engine_step.py 1 > $OUT/ncu_sort.log 2>&1
ls -la $OUT
