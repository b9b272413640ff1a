mkdir -p gpurun_out/r02s
NCU="ncu --set full --clock-control none --import-source on"
# ncu_layers.py: 6 layer shapes x 12 launches each; launch #11 (0-based 10) is a warm 96->96 @ level 0, #23 a warm 32->32, #47 128->128 @ L2
$NCU -k regex:conv_tc -s 10 -c 1 -o gpurun_out/r02s/conv_l0_96 python tools/ncu_layers.py --lex > gpurun_out/r02s/ncu1.log 2>&1
$NCU -k regex:conv_tc -s 22 -c 1 -o gpurun_out/r02s/conv_l0_32 python tools/ncu_layers.py --lex > gpurun_out/r02s/ncu2.log 2>&1
$NCU -k regex:conv_tc -s 46 -c 1 -o gpurun_out/r02s/conv_l2_128 python tools/ncu_layers.py --lex > gpurun_out/r02s/ncu3.log 2>&1
ls -la gpurun_out/r02s
