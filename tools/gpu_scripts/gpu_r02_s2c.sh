#!/bin/bash
O=gpurun_out/r02_s2c; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:conv_tc -s 10 -c 1 -o $O/conv_l0_96 python tools/ncu_layers.py --lex > $O/ncu1.log 2>&1; tail -3 $O/ncu1.log
$NCU -k regex:conv_tc -s 10 -c 1 -o $O/conv_1x1_32_256 python tools/ncu_1x1.py > $O/ncu2.log 2>&1; tail -3 $O/ncu2.log
python tools/ncu_1x1.py
ls -la $O
