#!/bin/bash
O=gpurun_out/r02_s2g; mkdir -p $O
for nb in 1 2; do
LIDAL_NB_MAX=$nb LIDAL_LIB=$PWD/lidal_b200/liblidal_b200_dbg.so LIDAL_DBG=128 timeout 300 python tools/ncu_layers.py --lex 2>&1 | grep "lvl\|conv dbg" | awk '/conv dbg/{c++; if (c%12==0) print; next} {print}' > $O/segments_nb$nb.txt
cat $O/segments_nb$nb.txt
done
