mkdir -p gpurun_out/r02d
for pm in 0 1; do for dbg in 128 167; do
echo "== PROD_MODE=$pm DBG=$dbg"; LIDAL_PROD_MODE=$pm LIDAL_DBG=$dbg timeout 120 python tools/ncu_layers.py --lex 2>&1 | grep "lvl\|conv dbg" | awk '/conv dbg/{c++; if (c%12==0) print; next} {print}'
done; done > gpurun_out/r02d/cycles.txt 2>&1
cat gpurun_out/r02d/cycles.txt
