mkdir -p gpurun_out/r02l
for cfg in "0 64" "1 64" "2 64"; do set -- $cfg
echo "== TMA_GATHER=$1 ZPAD=$2"; LIDAL_TMA_GATHER=$1 LIDAL_ZPAD=$2 timeout 120 python tools/ncu_layers.py --lex 2>&1 | grep lvl
done > gpurun_out/r02l/zpad.txt 2>&1
cat gpurun_out/r02l/zpad.txt
