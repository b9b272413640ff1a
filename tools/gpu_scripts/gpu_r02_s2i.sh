#!/bin/bash
O=gpurun_out/r02_s2i; mkdir -p $O
timeout 300 python tools/pipeline_jitter.py 16 > $O/jitter_default.txt 2>&1; cat $O/jitter_default.txt | tail -17
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python tools/pipeline_jitter.py 16 > $O/jitter_conn32.txt 2>&1; cat $O/jitter_conn32.txt | tail -17
