#!/bin/bash
O=gpurun_out/r02_s2k; mkdir -p $O
timeout 300 python tools/host_profile.py > $O/host_profile.txt 2>&1; cat $O/host_profile.txt | head -70
