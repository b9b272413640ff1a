#!/bin/bash
O=gpurun_out/r02_s2t; mkdir -p $O
timeout 300 python tools/host_profile.py > $O/host_profile.txt 2>&1; head -45 $O/host_profile.txt
