#!/bin/bash
O=gpurun_out/r02_s2o; mkdir -p $O
timeout 120 python tools/host_stall_probe.py 30 > $O/stall_probe.txt 2>&1; cat $O/stall_probe.txt | head -40
ps aux --sort=-%cpu | head -15 > $O/ps.txt; cat $O/ps.txt | cut -c1-150
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/fs/cgroup/cpu.stat 2>/dev/null | head -8
