#!/bin/bash
# one full ncu capture of the batched (channelizer) chain kernel; output: gpurun_out/c5/c5.ncu-rep
mkdir -p gpurun_out/c5 && cd gpurun_out/c5 || exit 1
ncu --kernel-name-base demangled --set full --clock-control none --import-source on \
    -k 'regex:k_chain1024<\(int\)3, \(bool\)1' -s 2 -c 1 -f -o c5 \
    python ../../bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload c5 > c5.log 2>&1
tail -3 c5.log; ls -la
