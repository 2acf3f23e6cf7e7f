#!/bin/bash
# Sweep descriptor candidates for the five tcgen05.mma operand forms (see tools/umma_probe.cu).
P=./build/umma_probe
run() { timeout 30 $P "$@" 2>&1 || echo "  -> exit $? for args: $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
run 0
# mode 1: both K-major SW64
run 1 16 512 32 16 512 32
run 1 0 512 32 0 512 32
run 1 8192 512 32 8192 512 32
# mode 2: TS, B MN-major SW64
run 2 0 0 0 8192 512 1024
run 2 0 0 0 512 8192 1024
# mode 3: A MN-major, B MN-major
run 3 8192 512 1024 8192 512 1024
run 3 512 8192 1024 512 8192 1024
run 3 8192 512 1024 512 8192 1024
run 3 512 8192 1024 8192 512 1024
# mode 4: A K-major thread-written, B MN-major
run 4 16 512 32 8192 512 1024
run 4 16 512 32 512 8192 1024
