#!/bin/bash
# tcgen05.mma.kind::i8 issue-cost table (tools/umma_bench.cu) -> gpurun_out/umma_bench.txt
mkdir -p gpurun_out
out=gpurun_out/umma_bench4.txt
: > $out
for L in 1 2; do for M in 64 128; do for N in 16 80 160 256; do timeout 60 tools/umma_bench $M $N 0 0 1 4000 1 $L >> $out 2>&1; done; done; done
for H in 4 8; do timeout 60 tools/umma_bench 128 160 0 $H 1 4000 1 0 >> $out 2>&1; done
timeout 60 tools/umma_bench 128 160 1 0 1 4000 1 0 >> $out 2>&1
timeout 60 tools/umma_bench 128 144 0 0 1 4000 1 0 >> $out 2>&1
timeout 60 tools/umma_bench 128 176 0 0 1 4000 1 0 >> $out 2>&1
timeout 60 tools/umma_bench 128 192 0 0 1 4000 1 0 >> $out 2>&1
cat $out
