#!/bin/bash
# runs every mode of the gather microbenchmark (each in its own process: a faulting mode must not stop the rest)
OUT=gpurun_out/${1:-mb}_gather.txt
: > $OUT
B=tools/bin/gather_microbench
for args in "ldg4 1024" "ldg4 512" "ldg4s 1024" "ldg8 1024" "ldg2 1024" "ldg1row 1024" "lds 1024" \
            "bulk 1024 2" "bulk 512 4" "bulk 512 8" "bulk 256 8" "bulk 128 8" "gather4 512 4" "gather4 512 8" "gather4 256 8" \
            "mix 1024 4" "mix 512 8"; do
  timeout 60 $B $args >> $OUT 2>&1 || echo "FAILED($?): $args" >> $OUT
done
cat $OUT
