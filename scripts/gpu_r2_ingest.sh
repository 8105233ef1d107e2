#!/bin/bash
# SM ingest microbenchmark (ring depth, box size, unicast vs cluster multicast of a shared A tile)
mkdir -p gpurun_out
timeout 300 scripts/sm_ingest_bench.bin 4800 > gpurun_out/r2_sm_ingest.log 2>&1; echo "exit=$?" >> gpurun_out/r2_sm_ingest.log
cat gpurun_out/r2_sm_ingest.log
