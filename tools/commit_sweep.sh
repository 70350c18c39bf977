#!/bin/bash
# variants of the concurrent commit: priority x persistent blocks per SM x block size
out=gpurun_out/r02_commit_sweep.jsonl
: > $out
run() { env "$@" timeout 120 python tools/commit_ab.py >> $out 2>> gpurun_out/r02_commit_sweep.err; }
run HODOR_CONCURRENT_COMMIT=0
run X=1
for R in 1 2 3 4; do run HODOR_COMMIT_PRIORITY=high HODOR_BACKFILL_PERSIST=$R; done
run HODOR_COMMIT_PRIORITY=high HODOR_BACKFILL_PERSIST=1 HODOR_BACKFILL_BLOCK=256
run HODOR_COMMIT_PRIORITY=high HODOR_BACKFILL_PERSIST=2 HODOR_BACKFILL_BLOCK=256
run HODOR_COMMIT_PRIORITY=equal HODOR_BACKFILL_PERSIST=2
run HODOR_COMMIT_PRIORITY=high
run X=1
