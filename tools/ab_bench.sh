#!/bin/bash
# usage: ab.sh name:libpath ...   (runs GPU tests for each non-default lib, then the bench)
for spec in "$@"; do
  v=${spec%%:*}; lib=${spec#*:}
  if [ "$lib" != default ]; then export HODOR_B200_LIB=$PWD/$lib; else unset HODOR_B200_LIB; fi
  if [ "$lib" != default ]; then timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "ntt or lde or fft or fri or shard" 2>&1 | tail -2; fi
  timeout 280 python bench.py --steps 5 --warmup 3 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_$v.json"))
print("$v", round(d["ms_per_step"],3), {k:round(x["ms_per_launch"],3) for k,x in d["roofline"]["kernels"].items()}, "ntt", round(d["ntt"]["ms_per_step"],3), "fri", round(d["fri"]["ms_per_step"],3))
PY
done
