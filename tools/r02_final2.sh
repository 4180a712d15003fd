#!/bin/bash
# Second closing call of round 2 on ONE GPU: the randomised parity run with the repaired harness, and the 16384^2 map of
# BASELINE config 4 on one GPU with the final kernels (the 1-GPU row of the strong-scaling table).
O=gpurun_out
timeout 100 python tools/gpu_fuzz_parity.py --seconds 60 --max-size 220 --seed 11 > $O/fuzz_r02_final.log 2>&1; echo "fuzz rc=$?"; tail -4 $O/fuzz_r02_final.log
timeout 150 python bench.py --size 16384 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench16384_r02_final.json 2> $O/bench16384_r02_final.err; echo "bench16384 rc=$?"
cut -c1-400 $O/bench16384_r02_final.json
