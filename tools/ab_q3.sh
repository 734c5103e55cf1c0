#!/bin/bash
# A/B of the FD quotient in three operations (exact_div.cuh: div_exact3_core) against the five-operation one, on one build: the
# verdict depends on dX (1e-7 qualifies, 1e-5 does not); beside it a build without the pair (-DLORENTZ_NO_Q3, tools/build_variants.sh
# residual_kernels NOQ3 -DLORENTZ_NO_Q3). Output: gpurun_out/ab_q3.log
out=gpurun_out/ab_q3.log
: > $out
V=parallelnonlinearoptimizationlibrary_b200/lib/var/libpnol_NOQ3.so
for rep in 1 2; do
echo "== LM step, three-operation quotient (dx 1e-7), kernel pair" >> $out
timeout 100 python tools/time_lm_step.py >> $out 2>&1
echo "== LM step, five-operation quotient (dx 1e-5), kernel pair" >> $out
PROF_DX=1e-5 timeout 100 python tools/time_lm_step.py >> $out 2>&1
echo "== LM step, five-operation quotient only (-DLORENTZ_NO_Q3 build, dx 1e-7)" >> $out
PNOL_B200_LIB=$V VARIANT=NOQ3 timeout 100 python tools/time_lm_step.py >> $out 2>&1
done
echo "== fd_jacobian alone: 3-op / 5-op / NOQ3" >> $out
timeout 100 python tools/time_jacobian.py >> $out 2>&1
PROF_DX=1e-5 timeout 100 python tools/time_jacobian.py >> $out 2>&1
PNOL_B200_LIB=$V timeout 100 python tools/time_jacobian.py >> $out 2>&1
echo "== one rank's share of the 8-GPU run (m = 500k): 3-op / NOQ3" >> $out
PROF_M=500000 timeout 100 python tools/time_lm_step.py >> $out 2>&1
PROF_M=500000 PNOL_B200_LIB=$V VARIANT=NOQ3 timeout 100 python tools/time_lm_step.py >> $out 2>&1
echo "== parity" >> $out
timeout 400 python -m pytest tests/test_gpu_parity_core.py tests/test_gpu_edge_cases.py tests/test_gpu_fullsize.py tests/test_gpu_baseline_shapes.py -x -q -m gpu 2>&1 | tail -4 >> $out
cat $out
