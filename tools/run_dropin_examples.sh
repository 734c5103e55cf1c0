#!/bin/bash
# Runs every example driver of the reference through oracle/_ref/pnol_examples_dropin (the reference's Examples.cpp compiled against
# include/pnol and linked to the B200 libraries, oracle/dropin_examples.cpp) and keeps the tail of each output under gpurun_out/.
mkdir -p gpurun_out/dropin
for d in testBFGS testBFGS_booth testBFGS_MPI testBFGSBnd testBFGSBndMPISW testBFGSBnd_MPI testLMExp testLMExpMPI testLMCubicLinearCoef \
         testGA testGAParallel testSimplexSearch testHessian testCreateObject testGradientEvaluation testGradientApproxMultMPI \
         testGradientApproxMultMPIRecur; do
	timeout 60 oracle/_ref/pnol_examples_dropin $d ${POOL:-8} > gpurun_out/dropin/$d.out 2> gpurun_out/dropin/$d.err
	echo "$d rc=$?" >> gpurun_out/dropin/summary.txt
	tail -c 3000 gpurun_out/dropin/$d.out > gpurun_out/dropin/$d.tail && rm gpurun_out/dropin/$d.out
done
cat gpurun_out/dropin/summary.txt
