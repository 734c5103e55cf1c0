// exact_div.cuh -- division by a loop-invariant divisor with the bits of the IEEE quotient, in 5 FP64 issue slots
// instead of the ~9 DFMA + MUFU + fix-up branch of the generic division sequence.
//
// With r = RN(1/d) (one true division, hoisted), q0 = RN(x r) is within 2 ulp of x/d, one FMA correction step makes
// q1 faithful, and by Markstein's theorem (IBM J. R&D 34(1), 1990; Muller et al., Handbook of Floating-Point
// Arithmetic, "division by a constant") the second step q2 = RN(q1 + RN(x - q1 d) r) is the correctly rounded
// quotient, provided no intermediate over/underflows and the significand of d is not all ones. Divisors or
// dividends outside those conditions take the ordinary division. tests/test_gpu_exact_div.py checks 10^8 random and
// adversarial pairs against `/` on the device; oracle-side the FD quotient stays a plain division
// (Source/PNOL_Objective.cpp:31, :192).
#pragma once

namespace pnol {

struct RecipDiv {
	double d;   // divisor
	double r;   // RN(1/d), or 0 when the fast path must not be used for this divisor
};

__device__ __forceinline__ RecipDiv make_recip(double d)
{
	const unsigned long long bits = (unsigned long long) __double_as_longlong(d);
	const int e = (int) ((bits >> 52) & 0x7ff);
	const bool all_ones = (bits & 0xFFFFFFFFFFFFFULL) == 0xFFFFFFFFFFFFFULL;
	const bool ok = e > 1023 - 200 && e < 1023 + 200 && !all_ones;
	RecipDiv rd;
	rd.d = d;
	rd.r = ok ? 1.0 / d : 0.0;
	return rd;
}

__device__ __forceinline__ double div_exact(double x, const RecipDiv & rd)
{
	const int ex = (__double2hiint(x) >> 20) & 0x7ff;
	const bool in_range = ex > 1023 - 700 && ex < 1023 + 700;
	if (rd.r != 0.0 && (in_range || x == 0.0)) {
		const double q0 = x * rd.r;
		const double r0 = fma(-q0, rd.d, x);
		const double q1 = fma(r0, rd.r, q0);
		const double r1 = fma(-q1, rd.d, x);
		const double q2 = fma(r1, rd.r, q1);
		return x == 0.0 ? q0 : q2;      // q0 carries the sign of a zero quotient
	}
	return x / rd.d;
}

} // namespace pnol
