/*
 * pnol_math.h -- transcendental functions with ONE definition for host and device.
 *
 * FD derivatives amplify a 1-ulp difference in f by 1/h (hard part 1), so an objective
 * that wants its device functor to agree with its host objEval to the last bit cannot call libm (glibc and
 * CUDA differ in the last ulp). These functions use only + - * / rint and bit operations, are compiled without
 * FMA contraction on both sides (nvcc -fmad=false, g++ -ffp-contract=off), and therefore return identical bits
 * on CPU and GPU. Accuracy is a few ulp -- they are not correctly rounded and make no such claim.
 */
#ifndef PNOL_MATH_H_
#define PNOL_MATH_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PNOL_HD __host__ __device__ __forceinline__
#else
#define PNOL_HD inline
#endif

namespace pnol {

PNOL_HD double bits_to_double(uint64_t b)
{
#if defined(__CUDA_ARCH__)
	return __longlong_as_double((long long) b);
#else
	double d; memcpy(&d, &b, sizeof d); return d;
#endif
}

/* 2^k for -1022 <= k <= 1023 */
PNOL_HD double pow2i(int k) { return bits_to_double((uint64_t)(k + 1023) << 52); }

/* cos(2 pi x): exact reduction to r in [0, 1/8], then a Taylor polynomial in (2 pi r)^2 */
PNOL_HD double cos2pi(double x)
{
	double r = x - rint(x);                 /* exact, |r| <= 0.5 */
	r = fabs(r);
	bool neg = r > 0.25;
	if (neg) r = 0.5 - r;                   /* exact */
	bool use_sin = r > 0.125;
	if (use_sin) r = 0.25 - r;              /* exact */
	double t = 6.283185307179586 * r;       /* in [0, pi/4] */
	double t2 = t * t;
	double v;
	if (use_sin) {
		/* sin t = t (1 - t2/3! + t2^2/5! - ... - t2^8/17!) */
		double p = -2.8114572543455206e-15;          /* -1/17! */
		p = p * t2 + 7.6471637318198164e-13;          /*  1/15! */
		p = p * t2 + -1.6059043836821613e-10;         /* -1/13! */
		p = p * t2 + 2.5052108385441720e-08;          /*  1/11! */
		p = p * t2 + -2.7557319223985893e-06;         /* -1/9!  */
		p = p * t2 + 1.9841269841269841e-04;          /*  1/7!  */
		p = p * t2 + -8.3333333333333332e-03;         /* -1/5!  */
		p = p * t2 + 1.6666666666666666e-01;          /*  1/3!  (sign folded below) */
		/* p now holds 1/3! - t2/5! + ...; sin t = t - t^3 * p */
		v = t - (t * t2) * p;
	} else {
		/* cos t = 1 - t2/2! + t2^2/4! - ... + t2^8/16! */
		double p = 4.7794773323873853e-14;           /*  1/16! */
		p = p * t2 + -1.1470745597729725e-11;         /* -1/14! */
		p = p * t2 + 2.0876756987868100e-09;          /*  1/12! */
		p = p * t2 + -2.7557319223985888e-07;         /* -1/10! */
		p = p * t2 + 2.4801587301587302e-05;          /*  1/8!  */
		p = p * t2 + -1.3888888888888889e-03;         /* -1/6!  */
		p = p * t2 + 4.1666666666666664e-02;          /*  1/4!  */
		p = p * t2 + -0.5;                            /* -1/2!  */
		v = 1.0 + t2 * p;
	}
	return neg ? -v : v;
}

/* exp(x): k = rint(x log2 e), Cody-Waite reduction, degree-13 Taylor polynomial, scaling by 2^k in two steps */
PNOL_HD double exp_hd(double x)
{
	if (x != x) return x;
	if (x > 709.782712893384) return bits_to_double(0x7FF0000000000000ULL);
	if (x < -745.2) return 0.0;
	double kf = rint(x * 1.4426950408889634);
	double r = x - kf * 6.93147180369123816490e-01;      /* ln2 high part (low 21 bits zero): product exact */
	r = r - kf * 1.90821492927058770002e-10;             /* ln2 low part */
	double p = 1.6059043836821613e-10;                   /* 1/13! */
	p = p * r + 2.0876756987868100e-09;                  /* 1/12! */
	p = p * r + 2.5052108385441720e-08;                  /* 1/11! */
	p = p * r + 2.7557319223985888e-07;                  /* 1/10! */
	p = p * r + 2.7557319223985893e-06;                  /* 1/9!  */
	p = p * r + 2.4801587301587302e-05;                  /* 1/8!  */
	p = p * r + 1.9841269841269841e-04;                  /* 1/7!  */
	p = p * r + 1.3888888888888889e-03;                  /* 1/6!  */
	p = p * r + 8.3333333333333332e-03;                  /* 1/5!  */
	p = p * r + 4.1666666666666664e-02;                  /* 1/4!  */
	p = p * r + 1.6666666666666666e-01;                  /* 1/3!  */
	p = p * r + 0.5;
	p = p * r + 1.0;
	p = p * r + 1.0;
	int k = (int) kf;
	int k1 = k / 2;
	int k2 = k - k1;
	return (p * pow2i(k1)) * pow2i(k2);
}

} // namespace pnol

#endif /* PNOL_MATH_H_ */
