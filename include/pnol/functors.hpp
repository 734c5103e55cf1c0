/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
/*
 * functors.hpp -- device functors: ONE __host__ __device__ definition per objective.
 *
 * A functor is the device twin of a PNOL Objective / MultiObjective (Source/PNOL_Objective.hpp:25-62 of
 * the reference). The same source is used by
 *   - the CUDA kernels (one perturbed point per thread; the point is seen through an accessor so the base
 *     point can sit in shared memory while a single coordinate is perturbed in a register), and
 *   - the host objEval of the matching C++ plugin class (include/pnol/ExampleObjectives.hpp),
 * so that host and device agree bit for bit (compiled with -fmad=false / -ffp-contract=off).
 *
 * Scalar functor concept:
 *     struct F { static constexpr int kKind;  template<class Acc> static double eval(const FunctorParams&, const Acc& X, int n); };
 * Residual functor concept (row i of the residual vector, data columns in FunctorParams::col):
 *     struct R { static constexpr int kKind;  template<class Acc> static double residual(const FunctorParams&, const Acc& X, int n, long long i); };
 * Acc needs only `double operator[](int) const`.
 */
#ifndef PNOL_FUNCTORS_HPP_
#define PNOL_FUNCTORS_HPP_

#include "pnol_math.h"
#include "../pnol_b200.h"

namespace pnol {

/* parameters of a functor as the kernels see them (POD, passed by value as a kernel argument): scalars[PNOL_MAX_SCALARS],
 * ints[PNOL_MAX_INTS], col[PNOL_MAX_COLUMNS] (data columns; device pointers on the device side), m -- the C struct of pnol_b200.h */
typedef pnol_functor_params FunctorParams;

/* ---- accessors ---- */
struct PtrAcc {                      /* plain array */
	const double * p;
	PNOL_HD double operator[](int j) const { return p[j]; }
};
struct PerturbAcc {                  /* base point with ONE coordinate replaced (forward-difference stencil) */
	const double * base; int i; double xi;
	PNOL_HD double operator[](int j) const { return j == i ? xi : base[j]; }
};
struct Perturb2Acc {                 /* base point with TWO coordinates replaced (FD Hessian) */
	const double * base; int i; double xi; int j2; double xj;
	PNOL_HD double operator[](int j) const { return j == i ? xi : (j == j2 ? xj : base[j]); }
};

/* ---- scalar objectives ---- */

/* Source/ExampleObjectives.hpp:87-103: sum_k 100 (x_{k+1} - x_k^2)^2 + (1 - x_k)^2, sequential sum.
 * pow(.,2) is written as a product: GCC folds the literal pow(x,2) to x*x at -O2. */
struct RosenbrockFunctor {
	static constexpr int kKind = PNOL_F_ROSENBROCK;
	template <class Acc> PNOL_HD static double eval(const FunctorParams &, const Acc & X, int n)
	{
		double value = 0;
		double xk = X[0];
		for (int k = 0; k < n - 1; k++) {
			double xk1 = X[k + 1];
			double a = xk1 - xk * xk;
			double b = 1 - xk;
			value = value + (100.0 * (a * a) + b * b);
			xk = xk1;
		}
		return value;
	}
};

/* Source/ExampleObjectives.hpp:214-224 with the runtime pow(X[k], power) replaced by repeated
 * multiplication ((x*x)*x...): libm pow is not reproducible on the device. */
struct PowerFunctor {
	static constexpr int kKind = PNOL_F_POWER;
	/* separable: f = init + term(x_0) + term(x_1) + ... summed in index order (see the note above RastriginFunctor) */
	static constexpr bool kSeparable = true;
	PNOL_HD static double sep_init(const FunctorParams &, int) { return 0.0; }
	PNOL_HD static double sep_term(const FunctorParams & P, double x)
	{
		int power = (int) P.ints[0];
		double v = 1.0;
		for (int q = 0; q < power; q++) v = v * x;
		return v;
	}
	template <class Acc> PNOL_HD static double eval(const FunctorParams & P, const Acc & X, int n)
	{
		double value = sep_init(P, n);
		for (int k = 0; k < n; k++) value = value + sep_term(P, X[k]);
		return value;
	}
};

/* Source/ExampleObjectives.hpp:58-69 */
struct BoothFunctor {
	static constexpr int kKind = PNOL_F_BOOTH;
	template <class Acc> PNOL_HD static double eval(const FunctorParams &, const Acc & X, int)
	{
		double x = X[0], y = X[1];
		double a = x + 2 * y - 7;
		double b = 2 * x + y - 5;
		return a * a + b * b;
	}
};

/* Source/ExampleObjectives.hpp:27-39 */
struct GoldsteinFunctor {
	static constexpr int kKind = PNOL_F_GOLDSTEIN;
	template <class Acc> PNOL_HD static double eval(const FunctorParams &, const Acc & X, int)
	{
		double x = X[0], y = X[1];
		double s = x + y + 1;
		double d = 2 * x - 3 * y;
		return (1 + (s * s) * (19 - 14 * x + 3 * (x * x) - 14 * y + 6 * x * y + 3 * (y * y))) *
		       (30 + (d * d) * (18 - 32 * x + 12 * (x * x) + 48 * y - 36 * x * y + 27 * (y * y)));
	}
};

/* ours (BASELINE.json config 4): f = 10 n + sum_k (x_k^2 - 10 cos(2 pi x_k)), sequential sum.
 * A functor may declare itself SEPARABLE (kSeparable, sep_init, sep_term): f = init + term(x_0) + term(x_1) + ... with the sum
 * taken in index order. The population sweep then computes the terms one gene per LANE straight from coalesced loads and
 * only the (cheap, order-preserving) summation one individual per lane -- same bits as eval(), which is defined through the
 * same two functions. */
struct RastriginFunctor {
	static constexpr int kKind = PNOL_F_RASTRIGIN;
	static constexpr bool kSeparable = true;
	PNOL_HD static double sep_init(const FunctorParams &, int n) { return 10.0 * n; }
	PNOL_HD static double sep_term(const FunctorParams &, double x) { return x * x - 10.0 * cos2pi(x); }
	template <class Acc> PNOL_HD static double eval(const FunctorParams & P, const Acc & X, int n)
	{
		double value = sep_init(P, n);
		PNOL_UNROLL(4)
		for (int k = 0; k < n; k++) value = value + sep_term(P, X[k]);
		return value;
	}
};

/* Source/ExampleObjectives.hpp:287-298 with exp -> exp_hd; columns {x, y} */
struct ExpCurveSingleFunctor {
	static constexpr int kKind = PNOL_F_EXPCURVE_SINGLE;
	template <class Acc> PNOL_HD static double eval(const FunctorParams & P, const Acc & X, int)
	{
		double Fnorm = 0;
		double x0 = X[0], x1 = X[1], x2 = X[2];
		for (long long k = 0; k < P.m; k++) {
			double func = x0 * exp_hd(x1 * P.col[0][k]) + x2;
			double d = P.col[1][k] - func;
			Fnorm = Fnorm + d * d;
		}
		return Fnorm;
	}
};

/* ---- residual models ---- */

/* Source/ExampleObjectives.hpp:123-132 with exp -> exp_hd; columns {x, y} */
struct ExpCurveFunctor {
	static constexpr int kKind = PNOL_F_EXPCURVE;
	template <class Acc> PNOL_HD static double residual(const FunctorParams & P, const Acc & X, int, long long i)
	{
		double func = X[0] * exp_hd(X[1] * P.col[0][i]) + X[2];
		return P.col[1][i] - func;
	}
};

/* Source/ExampleObjectives.hpp:170-179; columns {x^3 computed by the host's pow(x,3), x, y}; pow(x,2) -> x*x */
struct CubicFunctor {
	static constexpr int kKind = PNOL_F_CUBIC;
	template <class Acc> PNOL_HD static double residual(const FunctorParams & P, const Acc & X, int, long long i)
	{
		double x3 = P.col[0][i], x = P.col[1][i];
		double func = X[0] * x3 + X[1] * (x * x) + X[2] * x + X[3];
		return P.col[2][i] - func;
	}
};

/* ours (BASELINE.json configs 2 and 5): r_i = y_i - S_i,  S_i = balanced-tree sum over k < K of
 *   term_k = a_k / (1 + w (t_i - c_k)^2),   x = (a_0, c_0, a_1, c_1, ...), n = 2K, K a power of two.
 * The tree is the adjacent-pairs tree: level l sums nodes 2j and 2j+1 of level l-1. A single-parameter
 * perturbation changes one leaf, hence only log2 K nodes: the structured Jacobian kernel exploits that and
 * returns the same bits as n+1 black-box evaluations of this function. */
PNOL_HD double lorentz_term(double a, double c, double w, double t)
{
	double d = t - c;
	double q = d * d;
	double e = w * q;
	double den = 1.0 + e;
	return a / den;
}

struct LorentzSumFunctor {
	static constexpr int kKind = PNOL_F_LORENTZ_SUM;
	static constexpr int kMaxLog2K = 12;
	template <class Acc> PNOL_HD static double residual(const FunctorParams & P, const Acc & X, int n, long long i)
	{
		double w = P.scalars[0];
		double t = P.col[0][i];
		int K = n / 2;
		/* adjacent-pairs tree evaluated as a binary counter: stack[l] holds a finished subtree of 2^l leaves */
		double stack[kMaxLog2K + 1];
		for (int k = 0; k < K; k++) {
			double v = lorentz_term(X[2 * k], X[2 * k + 1], w, t);
			int l = 0;
			int kk = k;
			while (kk & 1) { v = stack[l] + v; kk >>= 1; l++; }
			stack[l] = v;
		}
		int top = 0;
		while ((1 << top) < K) top++;
		return P.col[1][i] - stack[top];
	}
};

} // namespace pnol

#endif /* PNOL_FUNCTORS_HPP_ */
