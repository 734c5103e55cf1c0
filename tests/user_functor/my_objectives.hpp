/* Example of objectives that are NOT built into libpnol_b200.so (tests/test_gpu_user_functor.py).
 * One __host__ __device__ definition per objective, written to the functor concept of include/pnol/functors.hpp; the reference's
 * user writes the same arithmetic inside Objective::objEval / MultiObjective::objEval (Source/PNOL_Objective.hpp:29, :57). */
#ifndef MY_OBJECTIVES_HPP_
#define MY_OBJECTIVES_HPP_

#include "pnol/functors.hpp"

enum { MY_F_TRID = PNOL_F_USER_SCALAR_BASE + 1, MY_F_STYBLINSKI = PNOL_F_USER_SCALAR_BASE + 2, MY_F_GAUSSFIT = PNOL_F_USER_RESIDUAL_BASE + 1 };

/* Trid: f = sum_i (x_i - 1)^2 - sum_{i >= 1} x_i x_{i-1}   (not separable: the generic sweep kernels) */
struct TridFunctor {
	static constexpr int kKind = MY_F_TRID;
	template <class Acc> PNOL_HD static double eval(const pnol::FunctorParams &, const Acc & X, int n)
	{
		double a = 0, b = 0;
		double prev = X[0];
		a = a + (prev - 1.0) * (prev - 1.0);
		for (int i = 1; i < n; i++) {
			const double xi = X[i];
			a = a + (xi - 1.0) * (xi - 1.0);
			b = b + xi * prev;
			prev = xi;
		}
		return a - b;
	}
};

/* Styblinski-Tang with a user scale: f = s * sum_i (x_i^4 - 16 x_i^2 + 5 x_i); separable -> the row-wise population sweep */
struct StyblinskiFunctor {
	static constexpr int kKind = MY_F_STYBLINSKI;
	static constexpr bool kSeparable = true;
	PNOL_HD static double sep_init(const pnol::FunctorParams &, int) { return 0.0; }
	PNOL_HD static double sep_term(const pnol::FunctorParams & P, double x)
	{
		const double x2 = x * x;
		return P.scalars[0] * (x2 * x2 - 16.0 * x2 + 5.0 * x);
	}
	template <class Acc> PNOL_HD static double eval(const pnol::FunctorParams & P, const Acc & X, int n)
	{
		double v = sep_init(P, n);
		for (int k = 0; k < n; k++) v = v + sep_term(P, X[k]);
		return v;
	}
};

/* three-parameter Gaussian peak fit: r_i = y_i - (A exp(-(t_i - mu)^2 s) + c0), x = (A, mu, s), scalars[0] = c0, columns {t, y};
 * exp through the library's shared host/device exp (include/pnol/pnol_math.h) so that host and device agree bit for bit */
struct GaussFitFunctor {
	static constexpr int kKind = MY_F_GAUSSFIT;
	template <class Acc> PNOL_HD static double residual(const pnol::FunctorParams & P, const Acc & X, int, long long i)
	{
		const double d = P.col[0][i] - X[1];
		const double model = X[0] * pnol::exp_hd(-(d * d) * X[2]) + P.scalars[0];
		return P.col[1][i] - model;
	}
};

#endif
