// TEST INFRASTRUCTURE -- objectives of the oracle (restated; shared by oracle/pnol_oracle.cpp and oracle/ref_cli.cpp).
// Never included by the product. See oracle/pnol_oracle.cpp for the pinning statement.
#ifndef PNOL_ORACLE_OBJECTIVES_H_
#define PNOL_ORACLE_OBJECTIVES_H_

#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector>

// ------------------------------------------------------------------------------------------------
// shared-definition transcendental functions, RESTATED (same operations as include/pnol/pnol_math.h)
// ------------------------------------------------------------------------------------------------
static double o_bits(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }
static double o_pow2i(int k) { return o_bits((uint64_t)(k + 1023) << 52); }

static double o_cos2pi(double x)
{
	double r = x - rint(x);
	r = fabs(r);
	int neg = r > 0.25;
	if (neg) r = 0.5 - r;
	double t = 3.141592653589793 * r;
	double t2 = t * t;
	double p = 4.7794773323873853e-14;
	p = fma(p, t2, -1.1470745597729725e-11);
	p = fma(p, t2, 2.0876756987868100e-09);
	p = fma(p, t2, -2.7557319223985888e-07);
	p = fma(p, t2, 2.4801587301587302e-05);
	p = fma(p, t2, -1.3888888888888889e-03);
	p = fma(p, t2, 4.1666666666666664e-02);
	p = fma(p, t2, -0.5);
	double c = fma(t2, p, 1.0);
	double v = fma(c + c, c, -1.0);
	return neg ? -v : v;
}

static double o_exp(double x)
{
	if (x != x) return x;
	if (x > 709.782712893384) return o_bits(0x7FF0000000000000ULL);
	if (x < -745.2) return 0.0;
	double kf = rint(x * 1.4426950408889634);
	double r = x - kf * 6.93147180369123816490e-01;
	r = r - kf * 1.90821492927058770002e-10;
	double p = 1.6059043836821613e-10;
	p = p * r + 2.0876756987868100e-09;
	p = p * r + 2.5052108385441720e-08;
	p = p * r + 2.7557319223985888e-07;
	p = p * r + 2.7557319223985893e-06;
	p = p * r + 2.4801587301587302e-05;
	p = p * r + 1.9841269841269841e-04;
	p = p * r + 1.3888888888888889e-03;
	p = p * r + 8.3333333333333332e-03;
	p = p * r + 4.1666666666666664e-02;
	p = p * r + 1.6666666666666666e-01;
	p = p * r + 0.5;
	p = p * r + 1.0;
	p = p * r + 1.0;
	int k = (int) kf;
	int k1 = k / 2;
	int k2 = k - k1;
	return (p * o_pow2i(k1)) * o_pow2i(k2);
}

// ------------------------------------------------------------------------------------------------
// objectives. kind ids are those of include/pnol_b200.h (restated as literals on purpose).
// ------------------------------------------------------------------------------------------------
struct OFunctor {
	int kind;
	const double * scalars;      // up to 8
	const long long * ints;      // up to 4
	const double * const * cols; // data columns
	long long m;
};

static double o_scalar(const OFunctor & f, const double * X, int n)
{
	switch (f.kind) {
		case 1: { // RosenbrockObject::objEval, Source/ExampleObjectives.hpp:87-103
			double value = 0;
			for (int k = 0; k < n - 1; k++)
				value = value + (100.0 * pow(X[k + 1] - pow(X[k], 2), 2) + pow(1 - X[k], 2));
			return value;
		}
		case 2: { // PowerObject::objEval, Source/ExampleObjectives.hpp:214-224, pow(x,power) as repeated products
			int power = (int) f.ints[0];
			double value = 0;
			for (int k = 0; k < n; k++) {
				double v = 1.0;
				for (int q = 0; q < power; q++) v = v * X[k];
				value = value + v;
			}
			return value;
		}
		case 3: { // BoothFunction::objEval, Source/ExampleObjectives.hpp:58-69
			double x = X[0], y = X[1];
			return pow(x + 2 * y - 7, 2) + pow(2 * x + y - 5, 2);
		}
		case 4: { // GoldsteinFunction::objEval, Source/ExampleObjectives.hpp:27-39
			double x = X[0], y = X[1];
			return (1 + pow(x + y + 1, 2) * (19 - 14 * x + 3 * pow(x, 2) - 14 * y + 6 * x * y + 3 * pow(y, 2))) *
			       (30 + pow(2 * x - 3 * y, 2) * (18 - 32 * x + 12 * pow(x, 2) + 48 * y - 36 * x * y + 27 * pow(y, 2)));
		}
		case 5: { // Rastrigin (ours, BASELINE.json config 4)
			double value = 10.0 * n;
			for (int k = 0; k < n; k++) value = value + (X[k] * X[k] - 10.0 * o_cos2pi(X[k]));
			return value;
		}
		case 6: { // ExpCurveObjectiveSingle::objEval, Source/ExampleObjectives.hpp:287-298 (exp -> shared exp)
			double Fnorm = 0;
			for (long long k = 0; k < f.m; k++) {
				double func = X[0] * o_exp(X[1] * f.cols[0][k]) + X[2];
				Fnorm = Fnorm + pow(f.cols[1][k] - func, 2);
			}
			return Fnorm;
		}
	}
	return NAN;
}

static double o_lorentz_term(double a, double c, double w, double t)
{
	double d = t - c;
	return a / (1.0 + w * (d * d));
}

// adjacent-pairs tree sum (recursive halves == pairs tree for a power-of-two count)
static double o_tree(const double * v, int count)
{
	if (count == 1) return v[0];
	int h = count / 2;
	return o_tree(v, h) + o_tree(v + h, h);
}

static void o_residual(const OFunctor & f, const double * X, int n, double * F)
{
	switch (f.kind) {
		case 101: // ExpCurveObjective::objEval, Source/ExampleObjectives.hpp:123-132 (exp -> shared exp)
			for (long long k = 0; k < f.m; k++) {
				double func = X[0] * o_exp(X[1] * f.cols[0][k]) + X[2];
				F[k] = f.cols[1][k] - func;
			}
			break;
		case 102: // CubicObjective::objEval, Source/ExampleObjectives.hpp:170-179; cols {pow(x,3), x, y}
			for (long long k = 0; k < f.m; k++) {
				double x = f.cols[1][k];
				double func = X[0] * f.cols[0][k] + X[1] * pow(x, 2) + X[2] * x + X[3];
				F[k] = f.cols[2][k] - func;
			}
			break;
		case 103: { // Lorentz sum (ours, configs 2 and 5)
			int K = n / 2;
			std::vector<double> terms(K);
			double w = f.scalars[0];
			for (long long i = 0; i < f.m; i++) {
				double t = f.cols[0][i];
				for (int k = 0; k < K; k++) terms[k] = o_lorentz_term(X[2 * k], X[2 * k + 1], w, t);
				F[i] = f.cols[1][i] - o_tree(terms.data(), K);
			}
			break;
		}
	}
}


#endif
