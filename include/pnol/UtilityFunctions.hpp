/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
/*
 * UtilityFunctions.hpp -- the small host helpers the PNOL plugin API expects to find in scope.
 *
 * The reference takes them from an un-vendored sibling library ("UtilityFunctions/utilityFunctions.hpp",
 * /root/reference/Source/LevenbergMarquardtMPI.hpp:20); user objectives call linspace() / print1DVector()
 * unqualified (Source/ExampleObjectives.hpp:141). These are OUR definitions (conventions stated in DESIGN.md):
 * sequential sums, first extremum on ties, linspace a + i (b-a)/(N-1) with the end point forced to b.
 * Only O(n) host control arithmetic lives here; nothing on the device path.
 */
#ifndef PNOL_UTILITYFUNCTIONS_HPP_
#define PNOL_UTILITYFUNCTIONS_HPP_

#include <cmath>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>

inline double vector2Norm(const std::vector<double> & v)
{
	double s = 0;
	for (size_t i = 0; i < v.size(); i++) s = s + v[i] * v[i];
	return std::sqrt(s);
}
inline double dotProd(const std::vector<double> & a, const std::vector<double> & b)
{
	double s = 0;
	for (size_t i = 0; i < a.size(); i++) s = s + a[i] * b[i];
	return s;
}
inline int mod(int a, int b) { int r = a % b; if (r < 0) r += b; return r; }
inline double sign(double x) { if (x > 0) return 1.0; if (x < 0) return -1.0; return 0.0; }
inline void linspace(double a, double b, int N, std::vector<double> & out)
{
	out.resize(N > 0 ? N : 0);
	if (N <= 0) return;
	if (N == 1) { out[0] = a; return; }
	double h = (b - a) / (N - 1);
	for (int i = 0; i < N; i++) out[i] = a + i * h;
	out[N - 1] = b;
}
inline void vectorMin(const std::vector<double> & v, int N, double & val, int & idx)
{
	val = v[0]; idx = 0;
	for (int i = 1; i < N; i++) if (v[i] < val) { val = v[i]; idx = i; }
}
inline void vectorMax(const std::vector<double> & v, int N, double & val, int & idx)
{
	val = v[0]; idx = 0;
	for (int i = 1; i < N; i++) if (v[i] > val) { val = v[i]; idx = i; }
}
template <typename T> void print1DVector(const std::vector<T> & v)
{
	std::cout << "[";
	for (size_t i = 0; i < v.size(); i++) { std::cout << std::setprecision(17) << v[i]; if (i + 1 < v.size()) std::cout << ", "; }
	std::cout << "]" << std::endl;
}
inline void print2DVector(const std::vector<std::vector<double> > & A)
{
	for (size_t i = 0; i < A.size(); i++) print1DVector(A[i]);
}
inline void setIdentity(std::vector<std::vector<double> > & A)
{
	for (size_t i = 0; i < A.size(); i++)
		for (size_t j = 0; j < A[i].size(); j++) A[i][j] = (i == j) ? 1.0 : 0.0;
}
// small host inverse for user code (the reference's example driver testHessian calls it, Source/Examples.cpp:296):
// Gauss-Jordan with partial pivoting on [A | I]. The algorithms here never call it -- an initial inverse Hessian is
// solved on the device (pnol::InverseHessian::setFromInverseOfFDHessian).
inline void matrixInverse(const std::vector<std::vector<double> > & A, std::vector<std::vector<double> > & Ainv)
{
	size_t n = A.size();
	std::vector<std::vector<double> > M(A);
	Ainv.assign(n, std::vector<double>(n, 0.0));
	for (size_t i = 0; i < n; i++) Ainv[i][i] = 1.0;
	for (size_t c = 0; c < n; c++)
	{
		size_t piv = c;
		for (size_t r = c + 1; r < n; r++) if (std::fabs(M[r][c]) > std::fabs(M[piv][c])) piv = r;
		M[c].swap(M[piv]); Ainv[c].swap(Ainv[piv]);
		double d = M[c][c];
		for (size_t j = 0; j < n; j++) { M[c][j] = M[c][j] / d; Ainv[c][j] = Ainv[c][j] / d; }
		for (size_t r = 0; r < n; r++)
		{
			if (r == c) continue;
			double f = M[r][c];
			if (f == 0.0) continue;
			for (size_t j = 0; j < n; j++) { M[r][j] = M[r][j] - f * M[c][j]; Ainv[r][j] = Ainv[r][j] - f * Ainv[c][j]; }
		}
	}
}

#endif
