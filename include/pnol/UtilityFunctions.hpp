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

#endif
