// host_logic_device.cpp -- TEST INFRASTRUCTURE, never shipped, never loaded by the product (the product's libraries are
// lib/libpnol_b200.so + lib/libpnol_b200_host.so and have no CPU path).
//
// The part of the C-ABI (include/pnol_b200.h) that the BFGS family, SimplexSearch, LevMarq[MPI] and the stencil members of the host
// C++ mirror call, answered by the CPU oracle (pnol_oracle.cpp) instead of CUDA kernels. tests/test_host_logic_cpu.py links the
// UNMODIFIED host sources (parallelnonlinearoptimizationlibrary_b200/host/*.cpp) against this file into a test-only library and runs
// the reference's control flow on a box without a GPU: with the oracle's arithmetic (= the oracle shim's: sequential sums, the
// literal two-product updateHessianInv) under them, the host controllers reproduce the verbatim reference's iterates BIT FOR BIT,
// which isolates every difference seen on the GPU to the summation order of the device's dense algebra. The Levenberg-Marquardt
// entry points (residuals, FD Jacobian, one iteration's device work) are answered the same way, the GA state machine by replaying the
// oracle's restatement of the whole loop up to the requested generation. The stand-alone GA stage entry points are not provided.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/pnol_b200.h"

extern "C" {
// pnol_oracle.cpp
double oracle_obj_eval(int kind, const double * scalars, const long long * ints, const double * const * cols, long long m, const double * X, int n);
void oracle_eval_batch(int kind, const double * scalars, const long long * ints, const double * const * cols, long long m,
                       const double * pts, long long B, int n, long long ld, const unsigned char * indicator, double * f_out);
void oracle_fd_gradient(int kind, const double * scalars, const long long * ints, const double * const * cols, long long m,
                        const double * X, const double * dX, int N, double * dFdX, double * f0);
double oracle_eval_recur(int kind, const double * scalars, const long long * ints, const double * const * cols, long long m,
                         const double * X, int N, const double * constantX, const unsigned char * ind, int Nparam);
void oracle_fd_gradient_recur(int kind, const double * scalars, const long long * ints, const double * const * cols, long long m,
                              const double * X, const double * dX, int N, const double * constantX, const unsigned char * ind,
                              int Nparam, double * dFdX, double * f0);
void oracle_fd_hessian(int kind, const double * scalars, const long long * ints, const double * const * cols, long long m,
                       const double * X, const double * dX, int N, double * B);
void oracle_alpha_pool(int kind, const double * scalars, const long long * ints, const double * const * cols, long long m,
                       const double * X, const double * p, int n, const double * alpha, int npool, double dalpha,
                       const unsigned char * eval_ind, const double * constantX, const unsigned char * const_ind, int nfull,
                       double * phi, double * dphi, int * bad);
void oracle_residual(int kind, const double * scalars, const long long * ints, const double * const * cols, long long m,
                     const double * X, int n, double * F);
void oracle_fd_jacobian(int kind, const double * scalars, const long long * ints, const double * const * cols, long long m,
                        const double * X, const double * dX, int N, double * J, double * Fout);
void oracle_lm_normal_eq(const double * J, const double * F, long long m, int n, double lambda, double * JTJ, double * A, double * rhs);
void oracle_update_hinv(double * D, const double * g, const double * s, int n);
void oracle_matvec_neg(const double * D, const double * g, int n, double * p);
void oracle_lu_solve(const double * A, const double * b, int n, double * x);
int oracle_check_box_bounds(double * X, const double * Xlb, const double * Xub, int n);
double oracle_compute_alpha_bnd(const double * X, const double * Xlb, const double * Xub, const double * p, int Nprm);
double oracle_stream_uniform(uint64_t seed, uint64_t k, double scale);
int oracle_ga(int kind, const double * scalars, const long long * ints, const double * const * cols, long long m,
              double * X, const double * Xlb, const double * Xub, int Nparam, int Npop, int maxGenerations, double eliteFrac,
              double crossFrac, double eliteMutationFrac, double mutationSize, double eliteMutationSize,
              double NstaticGenerations, int stopAfter, const double * values, uint64_t n_values, uint64_t seed, double scale,
              double * f0_out, double * fOpt_out, double * Xpop_out, double * F_out, int * cross_idx, int * mut_idx,
              int * elite_idx, uint64_t * stream_pos_out);
}

struct pnol_ctx { std::string err; };
struct pnol_functor {
	pnol_functor_desc d;
	std::vector<std::vector<double> > owned;      // residual models: the data columns are copied (as the product copies them to the device)
};
struct pnol_ga {
	pnol_ctx * ctx; const pnol_functor * f; pnol_ga_params prm; int n; pnol_stream_desc stream;
	std::vector<double> lb, ub, values, x0, pop, F;
	int generations; double f0; pnol_ga_status st;
};

#define FARGS(f) (f)->d.kind, (f)->d.scalars, (f)->d.ints, (f)->d.columns, (f)->d.m

static int unavailable(pnol_ctx * ctx, const char * what)
{
	if (ctx) ctx->err = std::string(what) + ": not provided by the CPU stand-in of the host-logic tests (oracle/host_logic_device.cpp)";
	return PNOL_ERR_NO_FUNCTOR;
}

extern "C" {

int pnol_ctx_create(pnol_ctx ** ctx, int) { *ctx = new pnol_ctx; return PNOL_OK; }
void pnol_ctx_destroy(pnol_ctx * ctx) { delete ctx; }
const char * pnol_last_error(pnol_ctx * ctx) { return ctx ? ctx->err.c_str() : ""; }
int pnol_comm_rank(pnol_ctx *) { return 0; }
int pnol_comm_size(pnol_ctx *) { return 1; }
int pnol_comm_set_local(pnol_ctx *, int) { return 0; }
int pnol_ga_set_sharding(pnol_ctx *, int) { return PNOL_OK; }
int pnol_comm_broadcast(pnol_ctx *, double *, size_t, int) { return PNOL_OK; }
int pnol_comm_allreduce_sum(pnol_ctx *, double *, size_t) { return PNOL_OK; }
int pnol_comm_allgather(pnol_ctx *, const double * send, double * recv, size_t n) { std::memmove(recv, send, n * sizeof(double)); return PNOL_OK; }

// "device" memory is host memory here
int pnol_malloc(pnol_ctx *, void ** p, size_t bytes) { *p = std::malloc(bytes ? bytes : 1); return *p ? PNOL_OK : PNOL_ERR_CUDA; }
int pnol_free(pnol_ctx *, void * p) { std::free(p); return PNOL_OK; }
int pnol_memcpy(pnol_ctx *, void * dst, const void * src, size_t bytes) { std::memmove(dst, src, bytes); return PNOL_OK; }
int pnol_copy_start(pnol_ctx *, void * dst, const void * src, size_t bytes) { std::memmove(dst, src, bytes); return PNOL_OK; }      // (nothing runs beside the host here)
int pnol_copy_wait(pnol_ctx *) { return PNOL_OK; }
int pnol_memset(pnol_ctx *, void * p, int value, size_t bytes) { std::memset(p, value, bytes); return PNOL_OK; }

int pnol_functor_create(pnol_ctx * ctx, const pnol_functor_desc * desc, pnol_functor ** out)
{
	if (!desc || desc->kind < 1) return unavailable(ctx, "pnol_functor_create");
	pnol_functor * f = new pnol_functor;
	f->d = *desc;
	for (int c = 0; c < desc->n_columns && c < PNOL_MAX_COLUMNS; c++) {
		f->owned.emplace_back(desc->columns[c], desc->columns[c] + desc->m);
	}
	for (size_t c = 0; c < f->owned.size(); c++) f->d.columns[c] = f->owned[c].data();
	*out = f;
	return PNOL_OK;
}
void pnol_functor_destroy(pnol_functor * f) { delete f; }
long long pnol_functor_rows(const pnol_functor * f) { return f->d.kind >= 100 ? f->d.m : 0; }

int pnol_eval_batch(pnol_ctx *, const pnol_functor * f, const double * pts, long long B, int n, long long ld, const unsigned char * indicator,
                    double * f_out)
{
	oracle_eval_batch(FARGS(f), pts, B, n, ld, indicator, f_out);
	return PNOL_OK;
}
int pnol_fd_gradient(pnol_ctx *, const pnol_functor * f, const double * x, const double * dx, int n, double * g_out, double * f0_out)
{
	oracle_fd_gradient(FARGS(f), x, dx, n, g_out, f0_out);
	return PNOL_OK;
}
int pnol_eval_recur(pnol_ctx *, const pnol_functor * f, const double * xr, int nr, const double * const_x, const unsigned char * const_ind,
                    int nfull, double * f_out)
{
	*f_out = oracle_eval_recur(FARGS(f), xr, nr, const_x, const_ind, nfull);
	return PNOL_OK;
}
int pnol_fd_gradient_recur(pnol_ctx *, const pnol_functor * f, const double * xr, const double * dxr, int nr, const double * const_x,
                           const unsigned char * const_ind, int nfull, double * g_out, double * f0_out)
{
	oracle_fd_gradient_recur(FARGS(f), xr, dxr, nr, const_x, const_ind, nfull, g_out, f0_out);
	return PNOL_OK;
}
int pnol_fd_hessian(pnol_ctx *, const pnol_functor * f, const double * x, const double * dx, int n, double * B_out)
{
	oracle_fd_hessian(FARGS(f), x, dx, n, B_out);
	return PNOL_OK;
}
int pnol_alpha_pool(pnol_ctx *, const pnol_functor * f, const double * x, const double * p, int n, const double * alpha, int npool,
                    double dalpha, const unsigned char * eval_ind, const double * const_x, const unsigned char * const_ind, int nfull,
                    double * phi, double * dphi, int * bad_out)
{
	int bad = 0;
	oracle_alpha_pool(FARGS(f), x, p, n, alpha, npool, dalpha, eval_ind, const_x, const_ind, const_ind ? nfull : n, phi, dphi, &bad);
	if (bad_out) *bad_out = bad;
	return PNOL_OK;
}
int pnol_matvec_neg(pnol_ctx *, const double * D, const double * g, int n, double * p) { oracle_matvec_neg(D, g, n, p); return PNOL_OK; }
// both update modes map to the reference's literal two-product form: that is what the verbatim reference computes
int pnol_bfgs_update_hinv(pnol_ctx *, double * D, const double * g, const double * s, int n, int) { oracle_update_hinv(D, g, s, n); return PNOL_OK; }
// the reference inverts the FD Hessian with matrixInverse (LU, partial pivoting); column-wise LU solves give the same numbers
int pnol_spd_solve(pnol_ctx *, const double * A, const double * rhs, int n, double * x, int * info)
{
	oracle_lu_solve(A, rhs, n, x);
	if (info) *info = 0;
	return PNOL_OK;
}
// matrixInverse of the FD Hessian: one LU solve per unit vector (the shim's definition)
int pnol_lu_inverse(pnol_ctx *, const double * A, int n, double * Ainv, int * info)
{
	std::vector<double> e(n), col(n);
	for (int j = 0; j < n; j++) {
		for (int i = 0; i < n; i++) e[i] = (i == j) ? 1.0 : 0.0;
		oracle_lu_solve(A, e.data(), n, col.data());
		for (int i = 0; i < n; i++) Ainv[(size_t) i * n + j] = col[i];
	}
	if (info) *info = 0;
	return PNOL_OK;
}
int pnol_check_box_bounds(double * x, const double * xlb, const double * xub, int n, int * n_replaced)
{
	int c = oracle_check_box_bounds(x, xlb, xub, n);
	if (n_replaced) *n_replaced = c;
	return PNOL_OK;
}
double pnol_compute_alpha_bnd(const double * x, const double * xlb, const double * xub, const double * p, int n)
{
	return oracle_compute_alpha_bnd(x, xlb, xub, p, n);
}
double pnol_stream_uniform(uint64_t seed, uint64_t k, double scale) { return oracle_stream_uniform(seed, k, scale); }

// ---- Levenberg-Marquardt (Source/LevenbergMarquardtMPI.cpp:42-108): residuals, FD Jacobian, one iteration's device work ----
static double sumsq_sequential(const double * F, long long m)          // vector2Norm(F)^2 before the sqrt: one running sum
{
	double s = 0;
	for (long long k = 0; k < m; k++) s = s + F[k] * F[k];
	return s;
}
int pnol_residual_eval(pnol_ctx * ctx, const pnol_functor * f, const double * x, int n, double * F, double * sumsq_out)
{
	if (f->d.kind < 100) return unavailable(ctx, "pnol_residual_eval on a scalar functor");
	oracle_residual(FARGS(f), x, n, F);
	if (sumsq_out) *sumsq_out = sumsq_sequential(F, f->d.m);
	return PNOL_OK;
}
int pnol_fd_jacobian(pnol_ctx * ctx, const pnol_functor * f, const double * x, const double * dx, int n, double * J, double * F, int)
{
	if (f->d.kind < 100) return unavailable(ctx, "pnol_fd_jacobian on a scalar functor");
	oracle_fd_jacobian(FARGS(f), x, dx, n, J, F);
	return PNOL_OK;
}
int pnol_lm_step(pnol_ctx * ctx, const pnol_functor * f, const double * x, const double * dx, int n, double * J, const double * F,
                 double * Ftrial, double lambda, int, int reuse_jtj, double * JTJ, double * sigma_out, double * x_trial_out,
                 double * sumsq_trial_out, int * spd_info_out)
{
	if (f->d.kind < 100) return unavailable(ctx, "pnol_lm_step on a scalar functor");
	const long long m = f->d.m;
	const size_t nn = (size_t) n * n;
	std::vector<double> A(nn), rhs(n), sigma(n), xt(n), Jown;
	if (!reuse_jtj) {
		if (!J) { Jown.resize((size_t) m * n); J = Jown.data(); }       // the caller keeps no J
		oracle_fd_jacobian(FARGS(f), x, dx, n, J, nullptr);
		oracle_lm_normal_eq(J, F, m, n, lambda, JTJ, A.data(), rhs.data());
		std::memcpy(JTJ + nn, rhs.data(), (size_t) n * sizeof(double));  // J^T J followed by -J^T F, as the product keeps them
	} else {
		for (int i = 0; i < n; i++)
			for (int j = 0; j < n; j++) A[(size_t) i * n + j] = (i == j) ? (1 + lambda) * JTJ[(size_t) i * n + j] : JTJ[(size_t) i * n + j];
		std::memcpy(rhs.data(), JTJ + nn, (size_t) n * sizeof(double));
	}
	oracle_lu_solve(A.data(), rhs.data(), n, sigma.data());              // luSolve (:88)
	for (int i = 0; i < n; i++) xt[i] = x[i] + sigma[i];                 // (:97-100)
	oracle_residual(FARGS(f), xt.data(), n, Ftrial);
	if (sumsq_trial_out) *sumsq_trial_out = sumsq_sequential(Ftrial, m);
	if (sigma_out) std::memcpy(sigma_out, sigma.data(), (size_t) n * sizeof(double));
	if (x_trial_out) std::memcpy(x_trial_out, xt.data(), (size_t) n * sizeof(double));
	if (spd_info_out) *spd_info_out = 0;
	return PNOL_OK;
}

// the while loop of Source/LevenbergMarquardtMPI.cpp:55-141 behind one call (the product runs it on device-resident state with the
// accept / reject rule in a kernel): here pnol_lm_step pass by pass plus the rule as the reference writes it
static int g_lm_stopped = 0;
static double g_lm_xdiff = 0.0;
int pnol_lm_iterate(pnol_ctx * ctx, const pnol_functor * f, double * x, const double * dx, int n, double * J, double * F, double * Ftrial,
                    double * JTJ, double * lambda_inout, double * chisq_inout, double lambda_factor, double x_min_diff, int iterations,
                    int jac_mode, int * accepted_out, int * rejected_out, int * swapped_out)
{
	if (f->d.kind < 100) return unavailable(ctx, "pnol_lm_iterate on a scalar functor");
	const long long m = f->d.m;
	std::vector<double> sigma(n), xt(n);
	double lambda = *lambda_inout, chisq = *chisq_inout;
	int acc = 0, rej = 0;
	g_lm_stopped = 0;
	g_lm_xdiff = 0.0;
	for (int it = 0; it < iterations && !g_lm_stopped; it++) {
		double ss = 0;
		int info = 0;
		int st = pnol_lm_step(ctx, f, x, dx, n, J, F, Ftrial, lambda, jac_mode, 0, JTJ, sigma.data(), xt.data(), &ss, &info);
		if (st != PNOL_OK) return st;
		const double root = std::sqrt(ss);
		const double chi = root * root;                              // pow(vector2Norm(F),2)  (:108)
		if (chi >= chisq || chi != chi) { lambda = lambda * lambda_factor; rej++; continue; }      // (:110-129)
		lambda = lambda / lambda_factor;                             // (:132-141)
		chisq = chi;
		for (int i = 0; i < n; i++) x[i] = xt[i];
		std::memcpy(F, Ftrial, (size_t) m * sizeof(double));
		acc++;
		double s2 = 0;
		for (int i = 0; i < n; i++) s2 = s2 + sigma[i] * sigma[i];
		g_lm_xdiff = std::sqrt(s2);
		if (x_min_diff > 0 && g_lm_xdiff < x_min_diff) g_lm_stopped = 1;
	}
	*lambda_inout = lambda;
	*chisq_inout = chisq;
	if (accepted_out) *accepted_out = acc;
	if (rejected_out) *rejected_out = rej;
	if (swapped_out) *swapped_out = 0;
	return PNOL_OK;
}
int pnol_lm_last_run(pnol_ctx *, int * stopped_out, double * xdiff_out)
{
	if (stopped_out) *stopped_out = g_lm_stopped;
	if (xdiff_out) *xdiff_out = g_lm_xdiff;
	return PNOL_OK;
}

// ---- genetic algorithm (Source/GeneticAlgorithmMPI.cpp:12-276): the oracle restates the whole loop (oracle_ga); the state machine
// of the C-ABI is answered by replaying it up to the requested generation (deterministic: same stream, same start) ----
int pnol_ga_create(pnol_ctx * ctx, const pnol_functor * f, const pnol_ga_params * params, int n, const double * xlb, const double * xub,
                   const pnol_stream_desc * stream, pnol_ga ** out)
{
	if (!f || !params || !stream || f->d.kind >= 100 || params->npop < 2) return unavailable(ctx, "pnol_ga_create (bad arguments)");
	pnol_ga * ga = new pnol_ga;
	ga->ctx = ctx; ga->f = f; ga->prm = *params; ga->n = n; ga->stream = *stream;
	ga->lb.assign(xlb, xlb + n); ga->ub.assign(xub, xub + n);
	if (stream->values) { ga->values.assign(stream->values, stream->values + stream->n_values); }
	ga->generations = 0;
	std::memset(&ga->st, 0, sizeof ga->st);
	*out = ga;
	return PNOL_OK;
}
void pnol_ga_destroy(pnol_ga * ga) { delete ga; }
static int ga_replay(pnol_ga * ga)
{
	const pnol_ga_params & p = ga->prm;
	std::vector<double> X(ga->x0);
	ga->pop.assign((size_t) p.npop * ga->n, 0.0);
	ga->F.assign(p.npop, 0.0);
	double f0 = 0, fOpt = 0;
	uint64_t pos = 0;
	int it = oracle_ga(FARGS(ga->f), X.data(), ga->lb.data(), ga->ub.data(), ga->n, p.npop, p.max_generations, p.elite_frac, p.cross_frac,
	                   p.elite_mutation_frac, p.mutation_size, p.elite_mutation_size, p.n_static_generations, ga->generations,
	                   ga->values.empty() ? nullptr : ga->values.data(), (uint64_t) ga->values.size(), ga->stream.seed, ga->stream.scale,
	                   &f0, &fOpt, ga->pop.data(), ga->F.data(), nullptr, nullptr, nullptr, &pos);
	if (it == -1) { ga->ctx->err = "GA fractions leave no room for random children"; return PNOL_ERR_INVALID; }
	if (it == -2) { ga->ctx->err = "random stream exhausted"; return PNOL_ERR_STREAM; }
	ga->f0 = f0;
	ga->st.generation = it;
	ga->st.stopped = it < ga->generations ? 1 : 0;        // the static-generation test ended the loop before the requested generation
	ga->st.f_best = fOpt;
	ga->st.stream_pos = pos;
	ga->st.n_elite = (int) std::ceil(p.elite_frac * p.npop);
	ga->st.n_elite_mut = (int) std::ceil(p.elite_mutation_frac * p.npop);
	ga->st.n_cross = (int) std::ceil(p.cross_frac * p.npop);
	ga->st.n_rand = p.npop - ga->st.n_elite - ga->st.n_elite_mut - ga->st.n_cross;
	return PNOL_OK;
}
int pnol_ga_init(pnol_ga * ga, const double * x0, double * f0_out)
{
	ga->x0.assign(x0, x0 + ga->n);
	ga->generations = 0;
	int st = ga_replay(ga);
	if (st == PNOL_OK && f0_out) *f0_out = ga->f0;
	return st;
}
int pnol_ga_generation(pnol_ga * ga) { ga->generations++; return ga_replay(ga); }
int pnol_ga_peer_mode(pnol_ga *) { return 0; }
int pnol_ga_status_get(pnol_ga * ga, pnol_ga_status * st) { *st = ga->st; return PNOL_OK; }
int pnol_ga_get_population(pnol_ga * ga, double * xpop, double * F)
{
	if (xpop) std::memcpy(xpop, ga->pop.data(), ga->pop.size() * sizeof(double));
	if (F) std::memcpy(F, ga->F.data(), ga->F.size() * sizeof(double));
	return PNOL_OK;
}
// the stand-alone stage entry points are exercised against the oracle on the GPU (tests/test_gpu_ga.py); not answered here
int pnol_ga_pop_sort(pnol_ctx * ctx, double *, double *, long long, int) { return unavailable(ctx, "pnol_ga_pop_sort"); }
int pnol_ga_check_bounds(pnol_ctx * ctx, double *, long long, int, const double *, const double *, unsigned char *, const pnol_stream_desc *,
                         uint64_t *) { return unavailable(ctx, "pnol_ga_check_bounds"); }
int pnol_ga_check_identical(pnol_ctx * ctx, double *, long long, int, const double *, const double *, unsigned char *, const pnol_stream_desc *,
                            uint64_t *) { return unavailable(ctx, "pnol_ga_check_identical"); }

} // extern "C"
