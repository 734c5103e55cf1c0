/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
// Runtime.cpp -- pnol::Runtime, DeviceFunctor, DeviceArray (host C++ above the C-ABI).
#include "pnol/Runtime.hpp"

#include <chrono>
#include <cstdlib>
#include <cstring>

namespace pnol {

Runtime::Runtime() : ctx_(nullptr), owned_(false), poolWidth_(8), haveStream_(false), hinvMode_(PNOL_HINV_RANK2), jacMode_(PNOL_JAC_AUTO), jacCache_(true)
{
	std::memset(&stream_, 0, sizeof stream_);
	const char * pw = std::getenv("PNOL_POOL_WIDTH");
	if (pw && std::atoi(pw) > 0) poolWidth_ = std::atoi(pw);
	const char * jc = std::getenv("PNOL_LM_JACOBIAN_CACHE");
	if (jc) jacCache_ = std::atoi(jc) != 0;
}

Runtime::~Runtime() { reset(); }

Runtime & Runtime::instance()
{
	static Runtime rt;
	return rt;
}

pnol_ctx * Runtime::ctx()
{
	if (!ctx_) {
		int dev = 0;
		const char * e = std::getenv("PNOL_DEVICE");
		if (!e) e = std::getenv("LOCAL_RANK");
		if (e) dev = std::atoi(e);
		int st = pnol_ctx_create(&ctx_, dev);
		if (st != PNOL_OK) {
			ctx_ = nullptr;
			throw Error(st, "pnol: cannot create a context on CUDA device " + std::to_string(dev) +
			                    " (a B200 is required; this library has no CPU fallback)");
		}
		owned_ = true;
	}
	return ctx_;
}

void Runtime::attach(pnol_ctx * ctx)
{
	reset();
	ctx_ = ctx;
	owned_ = false;
}

void Runtime::reset()
{
	if (ctx_ && owned_) pnol_ctx_destroy(ctx_);
	ctx_ = nullptr;
	owned_ = false;
}

pnol_stream_desc Runtime::defaultStream(double scale)
{
	if (haveStream_) return stream_;
	// one clock seed per process, drawn at the first use (the reference re-seeds with srand(time(0)) in every findMin call)
	static const uint64_t clockSeed = (uint64_t) std::chrono::system_clock::now().time_since_epoch().count();
	uint64_t seed = clockSeed;
	pnol_ctx * c = ctx();
	if (pnol_comm_size(c) > 1) {
		// rank 0's seed for everybody: two doubles carrying 32 bits each (exact), one broadcast
		double halves[2] = {(double) (uint32_t) (seed >> 32), (double) (uint32_t) (seed & 0xFFFFFFFFu)};
		check(pnol_comm_broadcast(c, halves, 2, 0));
		seed = ((uint64_t) (uint32_t) halves[0] << 32) | (uint64_t) (uint32_t) halves[1];
	}
	pnol_stream_desc s;
	s.values = nullptr; s.n_values = 0;
	s.seed = seed;
	s.scale = scale;
	return s;
}

void Runtime::check(int status) const
{
	if (status == PNOL_OK) return;
	std::string msg = ctx_ ? pnol_last_error(ctx_) : "";
	throw Error(status, "pnol error " + std::to_string(status) + ": " + msg);
}

pnol_functor * DeviceFunctor::get(int kind, const std::vector<double> & scalars, const std::vector<long long> & ints,
                                  const std::vector<const double *> & columns, long long m)
{
	if (f_) return f_;
	Runtime & rt = Runtime::instance();
	pnol_functor_desc d;
	std::memset(&d, 0, sizeof d);
	d.kind = kind;
	for (size_t i = 0; i < scalars.size() && i < PNOL_MAX_SCALARS; i++) d.scalars[i] = scalars[i];
	for (size_t i = 0; i < ints.size() && i < PNOL_MAX_INTS; i++) d.ints[i] = ints[i];
	d.n_columns = (int) columns.size();
	for (size_t i = 0; i < columns.size() && i < PNOL_MAX_COLUMNS; i++) d.columns[i] = columns[i];
	d.m = m;
	rt.check(pnol_functor_create(rt.ctx(), &d, &f_));
	return f_;
}

void DeviceFunctor::release()
{
	if (f_) pnol_functor_destroy(f_);
	f_ = nullptr;
}

void DeviceArray::resize(size_t n)
{
	if (n == n_ && p_) return;
	free();
	Runtime & rt = Runtime::instance();
	void * p = nullptr;
	rt.check(pnol_malloc(rt.ctx(), &p, n * sizeof(double)));
	p_ = (double *) p;
	n_ = n;
}

void DeviceArray::free()
{
	if (p_) {
		Runtime & rt = Runtime::instance();
		pnol_free(rt.ctx(), p_);
	}
	p_ = nullptr;
	n_ = 0;
}

void DeviceArray::upload(const double * host, size_t n)
{
	Runtime & rt = Runtime::instance();
	rt.check(pnol_memcpy(rt.ctx(), p_, host, n * sizeof(double)));
}

void DeviceArray::download(double * host, size_t n) const
{
	Runtime & rt = Runtime::instance();
	rt.check(pnol_memcpy(rt.ctx(), host, p_, n * sizeof(double)));
}

} // namespace pnol
