// FP64 pipe microbenchmark for sm_100a: DFMA latency / throughput as a function of warps per SM and independent chains per
// warp, and whether DMMA (FP64 tensor) and DFMA (FP64 ALU) overlap. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -o tools/ubench/fp64_ubench tools/ubench/fp64_ubench.cu ; run on the GPU box. Tuning aid only (not part of the library).
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(int iters, double * sink, long long * cycles)
{
	double a[ILP];
#pragma unroll
	for (int i = 0; i < ILP; i++) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
	const double b = 1.0000001, c = 1e-9;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int r = 0; r < 8; r++)
#pragma unroll
			for (int i = 0; i < ILP; i++) a[i] = fma(a[i], b, c);
	}
	long long t1 = clock64();
	double s = 0;
#pragma unroll
	for (int i = 0; i < ILP; i++) s += a[i];
	if (s == 123.456) sink[0] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

__device__ __forceinline__ void dmma(double & d0, double & d1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// per iteration: NM independent DMMA + NF independent DFMA (x8 unroll)
template <int NM, int NF>
__global__ void mix_kernel(int iters, double * sink, long long * cycles)
{
	double acc[NM > 0 ? NM : 1][2], f[NF > 0 ? NF : 1];
	for (int i = 0; i < NM; i++) { acc[i][0] = 0; acc[i][1] = 0; }
	for (int i = 0; i < NF; i++) f[i] = 1.0 + 1e-9 * (threadIdx.x + i);
	const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x, c = 1e-9;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int r = 0; r < 4; r++) {
#pragma unroll
			for (int i = 0; i < NM; i++) dmma(acc[i][0], acc[i][1], a, b);
#pragma unroll
			for (int i = 0; i < NF; i++) f[i] = fma(f[i], b, c);
		}
	}
	long long t1 = clock64();
	double s = 0;
	for (int i = 0; i < NM; i++) s += acc[i][0] + acc[i][1];
	for (int i = 0; i < NF; i++) s += f[i];
	if (s == 123.456) sink[0] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

// OP 0: DADD, 1: DMUL, 2: DFMA, 3: DSETP+predicated-select-free accumulate, 4: MUFU.RCP64H, 5: SHFL (xor, f64 = 2 SHFL)
template <int OP, int ILP>
__global__ void op_kernel(int iters, double * sink, long long * cycles)
{
	double a[ILP];
	int cnt = 0;
#pragma unroll
	for (int i = 0; i < ILP; i++) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
	const double b = 1.0000001, c = 1e-9;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int r = 0; r < 8; r++)
#pragma unroll
			for (int i = 0; i < ILP; i++) {
				if (OP == 0) a[i] = a[i] + c;
				if (OP == 1) a[i] = a[i] * b;
				if (OP == 2) a[i] = fma(a[i], b, c);
				if (OP == 3) { cnt += (a[i] < b + it) ? 1 : 0; }
				if (OP == 4) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a[i])); a[i] = y; }
				if (OP == 5) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (r & 15));
			}
	}
	long long t1 = clock64();
	double s = cnt;
#pragma unroll
	for (int i = 0; i < ILP; i++) s += a[i];
	if (s == 123.456) sink[0] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

// per iteration (x8 unroll): 4 independent DFMA + NI independent integer ops (LOP3/IADD chain per slot) + NS FP32 FFMA
template <int NI, int NS>
__global__ void issue_kernel(int iters, double * sink, long long * cycles)
{
	double f[4];
	unsigned u[NI > 0 ? NI : 1];
	float g[NS > 0 ? NS : 1];
	for (int i = 0; i < 4; i++) f[i] = 1.0 + 1e-9 * (threadIdx.x + i);
	for (int i = 0; i < NI; i++) u[i] = threadIdx.x * 2654435761u + i;
	for (int i = 0; i < NS; i++) g[i] = 1.0f + 1e-3f * (threadIdx.x + i);
	const double b = 1.0000001, c = 1e-9;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int r = 0; r < 8; r++) {
#pragma unroll
			for (int i = 0; i < 4; i++) f[i] = fma(f[i], b, c);
#pragma unroll
			for (int i = 0; i < NI; i++) u[i] = (u[i] ^ (u[i] >> 7)) + 0x9E3779B9u;     // 2 int instr (LOP3/SHF + IADD)
#pragma unroll
			for (int i = 0; i < NS; i++) g[i] = fmaf(g[i], 1.0001f, 1e-3f);
		}
	}
	long long t1 = clock64();
	double s = 0;
	for (int i = 0; i < 4; i++) s += f[i];
	for (int i = 0; i < NI; i++) s += u[i];
	for (int i = 0; i < NS; i++) s += g[i];
	if (s == 123.456) sink[0] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

// per iteration (x8 unroll): 4 independent DFMA + NL shared-memory loads (LDS.64, lane-dependent address) feeding nothing critical
template <int NL>
__global__ void lds_kernel(int iters, double * sink, long long * cycles)
{
	__shared__ double tab[64];
	if (threadIdx.x < 64) tab[threadIdx.x] = 1.0 + threadIdx.x;
	__syncthreads();
	double f[4], acc = 0;
	for (int i = 0; i < 4; i++) f[i] = 1.0 + 1e-9 * (threadIdx.x + i);
	const double b = 1.0000001, c = 1e-9;
	int idx = threadIdx.x & 1;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int r = 0; r < 8; r++) {
#pragma unroll
			for (int i = 0; i < 4; i++) f[i] = fma(f[i], b, c);
#pragma unroll
			for (int i = 0; i < NL; i++) {
				double v;
				asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned) __cvta_generic_to_shared(&tab[(idx * 8 + i + r) & 63])));
				acc = __longlong_as_double(__double_as_longlong(acc) ^ __double_as_longlong(v));
			}
		}
		idx ^= (it & 1);
	}
	long long t1 = clock64();
	double s = acc;
	for (int i = 0; i < 4; i++) s += f[i];
	if (s == 123.456) sink[0] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

// per iteration (x4 unroll): 8 independent DMMA + NI integer-op groups (SHF/LOP3/IADD, 3 instructions each) + NL LDS.64
template <int NI, int NL>
__global__ void dmma_other_kernel(int iters, double * sink, long long * cycles)
{
	__shared__ double tab[512];
	for (int i = threadIdx.x; i < 512; i += blockDim.x) tab[i] = 1.0 + i;
	__syncthreads();
	double acc[8][2];
	unsigned u[NI > 0 ? NI : 1];
	double l[NL > 0 ? NL : 1];
	for (int i = 0; i < 8; i++) { acc[i][0] = 0; acc[i][1] = 0; }
	for (int i = 0; i < NI; i++) u[i] = threadIdx.x * 2654435761u + i;
	for (int i = 0; i < NL; i++) l[i] = 0;
	const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
	const unsigned base = (unsigned) __cvta_generic_to_shared(&tab[threadIdx.x & 31]);
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int r = 0; r < 4; r++) {
#pragma unroll
			for (int i = 0; i < 8; i++) {
				dmma(acc[i][0], acc[i][1], a, b);
				if (i < NI) u[i] = (u[i] ^ (u[i] >> 7)) + 0x9E3779B9u;
				if (i < NL) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(l[i]) : "r"(base + 256 * (i + r)));
			}
		}
	}
	long long t1 = clock64();
	double s = 0;
	for (int i = 0; i < 8; i++) s += acc[i][0] + acc[i][1];
	for (int i = 0; i < NI; i++) s += u[i];
	for (int i = 0; i < NL; i++) s += l[i];
	if (s == 123.456) sink[0] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <class K> void run(const char * name, K kern, int warps, int iters, double inst_per_iter_per_warp, int sms)
{
	double * sink; long long * cyc;
	cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
	kern<<<sms, warps * 32>>>(iters, sink, cyc);
	cudaDeviceSynchronize();
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	kern<<<sms, warps * 32>>>(iters, sink, cyc);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
	double inst_per_sm = inst_per_iter_per_warp * iters * warps;
	printf("%-28s warps/SM=%2d  cycles=%9lld  warp-inst/cycle/SM=%.3f  cycles/inst/warp=%.2f  (%.3f ms) err=%s\n", name, warps, c, inst_per_sm / c,
	       c / (inst_per_iter_per_warp * iters), ms, cudaGetErrorString(cudaGetLastError()));
	cudaFree(sink); cudaFree(cyc);
}

int main()
{
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	int sms = p.multiProcessorCount;
	printf("device %s, %d SMs\n", p.name, sms);
	const int it = 4000;
	for (int w : {1, 4, 8, 16, 32}) {
		run("DFMA ILP=1", dfma_kernel<1>, w, it, 8.0 * 1, sms);
		run("DFMA ILP=2", dfma_kernel<2>, w, it, 8.0 * 2, sms);
		run("DFMA ILP=4", dfma_kernel<4>, w, it, 8.0 * 4, sms);
		run("DFMA ILP=8", dfma_kernel<8>, w, it, 8.0 * 8, sms);
	}
	for (int w : {1, 16}) {
		run("DADD ILP=1", op_kernel<0, 1>, w, it, 8.0, sms);
		run("DADD ILP=4", op_kernel<0, 4>, w, it, 32.0, sms);
		run("DMUL ILP=1", op_kernel<1, 1>, w, it, 8.0, sms);
		run("DMUL ILP=4", op_kernel<1, 4>, w, it, 32.0, sms);
		run("DSETP ILP=4", op_kernel<3, 4>, w, it, 32.0, sms);
		run("MUFU.RCP64H ILP=1", op_kernel<4, 1>, w, it, 8.0, sms);
		run("MUFU.RCP64H ILP=4", op_kernel<4, 4>, w, it, 32.0, sms);
		run("SHFL.f64 ILP=1", op_kernel<5, 1>, w, it, 8.0, sms);
		run("SHFL.f64 ILP=4", op_kernel<5, 4>, w, it, 32.0, sms);
	}
	// inst count below = the 4 DFMA only, so "cycles/inst/warp" shows what the extra instructions cost the DFMA stream
	for (int w : {4, 8, 16}) {
		run("4 DFMA + 0 other", issue_kernel<0, 0>, w, it, 8.0 * 4, sms);
		run("4 DFMA + 2x(2 INT)", issue_kernel<2, 0>, w, it, 8.0 * 4, sms);
		run("4 DFMA + 4x(2 INT)", issue_kernel<4, 0>, w, it, 8.0 * 4, sms);
		run("4 DFMA + 4 FFMA", issue_kernel<0, 4>, w, it, 8.0 * 4, sms);
		run("4 DFMA + 8 FFMA", issue_kernel<0, 8>, w, it, 8.0 * 4, sms);
	}
	for (int w : {16}) {
		run("4 DFMA + 4 LDS.64", lds_kernel<4>, w, it, 8.0 * 4, sms);
		run("4 DFMA + 8 LDS.64", lds_kernel<8>, w, it, 8.0 * 4, sms);
	}
	// inst count = the 8 DMMA only: cycles/inst/warp x (4 / warps per scheduler) = 16 means the tensor pipe is saturated
	for (int w : {16}) {
		run("8 DMMA + 0 other", dmma_other_kernel<0, 0>, w, it, 4.0 * 8, sms);
		run("8 DMMA + 4x3 INT", dmma_other_kernel<4, 0>, w, it, 4.0 * 8, sms);
		run("8 DMMA + 8x3 INT", dmma_other_kernel<8, 0>, w, it, 4.0 * 8, sms);
		run("8 DMMA + 4 LDS", dmma_other_kernel<0, 4>, w, it, 4.0 * 8, sms);
		run("8 DMMA + 8 LDS", dmma_other_kernel<0, 8>, w, it, 4.0 * 8, sms);
		run("8 DMMA + 8 LDS + 8x3 INT", dmma_other_kernel<8, 8>, w, it, 4.0 * 8, sms);
	}
	for (int w : {4}) {
		run("DMMA x8 only", mix_kernel<8, 0>, w, it, 4.0 * 8, sms);
		run("DFMA x8 only", mix_kernel<0, 8>, w, it, 4.0 * 8, sms);
		run("DMMA x8 + DFMA x8", mix_kernel<8, 8>, w, it, 4.0 * 16, sms);
		run("DMMA x8 + DFMA x2", mix_kernel<8, 2>, w, it, 4.0 * 10, sms);
		run("DMMA x8 + DFMA x4", mix_kernel<8, 4>, w, it, 4.0 * 12, sms);
	}
	return 0;
}
