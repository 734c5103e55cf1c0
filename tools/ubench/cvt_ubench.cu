// Throughput of the 64-bit conversions next to the splitmix64 arithmetic of the GA's counter stream (tuning aid, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/cvt_ubench tools/ubench/cvt_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long mix(unsigned long long k)
{
	unsigned long long z = 12345ULL + (k + 1ULL) * 0x9E3779B97F4A7C15ULL;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}
template <int MODE> __global__ void __launch_bounds__(256) k(unsigned long long n, double * out)
{
	unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x, stride = (unsigned long long) gridDim.x * blockDim.x;
	double acc = 0;
	unsigned long long iacc = 0;
	for (; i < n; i += stride) {
		const unsigned long long m = mix(i) >> 11;
		if (MODE == 0) iacc += m;                                            // integer work only
		if (MODE == 1) acc += (double) m;                                    // I2F.F64.U64
		if (MODE == 2) {                                                     // exact magic-number conversion (2 DADD)
			double d = __longlong_as_double((long long) (0x4330000000000000ULL | (m & 0xFFFFFFFFFFFFFULL))) - 4503599627370496.0;
			d += (m >> 52) ? 4503599627370496.0 : 0.0;
			acc += d;
		}
		if (MODE == 3) { const double x = (double) (unsigned) m * 0.37; iacc += (int) round(x); }       // I2F.F64.U32 + round + F2I
		if (MODE == 4) {                                                     // the same through magic adds
			const double x = (double) (unsigned) m * 0.37;
			const double t = __dadd_rz(__dadd_rz(x, 0.5), 4503599627370496.0);
			iacc += (unsigned) __double2loint(t);
		}
	}
	if (acc == 1.2345 || iacc == 77) out[0] = acc + (double) iacc;
}
int main()
{
	double * out; cudaMalloc(&out, 8);
	const unsigned long long n = 1ULL << 28;
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	const char * names[] = {"splitmix only", "+ I2F.F64.U64", "+ magic u64->f64", "+ I2F.U32, round, F2I", "+ I2F.U32, magic round"};
	for (int mode = 0; mode < 5; mode++)
		for (int rep = 0; rep < 2; rep++) {
			cudaEventRecord(a);
			if (mode == 0) k<0><<<148 * 8, 256>>>(n, out);
			if (mode == 1) k<1><<<148 * 8, 256>>>(n, out);
			if (mode == 2) k<2><<<148 * 8, 256>>>(n, out);
			if (mode == 3) k<3><<<148 * 8, 256>>>(n, out);
			if (mode == 4) k<4><<<148 * 8, 256>>>(n, out);
			cudaEventRecord(b); cudaEventSynchronize(b);
			float ms; cudaEventElapsedTime(&ms, a, b);
			if (rep) printf("%-28s %8.3f ms  %7.2f G draws/s\n", names[mode], ms, n / ms / 1e6);
		}
	return 0;
}
