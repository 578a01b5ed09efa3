// Micro-test (not part of the product): are FFMA2-based packed mul / add / sub bit-identical to the scalar
// round-to-nearest FMUL / FADD for arbitrary operands (random bit patterns incl. zeros, denormals, inf, NaN)?
//   mul2(a, b) = ffma2(a, b, -0)     add2(a, b) = ffma2(a, 1, b)     sub2(a, b) = ffma2(b, -1, a)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t rng(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
__device__ __forceinline__ float pick(uint32_t& s)
{
	const uint32_t r = rng(s);
	switch (r & 15u)
	{
	case 0: return 0.f; case 1: return -0.f; case 2: return __int_as_float(0x7f800000); case 3: return __int_as_float(0xff800000);
	case 4: return __int_as_float(0x7fc00000); case 5: return __int_as_float((rng(s) & 0x807fffffu));            // denormal
	case 6: return 1.f; case 7: return -1.f;
	case 8: case 9: case 10: return __int_as_float(rng(s));                                                      // anything
	default: return __int_as_float((rng(s) & 0x807fffffu) | ((100u + (rng(s) % 56u)) << 23));                    // moderate exponents
	}
}
__device__ __forceinline__ bool same(float a, float b)
{
	const uint32_t x = __float_as_uint(a), y = __float_as_uint(b);
	const bool nan_a = (x & 0x7fffffffu) > 0x7f800000u, nan_b = (y & 0x7fffffffu) > 0x7f800000u;
	return (nan_a && nan_b) || x == y;
}

// constants come in as kernel parameters so that neither NVVM nor ptxas can fold fma(a, b, -0) into a mul
// (ptxas 12.9 then contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false)
__global__ void test(unsigned long long* bad, int iters, float2 neg0, float2 one, float2 mone)
{
	uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
	unsigned long long n = 0, n2 = 0;
	for (int i = 0; i < iters; ++i)
	{
		const float2 a = make_float2(pick(s), pick(s)), b = make_float2(pick(s), pick(s));
		const float2 m = __ffma2_rn(a, b, neg0), p = __ffma2_rn(a, one, b), d = __ffma2_rn(b, mone, a);
		n += !same(m.x, __fmul_rn(a.x, b.x)) + !same(m.y, __fmul_rn(a.y, b.y));
		n += !same(p.x, __fadd_rn(a.x, b.x)) + !same(p.y, __fadd_rn(a.y, b.y));
		n += !same(d.x, __fsub_rn(a.x, b.x)) + !same(d.y, __fsub_rn(a.y, b.y));
		// chained: (a*b + c) - a*c with every operation rounded on its own
		const float2 c = make_float2(pick(s), pick(s));
		const float2 ch = __ffma2_rn(__ffma2_rn(a, c, neg0), mone, __ffma2_rn(__ffma2_rn(a, b, neg0), one, c));
		n += !same(ch.x, __fsub_rn(__fadd_rn(__fmul_rn(a.x, b.x), c.x), __fmul_rn(a.x, c.x)));
		n += !same(ch.y, __fsub_rn(__fadd_rn(__fmul_rn(a.y, b.y), c.y), __fmul_rn(a.y, c.y)));
		// the unguarded intrinsics, for the record (expected to mismatch: ptxas fuses them)
		const float2 un = __fadd2_rn(__fmul2_rn(a, b), c);
		n2 += !same(un.x, __fadd_rn(__fmul_rn(a.x, b.x), c.x)) + !same(un.y, __fadd_rn(__fmul_rn(a.y, b.y), c.y));
	}
	if (n) atomicAdd(bad, n);
	if (n2) atomicAdd(bad + 1, n2);
}

int main()
{
	unsigned long long* d; cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
	test<<<148 * 4, 256>>>(d, 4096, make_float2(-0.f, -0.f), make_float2(1.f, 1.f), make_float2(-1.f, -1.f));
	unsigned long long h[2] = { 1, 1 }; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
	printf("operand pairs tested: %llu, mismatches with FFMA2 + runtime constants: %llu, with __fmul2_rn/__fadd2_rn chained: %llu (%s)\n",
	       148ull * 4 * 256 * 4096 * 2, h[0], h[1], cudaGetErrorString(cudaGetLastError()));
	return h[0] != 0;
}
