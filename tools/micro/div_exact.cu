// Micro-test (not part of the product): rt::normalize_and_invert's shared-reciprocal divisions against the plain
// intrinsics (__fsqrt_rn / __fdiv_rn / __frcp_rn), bit for bit, over random vectors: magnitudes spread over the whole
// exponent range (so that both the fast range and its edges and the fallback are exercised), components down to zero,
// denormals, infinities and NaN included.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include "../../gp1_raytracer_2223_b200/csrc/rt_device.cuh"

__device__ __forceinline__ uint32_t rng(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
__device__ __forceinline__ float pick(uint32_t& s, uint32_t scale_exp)
{
	const uint32_t r = rng(s);
	switch (r & 31u)
	{
	case 0: return 0.f; case 1: return -0.f; case 2: return __int_as_float(0x7f800000); case 3: return __int_as_float(0x7fc00000);
	case 4: return __int_as_float(rng(s) & 0x807fffffu);                                                        // denormal
	case 5: return __int_as_float(rng(s));                                                                      // anything
	case 6: case 7: return __int_as_float((rng(s) & 0x807fffffu) | ((scale_exp - (rng(s) % 70u)) << 23));       // far below the vector's scale
	default: return __int_as_float((rng(s) & 0x807fffffu) | ((scale_exp - (rng(s) % 12u)) << 23));              // within 2^-12 of it
	}
}
__device__ __forceinline__ bool same(float a, float b)
{
	const uint32_t x = __float_as_uint(a), y = __float_as_uint(b);
	const bool nan_a = (x & 0x7fffffffu) > 0x7f800000u, nan_b = (y & 0x7fffffffu) > 0x7f800000u;
	return (nan_a && nan_b) || x == y;
}

__global__ void test(unsigned long long* out, int iters)
{
	uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 777u;
	unsigned long long bad = 0, fast = 0;
	for (int i = 0; i < iters; ++i)
	{
		// scale exponent: mostly around 1 (what the renderer sees), often near the fast range's edges (2^-40, 2^40), sometimes anywhere
		uint32_t e;
		const uint32_t k = rng(s) & 7u;
		if (k < 3) e = 127u + (rng(s) % 16u) - 8u;
		else if (k < 5) e = 127u - 40u + (rng(s) % 6u) - 2u;
		else if (k < 7) e = 127u + 40u + (rng(s) % 6u) - 4u;
		else e = 72u + (rng(s) % 180u);
		rt::V3 a = rt::v3(pick(s, e), pick(s, e), pick(s, e));
		const rt::V3 a0 = a;
		rt::V3 inv; bool finite;
		const float m = rt::normalize_and_invert(a, inv, finite);
		const float m_ref = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(a0.x, a0.x), __fmul_rn(a0.y, a0.y)), __fmul_rn(a0.z, a0.z)));
		const float qx = __fdiv_rn(a0.x, m_ref), qy = __fdiv_rn(a0.y, m_ref), qz = __fdiv_rn(a0.z, m_ref);
		const float ix = __frcp_rn(qx), iy = __frcp_rn(qy), iz = __frcp_rn(qz);
		const bool finite_ref = (fabsf(ix) < INFINITY) && (fabsf(iy) < INFINITY) && (fabsf(iz) < INFINITY);
		bad += !same(m, m_ref) + !same(a.x, qx) + !same(a.y, qy) + !same(a.z, qz) + !same(inv.x, ix) + !same(inv.y, iy) + !same(inv.z, iz) + (finite != finite_ref);
		// also 1 / x for the reference's __fdiv_rn(1, x) spelling
		bad += !same(ix, __fdiv_rn(1.f, qx));
		fast += (m_ref >= 9.0949470177292824e-13f && m_ref <= 1099511627776.f && fabsf(a0.x) >= 8.6736173798840355e-19f && fabsf(a0.y) >= 8.6736173798840355e-19f && fabsf(a0.z) >= 8.6736173798840355e-19f);
	}
	atomicAdd(out, bad);
	atomicAdd(out + 1, fast);
}

int main()
{
	unsigned long long* d; cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
	const int iters = 1 << 13;
	test<<<148 * 8, 256>>>(d, iters);
	unsigned long long h[2] = { 1, 1 }; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
	printf("vectors tested: %llu (%llu through the shared-reciprocal path), mismatches against __fsqrt_rn / __fdiv_rn / __frcp_rn: %llu (%s)\n",
	       148ull * 8 * 256 * iters, h[1], h[0], cudaGetErrorString(cudaGetLastError()));
	return h[0] != 0;
}
