// The small kernels around the pixel kernel: unstripe (multi-process gather tail), frame fill, the FP32 peak probe,
// the frame completion signal and the device-side TriangleMesh::UpdateTransforms (transform only).  Included by
// rt_api.cu alone (plain __global__ definitions: one translation unit).
#pragma once

#include "rt_kernel.cuh"

namespace rt
{
	// Root-rank tail of the band gather: band r holds strips r, r + world, ... packed; write the
	// frame in row order.  Pure copy (4 B read + 4 B write per pixel), 128-bit when aligned.
	__global__ void __launch_bounds__(256)
	unstripe_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int width, int height, int world,
	                int strips_per_rank, int vec)
	{
		const long long row_units = vec ? width / 4 : width;
		const long long units = row_units * height;
		const long long band_pixels = (long long)strips_per_rank * kBlockH * width;
		for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < units; u += (long long)gridDim.x * blockDim.x)
		{
			const int y = (int)(u / row_units);
			const int xu = (int)(u - (long long)y * row_units);
			const int strip = y / kBlockH, in_strip = y - strip * kBlockH;
			const int rank = strip % world, local = strip / world;
			const long long src_row = (long long)rank * band_pixels + ((long long)local * kBlockH + in_strip) * width;
			if (vec)
				reinterpret_cast<uint4*>(dst + (long long)y * width)[xu] = __ldg(reinterpret_cast<const uint4*>(src + src_row) + xu);
			else
				dst[(long long)y * width + xu] = __ldg(src + src_row + xu);
		}
	}

	// rt_clear_frame: poison pattern into the frame buffer (tests / bench: a frame check must not pass on a stale frame)
	__global__ void __launch_bounds__(256)
	fill32_kernel(uint32_t* dst, uint32_t value, size_t count)
	{
		for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) dst[i] = value;
	}

	// Roofline probe: 8 independent dependent-chains per thread, all in registers.
	template <bool FMA>
	__global__ void __launch_bounds__(256)
	fp32_peak_kernel(float* out, float a, float b, int iterations)
	{
		float acc[8];
		for (int i = 0; i < 8; ++i) acc[i] = (float)(threadIdx.x + i) * 1e-3f;
		for (int it = 0; it < iterations; ++it)
		{
#pragma unroll
			for (int i = 0; i < 8; ++i)
			{
				if (FMA) acc[i] = __fmaf_rn(acc[i], a, b);
				else acc[i] = __fadd_rn(__fmul_rn(acc[i], a), b);
			}
		}
		float s = 0.f;
		for (int i = 0; i < 8; ++i) s += acc[i];
		if (s == 123.456f) out[0] = s;   // never true in practice; keeps the chain alive
	}

	// One rank's "my strips are in the root's frame" signal: a system-scope fenced increment of the word
	// behind the frame's pixels (peer memory).  Stream order puts it after the pixel kernel.
	__global__ void frame_signal_kernel(unsigned int* word)
	{
		__threadfence_system();
		atomicAdd_system(word, 1u);
	}

	// TriangleMesh::UpdateTransforms on the device (DataTypes.h:216-230), one CTA per mesh: every triangle's
	// three vertices go through Matrix::TransformPoint (Matrix.cpp:49-56), its face normal through
	// TransformVector().Normalized() (Matrix.cpp:35-42, Vector3.cpp:42-46), in the reference's operation
	// order; the triangle record {v0|nx, e1|ny, e2|nz} and the box over the indexed vertices (started from
	// +FLT_MAX / +FLT_MIN like a BVH root, DataTypes.h:310-321) are written where the pixel kernel reads them.
	struct TransformParams
	{
		float m[16];                 // Matrix::data[0..3], row by row (x, y, z, w)
		const float* positions;      // 3 per vertex
		const int32_t* indices;      // 3 per triangle
		const float* normals;        // 3 per triangle
		int32_t triangle_count;
		float4* triangles;           // 3 float4 per triangle (this mesh's slice of the stream)
		float4* table;               // this mesh's 3 rows of the mesh table
		int32_t first_triangle;
	};

	__device__ __forceinline__ V3 transform_point(const float* m, float x, float y, float z)
	{
		return v3(add(add(add(mul(m[0], x), mul(m[4], y)), mul(m[8], z)), m[12]),
		          add(add(add(mul(m[1], x), mul(m[5], y)), mul(m[9], z)), m[13]),
		          add(add(add(mul(m[2], x), mul(m[6], y)), mul(m[10], z)), m[14]));
	}

	__global__ void __launch_bounds__(256)
	transform_mesh_kernel(const __grid_constant__ TransformParams p)
	{
		__shared__ float s_min[3][256], s_max[3][256];
		float bmin[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, bmax[3] = { FLT_MIN, FLT_MIN, FLT_MIN };
		for (int t = threadIdx.x; t < p.triangle_count; t += blockDim.x)
		{
			V3 v[3];
			for (int k = 0; k < 3; ++k)
			{
				const float* src = p.positions + 3 * (size_t)p.indices[3 * t + k];
				v[k] = transform_point(p.m, src[0], src[1], src[2]);
				bmin[0] = std_min(bmin[0], v[k].x); bmax[0] = std_max(bmax[0], v[k].x);
				bmin[1] = std_min(bmin[1], v[k].y); bmax[1] = std_max(bmax[1], v[k].y);
				bmin[2] = std_min(bmin[2], v[k].z); bmax[2] = std_max(bmax[2], v[k].z);
			}
			const float nx = p.normals[3 * t], ny = p.normals[3 * t + 1], nz = p.normals[3 * t + 2];
			V3 n = v3(add(add(mul(p.m[0], nx), mul(p.m[4], ny)), mul(p.m[8], nz)),
			          add(add(mul(p.m[1], nx), mul(p.m[5], ny)), mul(p.m[9], nz)),
			          add(add(mul(p.m[2], nx), mul(p.m[6], ny)), mul(p.m[10], nz)));
			normalize(n);
			const V3 e1 = v[1] - v[0], e2 = v[2] - v[0];     // Utils.h:143-144
			p.triangles[3 * t + 0] = make_float4(v[0].x, v[0].y, v[0].z, n.x);
			p.triangles[3 * t + 1] = make_float4(e1.x, e1.y, e1.z, n.y);
			p.triangles[3 * t + 2] = make_float4(e2.x, e2.y, e2.z, n.z);
		}
		for (int k = 0; k < 3; ++k) { s_min[k][threadIdx.x] = bmin[k]; s_max[k][threadIdx.x] = bmax[k]; }
		__syncthreads();
		for (int stride = blockDim.x / 2; stride > 0; stride >>= 1)
		{
			if (threadIdx.x < stride)
				for (int k = 0; k < 3; ++k)
				{
					s_min[k][threadIdx.x] = std_min(s_min[k][threadIdx.x], s_min[k][threadIdx.x + stride]);
					s_max[k][threadIdx.x] = std_max(s_max[k][threadIdx.x], s_max[k][threadIdx.x + stride]);
				}
			__syncthreads();
		}
		if (threadIdx.x == 0)
		{
			const float4 keep = p.table[1];
			p.table[0] = make_float4(s_min[0][0], s_max[0][0], s_min[1][0], s_max[1][0]);
			p.table[1] = make_float4(s_min[2][0], s_max[2][0], keep.z, keep.w);
		}
	}
}
