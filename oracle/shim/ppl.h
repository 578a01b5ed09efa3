// TEST INFRASTRUCTURE (oracle build only) -- not part of the shipped product.
//
// Stand-in for MSVC's <ppl.h>, which the reference includes at
// /root/reference/source/Renderer.cpp:17 and uses at Renderer.cpp:81
// (concurrency::parallel_for over the pixel range).  PPL's default
// auto_partitioner hands out contiguous sub-ranges with work stealing; the
// closest OpenMP equivalent is a dynamically scheduled loop over chunks.
// Thread count comes from OMP_NUM_THREADS / omp_set_num_threads().
#pragma once
#include <cstdint>

#ifndef GP1_PPL_CHUNK
#define GP1_PPL_CHUNK 128
#endif

namespace concurrency
{
	template <class Index, class Body>
	void parallel_for(Index first, Index last, const Body& body)
	{
		const long long lo = static_cast<long long>(first);
		const long long hi = static_cast<long long>(last);
#pragma omp parallel for schedule(dynamic, GP1_PPL_CHUNK)
		for (long long i = lo; i < hi; ++i)
		{
			body(static_cast<Index>(i));
		}
	}
}
