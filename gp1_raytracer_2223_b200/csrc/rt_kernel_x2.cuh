// Packed variant of the pixel kernel (RT_KERNEL_PACKED): TWO horizontally adjacent pixels per thread, their
// rays carried as packed float2 and advanced together with Blackwell's packed FP32 instruction (FFMA2,
// sm_100+).  Bit-identical to the scalar kernel and covered by the same parity tests.  Status (round 1):
// it halves the iterations of the box / plane / triangle loops (-35 % instructions there), but ray set-up,
// light set-up and shading are still evaluated per half with scalar code and the doubled state costs
// occupancy (121 registers), so on the bunny frame it is ~20 % slower than the scalar kernel, which
// therefore stays the default (profiles/r01_x2_*).
//
// Why: the scalar kernel (rt_kernel.cuh) is bound by instruction issue, not by the FP32 pipe
// (profiles/r01_v1_*, r01_v4_*: issue slots 72-78 % busy, FMA pipe 34 %).  FFMA2 does the work of two
// FMUL / FADD in one issue slot, and everything that is not arithmetic on ray data - primitive
// fetches, address math, loop control, the BVH link decode - is paid once per PAIR of rays.
//
// Exactness.  Every reference operation must stay an individually rounded IEEE-754 binary32 multiply,
// add or subtract.  FFMA2 can express exactly that:
//     a * b = fma(a, b, -0)        a + b = fma(a, 1, b)        a - b = fma(b, -1, a)
// (bit-identical for all operands incl. zeros, denormals, infinities, NaN:
// tools/micro/f32x2_exact.cu, 1.2e9 operand pairs, 0 mismatches).  The three constants arrive as
// kernel parameters, NOT literals: with literals the compiler folds fma(a, b, -0) back into mul.rn.f32x2,
// and ptxas 12.9 then contracts mul.rn.f32x2 + add.rn.f32x2 into ONE fused FFMA2 even under
// --fmad=false (same tool: 1e7 mismatches for __fadd2_rn(__fmul2_rn(a, b), c)).  With run-time
// constants there is nothing left to fuse.  Divisions, square roots, comparisons and min / max stay
// scalar per half, written exactly as in the scalar kernel.
//
// Control flow.  The two halves of a thread may disagree (one ray hits a box, the other does not).
// Every test therefore carries a per-half activity flag, loops run while either half needs them, and
// the BVH walk keeps, per half, the node at which a half that missed a box resumes (the escape link of
// that box), so each half tests exactly the boxes and triangles the reference tests for its ray, in
// the reference's order.
//
// Pixel mapping: a warp owns a 16x4 pixel tile (8 lanes x 2 pixels wide, 4 rows), a 128-thread CTA a
// 32x8 segment of an 8-row strip (same strips as the scalar kernel, so the multi-GPU split is unchanged).
#pragma once

#include "rt_kernel.cuh"

namespace rt
{
namespace x2
{
	constexpr int kThreads = 128;
	constexpr int kBlockW = 32;       // pixels per CTA segment (2 warps x 16)
	constexpr int kWarpW = 16;        // pixels per warp tile row

	struct V3x2
	{
		float2 x, y, z;
	};

	// Packed vector helpers on top of rt::Pk (run-time constants, see the header comment).
	struct Pk : rt::Pk
	{
		using rt::Pk::mul; using rt::Pk::add; using rt::Pk::sub;
		__device__ __forceinline__ V3x2 sub(const V3x2& a, const V3x2& b) const { V3x2 r; r.x = sub(a.x, b.x); r.y = sub(a.y, b.y); r.z = sub(a.z, b.z); return r; }
		// Vector3::Dot, Vector3.cpp:48-51
		__device__ __forceinline__ float2 dot(const V3x2& a, const V3x2& b) const { return add(add(mul(a.x, b.x), mul(a.y, b.y)), mul(a.z, b.z)); }
		// Vector3::Cross, Vector3.cpp:53-57: (s0, -s1, s2); -(p - q) is written q - p (same value, see rt_device.cuh on zero signs)
		__device__ __forceinline__ V3x2 cross(const V3x2& a, const V3x2& b) const
		{
			V3x2 r;
			r.x = sub(mul(a.y, b.z), mul(a.z, b.y));
			r.y = sub(mul(a.z, b.x), mul(a.x, b.z));
			r.z = sub(mul(a.x, b.y), mul(a.y, b.x));
			return r;
		}
	};

	using rt::splat;
	__device__ __forceinline__ V3x2 splat3(float x, float y, float z) { V3x2 r; r.x = splat(x); r.y = splat(y); r.z = splat(z); return r; }
	__device__ __forceinline__ V3x2 pack(V3 a, V3 b) { V3x2 r; r.x = make_float2(a.x, b.x); r.y = make_float2(a.y, b.y); r.z = make_float2(a.z, b.z); return r; }
	__device__ __forceinline__ V3 lo(const V3x2& a) { return v3(a.x.x, a.y.x, a.z.x); }
	__device__ __forceinline__ V3 hi(const V3x2& a) { return v3(a.x.y, a.y.y, a.z.y); }

	// Two rays.  `on0 / on1`: the half takes part at all (valid pixel, hit pixel, light not yet occluded).
	struct Rays
	{
		V3x2 o, d, inv;
		float2 tmax;
		bool fast;      // both halves have a finite 1/dir: FMNMX form of the slab test is exact (Ray::nan_safe)
	};

	__device__ __forceinline__ void set_inverse(Rays& r)
	{
		// Ray constructor, DataTypes.h:550-563
		r.inv.x = make_float2(quo(1.f, r.d.x.x), quo(1.f, r.d.x.y));
		r.inv.y = make_float2(quo(1.f, r.d.y.x), quo(1.f, r.d.y.y));
		r.inv.z = make_float2(quo(1.f, r.d.z.x), quo(1.f, r.d.z.y));
		r.fast = (fabsf(r.inv.x.x) < INFINITY) && (fabsf(r.inv.x.y) < INFINITY) && (fabsf(r.inv.y.x) < INFINITY) &&
		         (fabsf(r.inv.y.y) < INFINITY) && (fabsf(r.inv.z.x) < INFINITY) && (fabsf(r.inv.z.y) < INFINITY);
	}

	constexpr float kTMin = 0.0001f;   // Ray::min, DataTypes.h:545 and Renderer.cpp:136

	// ---- spheres and planes ----------------------------------------------------------------------------

	// HitTest_Sphere, Utils.h:52-63.  Returns per-half hit flags and t.
	__device__ __forceinline__ void sphere2(const Pk& K, const float4 sp, const Rays& r, bool on0, bool on1, bool& h0, bool& h1, float2& t)
	{
		const V3x2 ov = K.sub(splat3(sp.x, sp.y, sp.z), r.o);
		const float2 ov2 = K.dot(ov, ov);
		const float2 p = K.dot(r.d, ov);
		const float2 perp = K.sub(ov2, K.mul(p, p));
		const float r2 = mul(sp.w, sp.w);
		const float2 rem = K.sub(splat(r2), perp);
		h0 = on0 && !(r2 < perp.x);
		h1 = on1 && !(r2 < perp.y);
		t = make_float2(0.f, 0.f);
		if (h0) { t.x = sub(p.x, root(rem.x)); h0 = !(t.x < kTMin || t.x > r.tmax.x); }
		if (h1) { t.y = sub(p.y, root(rem.y)); h1 = !(t.y < kTMin || t.y > r.tmax.y); }
	}

	// HitTest_Plane, Utils.h:84-97 (division only where plane_may_hit cannot rule the hit out).
	__device__ __forceinline__ void plane2(const Pk& K, const float4 po, const float4 pn, const Rays& r, bool on0, bool on1, bool& h0, bool& h1, float2& t)
	{
		const V3x2 n = splat3(pn.x, pn.y, pn.z);
		const float2 num = K.dot(K.sub(splat3(po.x, po.y, po.z), r.o), n);
		const float2 den = K.dot(r.d, n);
		h0 = on0 && plane_may_hit(num.x, den.x, rt::mul(r.tmax.x, 1.000001f));
		h1 = on1 && plane_may_hit(num.y, den.y, rt::mul(r.tmax.y, 1.000001f));
		t = make_float2(0.f, 0.f);
		if (h0) { t.x = quo(num.x, den.x); h0 = (t.x >= kTMin && t.x < r.tmax.x); }
		if (h1) { t.y = quo(num.y, den.y); h1 = (t.y >= kTMin && t.y < r.tmax.y); }
	}

	// ---- boxes -------------------------------------------------------------------------------------------

	// SlabTest_TriangleMesh / SlabTest_BVH, Utils.h:194-243 for two rays.  The six (b - o) * inv products of
	// both rays are twelve packed operations; min / max stay scalar (FAST: FMNMX, else the literal ternaries).
	template <bool FAST>
	__device__ __forceinline__ void slab2(const Pk& K, float bminx, float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz,
	                                       const Rays& r, bool& h0, bool& h1)
	{
		const float2 tx1 = K.mul(K.sub(splat(bminx), r.o.x), r.inv.x), tx2 = K.mul(K.sub(splat(bmaxx), r.o.x), r.inv.x);
		const float2 ty1 = K.mul(K.sub(splat(bminy), r.o.y), r.inv.y), ty2 = K.mul(K.sub(splat(bmaxy), r.o.y), r.inv.y);
		const float2 tz1 = K.mul(K.sub(splat(bminz), r.o.z), r.inv.z), tz2 = K.mul(K.sub(splat(bmaxz), r.o.z), r.inv.z);
		if (FAST)
		{
			const float lo0 = fmaxf(fmaxf(fminf(tx1.x, tx2.x), fminf(ty1.x, ty2.x)), fminf(tz1.x, tz2.x));
			const float hi0 = fminf(fminf(fmaxf(tx1.x, tx2.x), fmaxf(ty1.x, ty2.x)), fmaxf(tz1.x, tz2.x));
			const float lo1 = fmaxf(fmaxf(fminf(tx1.y, tx2.y), fminf(ty1.y, ty2.y)), fminf(tz1.y, tz2.y));
			const float hi1 = fminf(fminf(fmaxf(tx1.y, tx2.y), fmaxf(ty1.y, ty2.y)), fmaxf(tz1.y, tz2.y));
			h0 = hi0 > 0 && hi0 >= lo0;
			h1 = hi1 > 0 && hi1 >= lo1;
		}
		else
		{
			float lo0 = std_min(tx1.x, tx2.x), hi0 = std_max(tx1.x, tx2.x);
			lo0 = std_max(lo0, std_min(ty1.x, ty2.x)); hi0 = std_min(hi0, std_max(ty1.x, ty2.x));
			lo0 = std_max(lo0, std_min(tz1.x, tz2.x)); hi0 = std_min(hi0, std_max(tz1.x, tz2.x));
			float lo1 = std_min(tx1.y, tx2.y), hi1 = std_max(tx1.y, tx2.y);
			lo1 = std_max(lo1, std_min(ty1.y, ty2.y)); hi1 = std_min(hi1, std_max(ty1.y, ty2.y));
			lo1 = std_max(lo1, std_min(tz1.y, tz2.y)); hi1 = std_min(hi1, std_max(tz1.y, tz2.y));
			h0 = hi0 > 0 && hi0 >= lo0;
			h1 = hi1 > 0 && hi1 >= lo1;
		}
	}

	// ---- triangles ----------------------------------------------------------------------------------------

	// HitTest_Triangle, Utils.h:109-160, for two rays against one triangle record.
	// on0 / on1 in: halves that test this triangle; h0 / h1 out: halves that hit it; t: their distances.
	template <int CULL>
	__device__ __forceinline__ void triangle2(const Pk& K, const Tri& T, const Rays& r, bool on0, bool on1, bool& h0, bool& h1, float2& t)
	{
		h0 = false; h1 = false;
		const float2 c = K.dot(splat3(T.a0.w, T.a1.w, T.a2.w), r.d);
		bool p0 = on0 && cull_pass<CULL>(c.x), p1 = on1 && cull_pass<CULL>(c.y);
		if (!(p0 | p1)) return;

		const V3x2 e1 = splat3(T.a1.x, T.a1.y, T.a1.z), e2 = splat3(T.a2.x, T.a2.y, T.a2.z);
		const V3x2 h = K.cross(r.d, e2);
		const float2 a = K.dot(e1, h);
		p0 = p0 && !(fabsf(a.x) < FLT_EPSILON);
		p1 = p1 && !(fabsf(a.y) < FLT_EPSILON);
		if (!(p0 | p1)) return;

		float2 f = make_float2(0.f, 0.f);
		if (p0) f.x = quo(1.f, a.x);
		if (p1) f.y = quo(1.f, a.y);
		const V3x2 s = K.sub(r.o, splat3(T.a0.x, T.a0.y, T.a0.z));
		const float2 u = K.mul(f, K.dot(s, h));
		p0 = p0 && !(u.x < 0.f || u.x > 1.f);
		p1 = p1 && !(u.y < 0.f || u.y > 1.f);
		if (!(p0 | p1)) return;

		const V3x2 q = K.cross(s, e1);
		const float2 v = K.mul(f, K.dot(r.d, q));
		const float2 uv = K.add(u, v);
		p0 = p0 && !(v.x < 0.f || uv.x > 1.f);
		p1 = p1 && !(v.y < 0.f || uv.y > 1.f);
		if (!(p0 | p1)) return;

		t = K.mul(f, K.dot(e2, q));
		h0 = p0 && !(t.x < kTMin || t.x >= r.tmax.x);
		h1 = p1 && !(t.y < kTMin || t.y >= r.tmax.y);
	}

	// ---- meshes: slab + linear (Utils.h:298-325) ------------------------------------------------------------

	// Closest hit: strict '<' on t keeps the first triangle on ties.
	template <int CULL>
	__device__ __forceinline__ void mesh_closest2(const Pk& K, const float4* tri, int count, const Rays& r, bool on0, bool on1,
	                                               float2& best_t, int& best0, int& best1)
	{
		for (int i = 0; i < count; ++i)
		{
			const Tri T = load_tri(tri + 3 * i);
			bool h0, h1; float2 t;
			triangle2<CULL>(K, T, r, on0, on1, h0, h1, t);
			if (h0 && t.x < best_t.x) { best_t.x = t.x; best0 = i; }
			if (h1 && t.y < best_t.y) { best_t.y = t.y; best1 = i; }
		}
	}

	// Any hit: a half stops at its first hit; returns through on0 / on1 the halves still unoccluded.
	template <int CULL>
	__device__ __forceinline__ void mesh_any2(const Pk& K, const float4* tri, int count, const Rays& r, bool& on0, bool& on1, bool& occ0, bool& occ1)
	{
		for (int i = 0; i < count && (on0 | on1); ++i)
		{
			const Tri T = load_tri(tri + 3 * i);
			bool h0, h1; float2 t;
			triangle2<CULL>(K, T, r, on0, on1, h0, h1, t);
			if (h0) { occ0 = true; on0 = false; }
			if (h1) { occ1 = true; on1 = false; }
		}
	}

	// ---- meshes: the reference's BVH walk (Utils.h:246-288) for two rays -----------------------------------
	//
	// One node pointer per thread, driven by whichever half still has business in the current subtree.
	// A half that misses a box (or is not taking part) sleeps until the walk reaches the node stored in its
	// resume slot - the escape link of the box it missed, which is where the reference's recursion continues
	// for that ray.  kAwake / kDone never equal a node index.
	constexpr int kAwake = -2, kDone = -3;

	template <int CULL, bool FAST, bool ANY>
	__device__ __forceinline__ void bvh2(const Pk& K, const float4* nodes, const float4* tri, const Rays& r, bool on0, bool on1,
	                                      float2& best_t, int& best0, int& best1, bool& occ0, bool& occ1)
	{
		int res0 = on0 ? kAwake : kDone, res1 = on1 ? kAwake : kDone;
		int node = 0;
		while (node >= 0)
		{
			if (res0 == node) res0 = kAwake;
			if (res1 == node) res1 = kAwake;
			// `node` is the record's byte offset (rt::BvhLink): offsets name nodes as well as indices do
			const float4* rec = node_at(nodes, node);
			const float4 n0 = __ldg(rec), n1 = __ldg(rec + 1);
			const int hit_link = __float_as_int(n1.z);
			const int escape = __float_as_int(n1.w);
			bool h0, h1;
			slab2<FAST>(K, n0.x, n0.z, n1.x, n0.y, n0.w, n1.y, r, h0, h1);
			h0 = h0 && (res0 == kAwake);
			h1 = h1 && (res1 == kAwake);
			if (!h0 && res0 == kAwake) res0 = escape;     // this ray skips the subtree, like the early return of Utils.h:251-254
			if (!h1 && res1 == kAwake) res1 = escape;
			if (!(h0 | h1)) { node = escape; continue; }
			if (!BvhLink::is_leaf(hit_link)) { node = hit_link; continue; }
			const int count = BvhLink::leaf_count(hit_link), first = BvhLink::leaf_first(hit_link);
			for (int k = 0; k < count && (h0 | h1); ++k)
			{
				const Tri T = load_tri(tri + 3 * (first + k));
				bool g0, g1; float2 t;
				triangle2<CULL>(K, T, r, h0, h1, g0, g1, t);
				if (ANY)
				{
					if (g0) { occ0 = true; res0 = kDone; h0 = false; }
					if (g1) { occ1 = true; res1 = kDone; h1 = false; }
				}
				else
				{
					if (g0 && t.x < best_t.x) { best_t.x = t.x; best0 = first + k; }
					if (g1 && t.y < best_t.y) { best_t.y = t.y; best1 = first + k; }
				}
			}
			if (ANY && res0 == kDone && res1 == kDone) return;
			node = escape;
		}
	}

	template <bool FAST, bool ANY>
	__device__ __forceinline__ void bvh2_cull(int cull, const Pk& K, const float4* nodes, const float4* tri, const Rays& r, bool on0, bool on1,
	                                           float2& best_t, int& best0, int& best1, bool& occ0, bool& occ1)
	{
		if (cull == RT_CULL_BACK_FACE) bvh2<RT_CULL_BACK_FACE, FAST, ANY>(K, nodes, tri, r, on0, on1, best_t, best0, best1, occ0, occ1);
		else if (cull == RT_CULL_FRONT_FACE) bvh2<RT_CULL_FRONT_FACE, FAST, ANY>(K, nodes, tri, r, on0, on1, best_t, best0, best1, occ0, occ1);
		else bvh2<RT_CULL_NONE, FAST, ANY>(K, nodes, tri, r, on0, on1, best_t, best0, best1, occ0, occ1);
	}

	// ---- Scene::GetClosestHit (Scene.cpp:29-66) for two primary rays ------------------------------------------

	// What hit a half: nothing, sphere i, plane i, or triangle j of mesh m (in upload order).
	constexpr int kNone = -1, kSphere = 0x10000000, kPlane = 0x20000000, kTriangle = 0x30000000, kKindMask = 0x30000000;
	constexpr int kMeshShift = 23, kIndexMask = (1 << kMeshShift) - 1;

	template <bool BVH>
	__device__ __forceinline__ void closest_hit2(const Pk& K, const SharedScene& sc, const SceneDevice& dev, const Rays& r, bool on0, bool on1,
	                                              float2& best_t, int& id0, int& id1)
	{
		best_t = splat(FLT_MAX);
		id0 = kNone; id1 = kNone;
		for (int i = 0; i < dev.n_spheres; ++i)
		{
			bool h0, h1; float2 t;
			sphere2(K, sc.sphere[i], r, on0, on1, h0, h1, t);
			if (h0 && t.x < best_t.x) { best_t.x = t.x; id0 = kSphere | i; }
			if (h1 && t.y < best_t.y) { best_t.y = t.y; id1 = kSphere | i; }
		}
		for (int i = 0; i < dev.n_planes; ++i)
		{
			bool h0, h1; float2 t;
			plane2(K, sc.plane_o[i], sc.plane_n[i], r, on0, on1, h0, h1, t);
			if (h0 && t.x < best_t.x) { best_t.x = t.x; id0 = kPlane | i; }
			if (h1 && t.y < best_t.y) { best_t.y = t.y; id1 = kPlane | i; }
		}
		for (int m = 0; m < dev.n_meshes; ++m)
		{
			const float4 b0 = sc.mesh[3 * m], b1 = sc.mesh[3 * m + 1], info = sc.mesh[3 * m + 2];
			const int first = __float_as_int(b1.z), count = __float_as_int(b1.w);
			const int cull = __float_as_int(info.x);
			const float4* tri = dev.triangles + 3 * (size_t)first;
			int t0 = -1, t1 = -1;
			bool dummy0 = false, dummy1 = false;
			if (BVH)
			{
				if (count == 0) continue;
				const float4* nodes = dev.bvh_nodes + 2 * (size_t)__float_as_int(info.z);
				if (r.fast) bvh2_cull<true, false>(cull, K, nodes, tri, r, on0, on1, best_t, t0, t1, dummy0, dummy1);
				else bvh2_cull<false, false>(cull, K, nodes, tri, r, on0, on1, best_t, t0, t1, dummy0, dummy1);
			}
			else
			{
				bool s0, s1;
				if (r.fast) slab2<true>(K, b0.x, b0.z, b1.x, b0.y, b0.w, b1.y, r, s0, s1);
				else slab2<false>(K, b0.x, b0.z, b1.x, b0.y, b0.w, b1.y, r, s0, s1);
				s0 = s0 && on0; s1 = s1 && on1;
				if (!(s0 | s1)) continue;
				if (cull == RT_CULL_BACK_FACE) mesh_closest2<RT_CULL_BACK_FACE>(K, tri, count, r, s0, s1, best_t, t0, t1);
				else if (cull == RT_CULL_FRONT_FACE) mesh_closest2<RT_CULL_FRONT_FACE>(K, tri, count, r, s0, s1, best_t, t0, t1);
				else mesh_closest2<RT_CULL_NONE>(K, tri, count, r, s0, s1, best_t, t0, t1);
			}
			if (t0 >= 0) id0 = kTriangle | (m << kMeshShift) | t0;
			if (t1 >= 0) id1 = kTriangle | (m << kMeshShift) | t1;
		}
	}

	// The HitRecord of one half (Utils.h:65-69, 89-93, 176-180; Scene.cpp:40).
	__device__ __forceinline__ Hit make_hit(const SharedScene& sc, const SceneDevice& dev, int id, float t, V3 o, V3 d)
	{
		Hit hit;
		hit.did = id != kNone;
		hit.t = t; hit.material = 0;
		hit.origin = v3(0.f, 0.f, 0.f); hit.normal = v3(0.f, 0.f, 0.f);
		if (!hit.did) return hit;
		hit.origin = o + d * t;
		const int kind = id & kKindMask, index = id & kIndexMask;
		if (kind == kSphere)
		{
			hit.material = sc.sphere_mat[index];
			hit.normal = hit.origin - v3(sc.sphere[index]);
			normalize(hit.normal);
		}
		else if (kind == kPlane)
		{
			hit.material = __float_as_int(sc.plane_o[index].w);
			hit.normal = v3(sc.plane_n[index]);
		}
		else
		{
			const int m = (id & ~kKindMask) >> kMeshShift;
			const float4 b1 = sc.mesh[3 * m + 1], info = sc.mesh[3 * m + 2];
			const Tri T = load_tri(dev.triangles + 3 * ((size_t)__float_as_int(b1.z) + index));
			hit.material = __float_as_int(info.y);
			hit.normal = v3(T.a0.w, T.a1.w, T.a2.w);
		}
		return hit;
	}

	// ---- Scene::DoesHit (Scene.cpp:68-96) for two shadow rays towards the same light -----------------------------

	template <bool BVH>
	__device__ __forceinline__ void does_hit2(const Pk& K, const SharedScene& sc, const SceneDevice& dev, const Rays& r, bool on0, bool on1,
	                                           bool& occ0, bool& occ1)
	{
		occ0 = false; occ1 = false;
		for (int i = 0; i < dev.n_spheres && (on0 | on1); ++i)
		{
			bool h0, h1; float2 t;
			sphere2(K, sc.sphere[i], r, on0, on1, h0, h1, t);
			if (h0) { occ0 = true; on0 = false; }
			if (h1) { occ1 = true; on1 = false; }
		}
		for (int i = 0; i < dev.n_planes && (on0 | on1); ++i)
		{
			bool h0, h1; float2 t;
			plane2(K, sc.plane_o[i], sc.plane_n[i], r, on0, on1, h0, h1, t);
			if (h0) { occ0 = true; on0 = false; }
			if (h1) { occ1 = true; on1 = false; }
		}
		for (int m = 0; m < dev.n_meshes && (on0 | on1); ++m)
		{
			const float4 b0 = sc.mesh[3 * m], b1 = sc.mesh[3 * m + 1], info = sc.mesh[3 * m + 2];
			const int first = __float_as_int(b1.z), count = __float_as_int(b1.w);
			int cull = __float_as_int(info.x);
			// Utils.h:114-127: shadow rays see the opposite cull mode
			cull = (cull == RT_CULL_FRONT_FACE) ? RT_CULL_BACK_FACE : (cull == RT_CULL_BACK_FACE ? RT_CULL_FRONT_FACE : cull);
			const float4* tri = dev.triangles + 3 * (size_t)first;
			if (BVH)
			{
				if (count == 0) continue;
				const float4* nodes = dev.bvh_nodes + 2 * (size_t)__float_as_int(info.z);
				float2 unused_t = splat(0.f); int u0 = -1, u1 = -1;
				bool o0 = false, o1 = false;
				if (r.fast) bvh2_cull<true, true>(cull, K, nodes, tri, r, on0, on1, unused_t, u0, u1, o0, o1);
				else bvh2_cull<false, true>(cull, K, nodes, tri, r, on0, on1, unused_t, u0, u1, o0, o1);
				if (o0) { occ0 = true; on0 = false; }
				if (o1) { occ1 = true; on1 = false; }
			}
			else
			{
				bool s0, s1;
				if (r.fast) slab2<true>(K, b0.x, b0.z, b1.x, b0.y, b0.w, b1.y, r, s0, s1);
				else slab2<false>(K, b0.x, b0.z, b1.x, b0.y, b0.w, b1.y, r, s0, s1);
				s0 = s0 && on0; s1 = s1 && on1;
				if (!(s0 | s1)) continue;
				bool o0 = false, o1 = false;
				if (cull == RT_CULL_BACK_FACE) mesh_any2<RT_CULL_BACK_FACE>(K, tri, count, r, s0, s1, o0, o1);
				else if (cull == RT_CULL_FRONT_FACE) mesh_any2<RT_CULL_FRONT_FACE>(K, tri, count, r, s0, s1, o0, o1);
				else mesh_any2<RT_CULL_NONE>(K, tri, count, r, s0, s1, o0, o1);
				if (o0) { occ0 = true; on0 = false; }
				if (o1) { occ1 = true; on1 = false; }
			}
		}
	}

	// ---- Renderer::RenderPixel (Renderer.cpp:100-182) for the pixels (px, py) and (px + 1, py) -------------------

	__device__ __forceinline__ V3 primary_direction(const FrameParams& p, int px, int py)
	{
		// Renderer.cpp:107-114
		const float cx = mul(mul(sub(mul(2.f, quo(add((float)px, 0.5f), (float)p.width)), 1.f), p.aspect), p.fov);
		const float cy = mul(sub(1.f, quo(mul(2.f, add((float)py, 0.5f)), (float)p.height)), p.fov);
		V3 d = v3(add(add(mul(p.right_x, cx), mul(p.up_x, cy)), p.fwd_x),
		          add(add(mul(p.right_y, cx), mul(p.up_y, cy)), p.fwd_y),
		          add(add(mul(p.right_z, cx), mul(p.up_z, cy)), p.fwd_z));
		normalize(d);
		return d;
	}

	__device__ __forceinline__ uint32_t pack_pixel(const FrameParams& p, V3 color, float shadow_factor, bool did)
	{
		if (did) color = color * shadow_factor;                                  // Renderer.cpp:173
		const float max_value = std_max(color.x, std_max(color.y, color.z));   // ColorRGB::MaxToOne, ColorRGB.h:12-17
		if (max_value > 1.f) { color.x = quo(color.x, max_value); color.y = quo(color.y, max_value); color.z = quo(color.z, max_value); }
		const uint32_t R = (uint32_t)__float2int_rz(mul(color.x, 255.f)) & 0xffu;
		const uint32_t G = (uint32_t)__float2int_rz(mul(color.y, 255.f)) & 0xffu;
		const uint32_t B = (uint32_t)__float2int_rz(mul(color.z, 255.f)) & 0xffu;
		return (R << p.r_shift) | (G << p.g_shift) | (B << p.b_shift) | p.alpha_mask;
	}

	template <int MODE>
	__device__ __forceinline__ V3 light_contribution(const SharedScene& sc, const Hit& hit, const float4 la, const float4 lb, V3 l, V3 view_neg)
	{
		Counters<false> cnt;
		if (MODE == RT_LIGHTING_COMBINED)
		{
			const float oa = std_max(dot(hit.normal, l), 0.f);
			const V3 e = radiance(la, lb, hit.origin);
			const V3 brdf = shade(sc.material[2 * hit.material], sc.material[2 * hit.material + 1], hit.normal, l, ViewInRegisters{ view_neg }, cnt);
			return v3(mul(mul(e.x, oa), brdf.x), mul(mul(e.y, oa), brdf.y), mul(mul(e.z, oa), brdf.z));   // Renderer.cpp:152
		}
		if (MODE == RT_LIGHTING_OBSERVED_AREA)
		{
			const float oa = std_max(dot(hit.normal, l), 0.f);
			return v3(oa, oa, oa);
		}
		if (MODE == RT_LIGHTING_RADIANCE) return radiance(la, lb, hit.origin);
		return shade(sc.material[2 * hit.material], sc.material[2 * hit.material + 1], hit.normal, l, ViewInRegisters{ view_neg }, cnt);
	}

	template <int MODE, int SHADOWS, bool BVH>
	__device__ __forceinline__ void render_pair(const Pk& K, const SharedScene& sc, const SceneDevice& dev, const FrameParams& p,
	                                             int px, int py, bool valid0, bool valid1, uint32_t& pixel0, uint32_t& pixel1)
	{
		const V3 d0 = primary_direction(p, px, py), d1 = primary_direction(p, px + 1, py);
		const V3 cam = v3(p.cam_ox, p.cam_oy, p.cam_oz);
		Rays view;
		view.o = pack(cam, cam);
		view.d = pack(d0, d1);
		view.tmax = splat(FLT_MAX);
		set_inverse(view);

		float2 best_t; int id0, id1;
		closest_hit2<BVH>(K, sc, dev, view, valid0, valid1, best_t, id0, id1);
		const Hit hit0 = make_hit(sc, dev, id0, best_t.x, cam, d0), hit1 = make_hit(sc, dev, id1, best_t.y, cam, d1);

		float sf0 = 1.f, sf1 = 1.f;
		V3 c0 = v3(0.f, 0.f, 0.f), c1 = v3(0.f, 0.f, 0.f);
		if (hit0.did | hit1.did)
		{
			const V3 oo0 = hit0.origin + hit0.normal * 0.0001f, oo1 = hit1.origin + hit1.normal * 0.0001f;   // Renderer.cpp:126
			const V3 vn0 = neg(d0), vn1 = neg(d1);
			for (int li = 0; li < dev.n_lights; ++li)
			{
				const float4 la = sc.light_a[li], lb = sc.light_b[li];
				const int ltype = __float_as_int(lb.w);
				const bool known = (ltype == RT_LIGHT_POINT || ltype == RT_LIGHT_DIRECTIONAL);
				// GetDirectionToLight, Utils.h:341-353
				V3 l0 = known ? (v3(la) - oo0) : v3(0.f, 0.f, 0.f), l1 = known ? (v3(la) - oo1) : v3(0.f, 0.f, 0.f);
				const float mag0 = normalize(l0), mag1 = normalize(l1);
				bool occ0 = false, occ1 = false;
				if (SHADOWS)
				{
					Rays sh;
					sh.o = pack(oo0, oo1);
					sh.d = pack(l0, l1);
					sh.tmax = make_float2(mag0, mag1);                      // Renderer.cpp:136
					set_inverse(sh);
					does_hit2<BVH>(K, sc, dev, sh, hit0.did, hit1.did, occ0, occ1);
				}
				if (hit0.did)
				{
					if (occ0) sf0 = mul(sf0, 0.95f);                        // Renderer.cpp:139-140
					else c0 = c0 + light_contribution<MODE>(sc, hit0, la, lb, l0, vn0);
				}
				if (hit1.did)
				{
					if (occ1) sf1 = mul(sf1, 0.95f);
					else c1 = c1 + light_contribution<MODE>(sc, hit1, la, lb, l1, vn1);
				}
			}
		}
		pixel0 = pack_pixel(p, c0, sf0, hit0.did);
		pixel1 = pack_pixel(p, c1, sf1, hit1.did);
	}

	template <int MODE, int SHADOWS, bool BVH>
	__global__ void __launch_bounds__(kThreads)
	render_kernel_x2(const __grid_constant__ SceneDevice dev, const __grid_constant__ FrameParams p)
	{
		__shared__ SharedScene sc;
		stage_scene<kThreads>(sc, dev, v3(p.cam_ox, p.cam_oy, p.cam_oz));
		__syncthreads();

		Pk K;
		K.neg0 = dev.k_neg0; K.one = dev.k_one; K.mone = dev.k_mone;

		const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
		const int tx = lane & 7, ty = lane >> 3;
		const int wx = warp & 1, wy = warp >> 1;
		const int px = blockIdx.x * kBlockW + wx * kWarpW + 2 * tx;
		const int local_y = wy * kTileH + ty;
		const int py = p.row_begin + ((int)blockIdx.y * p.strip_step + p.strip_first) * kBlockH + local_y;
		const bool row_ok = py < p.row_end;
		const bool valid0 = row_ok && (px < p.width), valid1 = row_ok && (px + 1 < p.width);

		uint32_t pixel0 = 0, pixel1 = 0;
		if (valid0) render_pair<MODE, SHADOWS, BVH>(K, sc, dev, p, px, py, valid0, valid1, pixel0, pixel1);

		const int dst_row = p.dst_full_frame ? py : ((int)blockIdx.y * kBlockH + local_y);
		uint32_t* row = p.dst + (size_t)dst_row * (size_t)p.width;
		if (p.vector_store)
		{
			// two lanes x two pixels -> one 128-bit store (width % 4 == 0, so a valid px % 4 == 0 has px + 3 < width)
			const uint32_t q0 = __shfl_down_sync(0xffffffffu, pixel0, 1);
			const uint32_t q1 = __shfl_down_sync(0xffffffffu, pixel1, 1);
			if (valid0 && (tx & 1) == 0) *reinterpret_cast<uint4*>(row + px) = make_uint4(pixel0, pixel1, q0, q1);
		}
		else
		{
			if (valid0) row[px] = pixel0;
			if (valid1) row[px + 1] = pixel1;
		}

		if (p.band_done) signal_band_done(p);
	}
}
}
