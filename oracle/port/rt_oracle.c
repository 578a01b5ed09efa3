/*
 * TEST INFRASTRUCTURE -- CPU restatement (plain C) of the reference's per-pixel path.
 * See rt_oracle.h for who may use this and for the parity status (pinned against the
 * compiled reference, tests/test_oracle_port.py).
 *
 * Every function cites the reference lines it restates.  Arithmetic is written in the
 * evaluation order the reference's C++ implies (left-to-right association, no
 * contraction: build with -ffp-contract=off), including its operator overloads:
 *   Vector3::Dot        (a.x*b.x + a.y*b.y) + a.z*b.z          source/Vector3.cpp:48-51
 *   Vector3::Normalize  three true divisions by sqrtf(...)      source/Vector3.cpp:32-46
 *   std::max(a,b)       (a < b) ? b : a   -- NOT fmaxf          (NaN behaviour differs)
 *   std::min(a,b)       (b < a) ? b : a
 */
#include "rt_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float x, y, z; } v3;
typedef struct { float r, g, b; } rgb;

static inline float std_max(float a, float b) { return (a < b) ? b : a; }
static inline float std_min(float a, float b) { return (b < a) ? b : a; }

static inline v3 v3_make(float x, float y, float z) { v3 v = { x, y, z }; return v; }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }   /* Vector3.cpp:118-121 */
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }   /* Vector3.cpp:113-116 */
static inline v3 v3_scale(v3 a, float s) { return v3_make(a.x * s, a.y * s, a.z * s); }     /* Vector3.cpp:103-106, Vector3.h:57-60 */
static inline float v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }        /* Vector3.cpp:48-51 */
static inline float v3_sqr(v3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }              /* Vector3.cpp:27-30 */
static inline float v3_mag(v3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }       /* Vector3.cpp:22-25 */

/* Vector3::Normalize / Normalized, Vector3.cpp:32-46. */
static inline float v3_normalize(v3* a)
{
	const float m = v3_mag(*a);
	a->x /= m; a->y /= m; a->z /= m;
	return m;
}

/* Vector3::Cross, Vector3.cpp:53-57, literally: UnitX*s0 - UnitY*s1 + UnitZ*s2 with the
 * zero products kept (they only matter for signed zeros / non-finite inputs). */
static inline v3 v3_cross(v3 a, v3 b)
{
	const float s0 = a.y * b.z - a.z * b.y;
	const float s1 = a.x * b.z - a.z * b.x;
	const float s2 = a.x * b.y - a.y * b.x;
	const v3 ux = v3_make(1.f * s0, 0.f * s0, 0.f * s0);
	const v3 uy = v3_make(0.f * s1, 1.f * s1, 0.f * s1);
	const v3 uz = v3_make(0.f * s2, 0.f * s2, 1.f * s2);
	return v3_add(v3_sub(ux, uy), uz);
}

/* Ray, DataTypes.h:539-565. */
typedef struct { v3 origin, direction, inv; float tmin, tmax; } ray_t;
static inline ray_t ray_make(v3 o, v3 d, float tmin, float tmax)
{
	ray_t r; r.origin = o; r.direction = d;
	r.inv = v3_make(1.f / d.x, 1.f / d.y, 1.f / d.z);
	r.tmin = tmin; r.tmax = tmax;
	return r;
}

/* HitRecord, DataTypes.h:567-575. */
typedef struct { v3 origin, normal; float t; int did_hit; unsigned char material; } hit_t;
static inline hit_t hit_default(void)
{
	hit_t h; h.origin = v3_make(0, 0, 0); h.normal = v3_make(0, 0, 0); h.t = FLT_MAX; h.did_hit = 0; h.material = 0;
	return h;
}

#define CNT(slot) do { if (cnt) cnt[slot]++; } while (0)

/* HitTest_Sphere, Utils.h:52-71. */
static int hit_sphere(const rt_spheres_soa* s, int i, const ray_t* ray, hit_t* rec, int ignore, uint64_t* cnt)
{
	const v3 c = v3_make(s->origin_x[i], s->origin_y[i], s->origin_z[i]);
	const v3 ov = v3_sub(c, ray->origin);
	const float ov2 = v3_sqr(ov);
	const float p = v3_dot(ray->direction, ov);
	const float perp = ov2 - p * p;
	const float r2 = s->radius[i] * s->radius[i];
	if (r2 < perp) { CNT(ignore ? RT_CNT_SPHERE_S_DISC : RT_CNT_SPHERE_P_DISC); return 0; }
	const float dist = sqrtf(r2 - perp);
	const float t = p - dist;
	if (t < ray->tmin || t > ray->tmax) { CNT(ignore ? RT_CNT_SPHERE_S_TREJ : RT_CNT_SPHERE_P_TREJ); return 0; }
	if (ignore) { CNT(RT_CNT_SPHERE_S_HIT); return 1; }
	CNT(RT_CNT_SPHERE_P_HIT);
	rec->did_hit = 1;
	rec->material = s->material_index[i];
	rec->origin = v3_add(ray->origin, v3_scale(ray->direction, t));
	rec->normal = v3_sub(rec->origin, c);
	rec->t = t;
	return 1;
}

/* HitTest_Plane, Utils.h:82-98. */
static int hit_plane(const rt_planes_soa* pl, int i, const ray_t* ray, hit_t* rec, int ignore, uint64_t* cnt)
{
	const v3 o = v3_make(pl->origin_x[i], pl->origin_y[i], pl->origin_z[i]);
	const v3 n = v3_make(pl->normal_x[i], pl->normal_y[i], pl->normal_z[i]);
	const float t = v3_dot(v3_sub(o, ray->origin), n) / v3_dot(ray->direction, n);
	CNT(ignore ? RT_CNT_PLANE_S_TEST : RT_CNT_PLANE_P_TEST);
	if (t >= ray->tmin && t < ray->tmax)
	{
		if (!ignore)
		{
			CNT(RT_CNT_PLANE_P_HIT);
			rec->did_hit = 1;
			rec->material = pl->material_index[i];
			rec->normal = n;
			rec->origin = v3_add(ray->origin, v3_scale(ray->direction, t));
			rec->t = t;
		}
		return 1;
	}
	return 0;
}

/* HitTest_Triangle, Utils.h:109-184. */
static int hit_triangle(v3 v0, v3 v1, v3 v2, v3 normal, int cull_mode, unsigned char material,
                        const ray_t* ray, hit_t* rec, int ignore, uint64_t* cnt)
{
	const int base = ignore ? RT_CNT_TRI_S_CULLED : RT_CNT_TRI_P_CULLED;
	const float cull_dot = v3_dot(normal, ray->direction);
	if (fabsf(cull_dot) < FLT_EPSILON) { CNT(base + 0); return 0; }

	int mode = cull_mode;                       /* Utils.h:114-127: inverted for shadow rays */
	if (ignore)
	{
		if (mode == RT_CULL_FRONT_FACE) mode = RT_CULL_BACK_FACE;
		else if (mode == RT_CULL_BACK_FACE) mode = RT_CULL_FRONT_FACE;
	}
	if (mode == RT_CULL_FRONT_FACE) { if (cull_dot < 0) { CNT(base + 0); return 0; } }
	else if (mode == RT_CULL_BACK_FACE) { if (cull_dot > 0) { CNT(base + 0); return 0; } }

	const v3 e1 = v3_sub(v1, v0);
	const v3 e2 = v3_sub(v2, v0);
	const v3 h = v3_cross(ray->direction, e2);
	const float a = v3_dot(e1, h);
	if (fabsf(a) < FLT_EPSILON) { CNT(base + 1); return 0; }

	const float f = 1.f / a;
	const v3 s = v3_sub(ray->origin, v0);
	const float u = f * v3_dot(s, h);
	if (u < 0.f || u > 1.f) { CNT(base + 2); return 0; }

	const v3 q = v3_cross(s, e1);
	const float v = f * v3_dot(ray->direction, q);
	if (v < 0.f || (u + v) > 1.f) { CNT(base + 3); return 0; }

	const float t = f * v3_dot(e2, q);
	if (t < ray->tmin || t >= ray->tmax) { CNT(base + 4); return 0; }
	CNT(base + 5);

	const v3 ip = v3_add(ray->origin, v3_scale(ray->direction, t));
	if (!ignore)
	{
		rec->material = material;
		rec->did_hit = 1;
		rec->normal = normal;
		rec->origin = ip;
		rec->t = t;
	}
	return 1;
}

/* SlabTest_TriangleMesh / SlabTest_BVH, Utils.h:194-216 and 221-243 (same arithmetic). */
static int slab_test(const float bmin[3], const float bmax[3], const ray_t* ray)
{
	const float tx1 = (bmin[0] - ray->origin.x) * ray->inv.x;
	const float tx2 = (bmax[0] - ray->origin.x) * ray->inv.x;
	float t_min = std_min(tx1, tx2);
	float t_max = std_max(tx1, tx2);
	const float ty1 = (bmin[1] - ray->origin.y) * ray->inv.y;
	const float ty2 = (bmax[1] - ray->origin.y) * ray->inv.y;
	t_min = std_max(t_min, std_min(ty1, ty2));
	t_max = std_min(t_max, std_max(ty1, ty2));
	const float tz1 = (bmin[2] - ray->origin.z) * ray->inv.z;
	const float tz2 = (bmax[2] - ray->origin.z) * ray->inv.z;
	t_min = std_max(t_min, std_min(tz1, tz2));
	t_max = std_min(t_max, std_max(tz1, tz2));
	return t_max > 0 && t_max >= t_min;
}

static inline v3 mesh_pos(const rt_mesh_desc* m, int vi)
{
	return v3_make(m->positions[3 * vi], m->positions[3 * vi + 1], m->positions[3 * vi + 2]);
}
static inline v3 mesh_nrm(const rt_mesh_desc* m, int ti)
{
	return v3_make(m->normals[3 * ti], m->normals[3 * ti + 1], m->normals[3 * ti + 2]);
}

void rto_mesh_bounds(const rt_mesh_desc* m, float out_min[3], float out_max[3])
{
	/* UpdateNodeBounds on the root, DataTypes.h:310-321, with MaxVector / MinVector from
	 * Vector3.cpp:13-14 (MinVector is +FLT_MIN, a reference quirk that only inflates boxes). */
	for (int k = 0; k < 3; ++k) { out_min[k] = FLT_MAX; out_max[k] = FLT_MIN; }
	for (int i = 0; i < 3 * m->triangle_count; ++i)
	{
		const float* p = &m->positions[3 * m->indices[i]];
		for (int k = 0; k < 3; ++k)
		{
			out_min[k] = std_min(out_min[k], p[k]);
			out_max[k] = std_max(out_max[k], p[k]);
		}
	}
}

/* IntersectionTest_BVH, Utils.h:246-288. */
static void bvh_traverse(const rto_mesh* mesh, unsigned node_idx, const ray_t* ray, int* did_hit,
                         hit_t* hit_record, hit_t* current, int ignore, uint64_t* cnt)
{
	const rto_bvh_node* node = &mesh->desc.bvh_nodes[node_idx];
	/* The reference keeps walking sibling subtrees after an any-hit leaf returned (Utils.h:274 only
	 * leaves the leaf); nothing it finds there can change the boolean, so stop here. */
	if (ignore && *did_hit) return;
	CNT(ignore ? RT_CNT_BVH_S_NODE : RT_CNT_BVH_P_NODE);
	if (!slab_test(node->min_aabb, node->max_aabb, ray)) return;
	if (node->idx_count > 0)
	{
		const rt_mesh_desc* m = &mesh->desc;
		for (int idx = 0; idx < (int)node->idx_count; idx += 3)
		{
			const int leaf = (int)node->first_idx + idx;
			if (hit_triangle(mesh_pos(m, m->indices[leaf]), mesh_pos(m, m->indices[leaf + 1]),
			                 mesh_pos(m, m->indices[leaf + 2]), mesh_nrm(m, leaf / 3),
			                 m->cull_mode, m->material_index, ray, current, ignore, cnt))
			{
				*did_hit = 1;
				if (ignore) return;
				if (current->t < hit_record->t) *hit_record = *current;
			}
		}
	}
	else
	{
		bvh_traverse(mesh, node->left_node, ray, did_hit, hit_record, current, ignore, cnt);
		bvh_traverse(mesh, node->left_node + 1, ray, did_hit, hit_record, current, ignore, cnt);
	}
}

typedef struct
{
	const rto_scene* scene;
	int mesh_path;
	const float (*bounds)[6];   /* per mesh: min xyz, max xyz (slab-linear path) */
} world_t;

/* HitTest_TriangleMesh, Utils.h:290-327. */
static int hit_mesh(const world_t* w, int mi, const ray_t* ray, hit_t* hit_record, int ignore, uint64_t* cnt)
{
	const rto_mesh* mesh = &w->scene->meshes[mi];
	const rt_mesh_desc* m = &mesh->desc;
	hit_t closest = hit_default();
	int did_hit = 0;
	if (w->mesh_path == RTO_MESH_BVH)
	{
		bvh_traverse(mesh, 0, ray, &did_hit, hit_record, &closest, ignore, cnt);
		return did_hit;
	}
	CNT(ignore ? RT_CNT_SLAB_S_TEST : RT_CNT_SLAB_P_TEST);
	if (!slab_test(&w->bounds[mi][0], &w->bounds[mi][3], ray)) return 0;
	CNT(ignore ? RT_CNT_SLAB_S_PASS : RT_CNT_SLAB_P_PASS);
	for (int idx = 0; idx < 3 * m->triangle_count; idx += 3)
	{
		if (hit_triangle(mesh_pos(m, m->indices[idx]), mesh_pos(m, m->indices[idx + 1]),
		                 mesh_pos(m, m->indices[idx + 2]), mesh_nrm(m, idx / 3),
		                 m->cull_mode, m->material_index, ray, &closest, ignore, cnt))
		{
			if (ignore) return 1;
			if (closest.t < hit_record->t) *hit_record = closest;
			did_hit = 1;
		}
	}
	return did_hit;
}

/* Scene::GetClosestHit, Scene.cpp:29-66 (one scratch record shared by all primitives). */
static void get_closest_hit(const world_t* w, const ray_t* ray, hit_t* closest, uint64_t* cnt)
{
	const rto_scene* sc = w->scene;
	hit_t scratch = hit_default();
	for (int i = 0; i < sc->spheres.count; ++i)
	{
		if (hit_sphere(&sc->spheres, i, ray, &scratch, 0, cnt))
		{
			if (scratch.t < closest->t)
			{
				*closest = scratch;
				v3_normalize(&closest->normal);
				CNT(RT_CNT_SPHERE_P_CLOSEST);
			}
		}
	}
	for (int i = 0; i < sc->planes.count; ++i)
	{
		if (hit_plane(&sc->planes, i, ray, &scratch, 0, cnt))
		{
			if (scratch.t < closest->t) *closest = scratch;
		}
	}
	for (int i = 0; i < sc->mesh_count; ++i)
	{
		if (hit_mesh(w, i, ray, &scratch, 0, cnt))
		{
			if (scratch.t < closest->t) *closest = scratch;
		}
	}
}

/* Scene::DoesHit, Scene.cpp:68-96. */
static int does_hit(const world_t* w, const ray_t* ray, uint64_t* cnt)
{
	const rto_scene* sc = w->scene;
	hit_t temp = hit_default();
	for (int i = 0; i < sc->spheres.count; ++i) if (hit_sphere(&sc->spheres, i, ray, &temp, 1, cnt)) return 1;
	for (int i = 0; i < sc->planes.count; ++i) if (hit_plane(&sc->planes, i, ray, &temp, 1, cnt)) return 1;
	for (int i = 0; i < sc->mesh_count; ++i) if (hit_mesh(w, i, ray, &temp, 1, cnt)) return 1;
	return 0;
}

static const float RT_PI = 3.14159265358979323846f;   /* MathHelpers.h:7 */

/* BRDF::Lambert, BRDFs.h:14-22: (cd * kd) / PI per channel. */
static inline rgb lambert_scalar(float kd, rgb cd) { rgb o = { (cd.r * kd) / RT_PI, (cd.g * kd) / RT_PI, (cd.b * kd) / RT_PI }; return o; }
static inline rgb lambert_color(rgb kd, rgb cd) { rgb o = { (cd.r * kd.r) / RT_PI, (cd.g * kd.g) / RT_PI, (cd.b * kd.b) / RT_PI }; return o; }

/* BRDF::GeometryFunction_SchlickGGX, BRDFs.h:78-86. */
static inline float schlick_ggx(v3 n, v3 v, float roughness)
{
	const float a = roughness * roughness;
	const float k = ((a + 1) * (a + 1)) / 8;
	const float clamped = std_max(v3_dot(n, v), 0.f);
	return clamped / ((clamped * (1 - k)) + k);
}

/* Material::Shade x4, Material.h:41-44, 60-63, 83-87, 107-123; BRDFs.h:33-40, 49-53, 62-68, 96-99. */
static rgb shade(const rt_material_desc* m, v3 n, v3 l, v3 v, uint64_t* cnt)
{
	const rgb color = { m->color[0], m->color[1], m->color[2] };
	switch (m->tag)
	{
	case RT_MATERIAL_SOLID_COLOR:
		CNT(RT_CNT_SHADE_SOLID);
		return color;
	case RT_MATERIAL_LAMBERT:
		CNT(RT_CNT_SHADE_LAMBERT);
		return lambert_scalar(m->p0, color);
	case RT_MATERIAL_LAMBERT_PHONG:
	{
		CNT(RT_CNT_SHADE_PHONG);
		const rgb d = lambert_scalar(m->p0, color);
		const float nl = std_max(v3_dot(n, l), 0.f);
		const v3 reflect = v3_sub(l, v3_scale(n, 2 * nl));
		const float cosa = std_max(v3_dot(reflect, v), 0.f);
		const float spec = m->p1 * powf(cosa, m->p2);
		const rgb ph = { 1.f * spec, 1.f * spec, 1.f * spec };
		rgb o = { d.r + ph.r, d.g + ph.g, d.b + ph.b };
		return o;
	}
	case RT_MATERIAL_COOK_TORRENCE:
	{
		CNT(RT_CNT_SHADE_COOK_TORRENCE);
		const float metal = m->p0, rough = m->p1;
		v3 h = v3_add(v, l);
		v3_normalize(&h);
		rgb f0 = color;
		if (metal == 0.f) { f0.r = 0.04f; f0.g = 0.04f; f0.b = 0.04f; }
		const float pw = powf(1 - std_max(v3_dot(h, v), 0.f), 5);
		const rgb F = { f0.r + ((1.f - f0.r) * pw), f0.g + ((1.f - f0.g) * pw), f0.b + ((1.f - f0.b) * pw) };
		const float a = rough * rough;
		const float a2 = a * a;
		const float nh = std_max(v3_dot(n, h), 0.f);
		const float inner = (nh * nh) * ((a * a) - 1) + 1;
		const float D = a2 / (RT_PI * (inner * inner));
		const float G = schlick_ggx(n, v, rough) * schlick_ggx(n, l, rough);
		const float denom = 4 * std_max(v3_dot(v, n), 0.0001f) * std_max(v3_dot(l, n), 0.0001f);
		const rgb spec = { ((F.r * D) * G) / denom, ((F.g * D) * G) / denom, ((F.b * D) * G) / denom };
		rgb kd = { 0.f, 0.f, 0.f };
		if (metal == 0.f) { kd.r = 1.f - F.r; kd.g = 1.f - F.g; kd.b = 1.f - F.b; }
		const rgb diff = lambert_color(kd, color);
		rgb o = { diff.r + spec.r, diff.g + spec.g, diff.b + spec.b };
		return o;
	}
	default:
	{
		rgb z = { 0.f, 0.f, 0.f };
		return z;
	}
	}
}

/* Renderer::RenderPixel, Renderer.cpp:100-182. */
static uint32_t render_pixel(const world_t* w, const rt_camera* cam, const rt_frame_desc* fr, uint32_t pixel_index, uint64_t* cnt)
{
	const rto_scene* sc = w->scene;
	const int W = fr->width, H = fr->height;
	const int px = (int)(pixel_index % (uint32_t)W);
	const int py = (int)(pixel_index / (uint32_t)W);
	CNT(RT_CNT_PIXELS);

	const float cx = (2.f * ((px + 0.5f) / W) - 1) * fr->aspect_ratio * cam->fov;
	const float cy = (1.f - (2.f * (py + 0.5f) / H)) * cam->fov;

	/* Matrix::TransformVector(cx, cy, 1), Matrix.cpp:35-42 */
	v3 dir = v3_make(cam->right[0] * cx + cam->up[0] * cy + cam->forward[0] * 1.f,
	                 cam->right[1] * cx + cam->up[1] * cy + cam->forward[1] * 1.f,
	                 cam->right[2] * cx + cam->up[2] * cy + cam->forward[2] * 1.f);
	v3_normalize(&dir);
	const ray_t view = ray_make(v3_make(cam->origin[0], cam->origin[1], cam->origin[2]), dir, 0.0001f, FLT_MAX);

	hit_t closest = hit_default();
	get_closest_hit(w, &view, &closest, cnt);

	float shadow_factor = 1.f;
	rgb fc = { 0.f, 0.f, 0.f };
	if (closest.did_hit)
	{
		CNT(RT_CNT_HIT_PIXELS);
		const v3 origin_offset = v3_add(closest.origin, v3_scale(closest.normal, 0.0001f));
		const v3 view_neg = v3_make(-dir.x, -dir.y, -dir.z);
		for (int li = 0; li < sc->lights.count; ++li)
		{
			CNT(RT_CNT_LIGHT_ITERATIONS);
			const v3 lo = v3_make(sc->lights.origin_x[li], sc->lights.origin_y[li], sc->lights.origin_z[li]);
			const int ltype = sc->lights.type[li];
			/* LightUtils::GetDirectionToLight, Utils.h:341-353 */
			v3 ld = (ltype == RT_LIGHT_POINT || ltype == RT_LIGHT_DIRECTIONAL) ? v3_sub(lo, origin_offset) : v3_make(0, 0, 0);
			const float magnitude = v3_normalize(&ld);

			if (fr->shadows_enabled)
			{
				CNT(RT_CNT_SHADOW_RAYS);
				const ray_t shadow = ray_make(origin_offset, ld, 0.0001f, magnitude);
				if (does_hit(w, &shadow, cnt))
				{
					CNT(RT_CNT_OCCLUDED);
					shadow_factor *= 0.95f;
					continue;
				}
			}
			CNT(RT_CNT_LIT);

			/* LightUtils::GetRadiance, Utils.h:355-369 */
			rgb radiance = { 0.f, 0.f, 0.f };
			const rgb lc = { sc->lights.color_r[li], sc->lights.color_g[li], sc->lights.color_b[li] };
			const float intensity = sc->lights.intensity[li];

			switch (fr->lighting_mode)
			{
			case RT_LIGHTING_COMBINED:
			{
				const float oa = std_max(v3_dot(closest.normal, ld), 0.f);
				if (ltype == RT_LIGHT_POINT)
				{
					const float s = intensity / v3_sqr(v3_sub(lo, closest.origin));
					radiance.r = lc.r * s; radiance.g = lc.g * s; radiance.b = lc.b * s;
				}
				else if (ltype == RT_LIGHT_DIRECTIONAL)
				{
					radiance.r = lc.r * intensity; radiance.g = lc.g * intensity; radiance.b = lc.b * intensity;
				}
				const rgb brdf = shade(&sc->materials[closest.material], closest.normal, ld, view_neg, cnt);
				fc.r += (radiance.r * oa) * brdf.r;
				fc.g += (radiance.g * oa) * brdf.g;
				fc.b += (radiance.b * oa) * brdf.b;
				break;
			}
			case RT_LIGHTING_OBSERVED_AREA:
			{
				const float oa = std_max(v3_dot(closest.normal, ld), 0.f);
				fc.r += oa; fc.g += oa; fc.b += oa;
				break;
			}
			case RT_LIGHTING_RADIANCE:
			{
				if (ltype == RT_LIGHT_POINT)
				{
					const float s = intensity / v3_sqr(v3_sub(lo, closest.origin));
					radiance.r = lc.r * s; radiance.g = lc.g * s; radiance.b = lc.b * s;
				}
				else if (ltype == RT_LIGHT_DIRECTIONAL)
				{
					radiance.r = lc.r * intensity; radiance.g = lc.g * intensity; radiance.b = lc.b * intensity;
				}
				fc.r += radiance.r; fc.g += radiance.g; fc.b += radiance.b;
				break;
			}
			case RT_LIGHTING_BRDF:
			{
				const rgb brdf = shade(&sc->materials[closest.material], closest.normal, ld, view_neg, cnt);
				fc.r += brdf.r; fc.g += brdf.g; fc.b += brdf.b;
				break;
			}
			default: break;
			}
		}
		fc.r *= shadow_factor; fc.g *= shadow_factor; fc.b *= shadow_factor;
	}

	/* ColorRGB::MaxToOne, ColorRGB.h:12-17 */
	const float max_value = std_max(fc.r, std_max(fc.g, fc.b));
	if (max_value > 1.f) { fc.r /= max_value; fc.g /= max_value; fc.b /= max_value; }

	/* static_cast<uint8_t>(c * 255) + SDL_MapRGB, Renderer.cpp:178-181.  The conversion goes
	 * through int like the x86-64 code the reference compiles to (cvttss2si, low byte). */
	const uint8_t R = (uint8_t)(int)(fc.r * 255);
	const uint8_t G = (uint8_t)(int)(fc.g * 255);
	const uint8_t B = (uint8_t)(int)(fc.b * 255);
	return ((uint32_t)R << fr->r_shift) | ((uint32_t)G << fr->g_shift) | ((uint32_t)B << fr->b_shift) | fr->alpha_mask;
}

int rto_render_rows(const rto_scene* scene, const rt_camera* camera, const rt_frame_desc* frame,
                    int32_t mesh_path, int32_t row_begin, int32_t row_count, uint32_t* dst,
                    int32_t threads, rt_counters* counters)
{
	if (!scene || !camera || !frame || !dst) return 1;
	if (frame->width <= 0 || frame->height <= 0 || row_begin < 0 || row_count < 0 || row_begin + row_count > frame->height) return 1;
	if (mesh_path == RTO_MESH_BVH)
	{
		for (int i = 0; i < scene->mesh_count; ++i) if (!scene->meshes[i].desc.bvh_nodes) return 1;
	}

	float (*bounds)[6] = NULL;
	if (scene->mesh_count > 0)
	{
		bounds = (float (*)[6])malloc(sizeof(float[6]) * (size_t)scene->mesh_count);
		for (int i = 0; i < scene->mesh_count; ++i)
		{
			const rt_mesh_desc* m = &scene->meshes[i].desc;
			if (m->aabb_min && m->aabb_max) { memcpy(&bounds[i][0], m->aabb_min, 12); memcpy(&bounds[i][3], m->aabb_max, 12); }
			else rto_mesh_bounds(m, &bounds[i][0], &bounds[i][3]);
		}
	}
	world_t w; w.scene = scene; w.mesh_path = mesh_path; w.bounds = (const float (*)[6])bounds;

	const int64_t first = (int64_t)row_begin * frame->width;
	const int64_t count = (int64_t)row_count * frame->width;
	if (counters) memset(counters, 0, sizeof(*counters));

#ifdef _OPENMP
	if (threads > 0) omp_set_num_threads(threads);
#else
	(void)threads;
#endif
#pragma omp parallel
	{
		uint64_t local[RT_COUNTER_SLOTS];
		memset(local, 0, sizeof local);
		uint64_t* cnt = counters ? local : NULL;
#pragma omp for schedule(dynamic, 128)
		for (int64_t i = 0; i < count; ++i)
			dst[i] = render_pixel(&w, camera, frame, (uint32_t)(first + i), cnt);
		if (counters)
		{
#pragma omp critical
			for (int k = 0; k < RT_COUNTER_SLOTS; ++k) counters->slot[k] += local[k];
		}
	}
	free(bounds);
	return 0;
}

void rto_transform_mesh(const float* positions, int32_t vertex_count, const float* normals, int32_t triangle_count,
                        const float* m, float* out_positions, float* out_normals)
{
	/* rows: m[0..3] = data[0] (x axis), m[4..7] = data[1], m[8..11] = data[2], m[12..15] = data[3] (translation) */
	for (int i = 0; i < vertex_count; ++i)
	{
		const float x = positions[3 * i], y = positions[3 * i + 1], z = positions[3 * i + 2];
		/* Matrix::TransformPoint, Matrix.cpp:49-56 */
		out_positions[3 * i + 0] = m[0] * x + m[4] * y + m[8] * z + m[12];
		out_positions[3 * i + 1] = m[1] * x + m[5] * y + m[9] * z + m[13];
		out_positions[3 * i + 2] = m[2] * x + m[6] * y + m[10] * z + m[14];
	}
	for (int i = 0; i < triangle_count; ++i)
	{
		const float x = normals[3 * i], y = normals[3 * i + 1], z = normals[3 * i + 2];
		/* Matrix::TransformVector, Matrix.cpp:35-42, then Vector3::Normalized, Vector3.cpp:42-46 */
		v3 n = v3_make(m[0] * x + m[4] * y + m[8] * z, m[1] * x + m[5] * y + m[9] * z, m[2] * x + m[6] * y + m[10] * z);
		v3_normalize(&n);
		out_normals[3 * i] = n.x; out_normals[3 * i + 1] = n.y; out_normals[3 * i + 2] = n.z;
	}
}

/* ---- TriangleMesh::BuildBVH, source/DataTypes.h:294-483 (the shipped configuration: BVH + USE_BINS) ---- */

typedef struct bvh_build
{
	const float* tp;        /* transformedPositions */
	int32_t* indices;       /* reordered in place, DataTypes.h:358-360 */
	float* normals;         /* reordered in place, DataTypes.h:355 */
	float* tnormals;        /* transformedNormals, reordered in place, DataTypes.h:356 */
	rto_bvh_node* nodes;
	int32_t capacity;
	int32_t used;
	int overflow;
} bvh_build;

static v3 bb_position(const bvh_build* b, int32_t index) { return v3_make(b->tp[3 * index], b->tp[3 * index + 1], b->tp[3 * index + 2]); }
static float v3_axis(v3 v, int axis) { return axis == 0 ? v.x : (axis == 1 ? v.y : v.z); }
static float std_minf(float a, float b) { return (b < a) ? b : a; }   /* std::min */
static float std_maxf(float a, float b) { return (a < b) ? b : a; }   /* std::max */

/* centroid of the triangle whose first index sits at `i`, DataTypes.h:349, 412-416, 432-435 */
static v3 bb_centroid(const bvh_build* b, uint32_t i)
{
	const v3 sum = v3_add(v3_add(bb_position(b, b->indices[i]), bb_position(b, b->indices[i + 1])), bb_position(b, b->indices[i + 2]));
	return v3_scale(sum, 0.3333f);
}

/* AABB, DataTypes.h:56-80 */
typedef struct aabb { v3 mn, mx; } aabb;
static aabb aabb_empty(void) { aabb a; a.mn = v3_make(FLT_MAX, FLT_MAX, FLT_MAX); a.mx = v3_make(FLT_MIN, FLT_MIN, FLT_MIN); return a; }   /* Vector3.cpp:13-14 */
static void aabb_grow(aabb* a, v3 p)
{
	a->mn = v3_make(std_minf(a->mn.x, p.x), std_minf(a->mn.y, p.y), std_minf(a->mn.z, p.z));
	a->mx = v3_make(std_maxf(a->mx.x, p.x), std_maxf(a->mx.y, p.y), std_maxf(a->mx.z, p.z));
}
static void aabb_grow_box(aabb* a, const aabb* o)
{
	a->mn = v3_make(std_minf(a->mn.x, o->mn.x), std_minf(a->mn.y, o->mn.y), std_minf(a->mn.z, o->mn.z));
	a->mx = v3_make(std_maxf(a->mx.x, o->mx.x), std_maxf(a->mx.y, o->mx.y), std_maxf(a->mx.z, o->mx.z));
}
static float aabb_area(const aabb* a)
{
	const v3 e = v3_sub(a->mx, a->mn);
	return e.x * e.y + e.y * e.z + e.z * e.x;
}

/* UpdateNodeBounds, DataTypes.h:310-321 */
static void bb_update_bounds(bvh_build* b, uint32_t node_index)
{
	rto_bvh_node* n = &b->nodes[node_index];
	aabb box = aabb_empty();
	for (uint32_t i = n->first_idx; i < n->first_idx + n->idx_count; ++i) aabb_grow(&box, bb_position(b, b->indices[i]));
	n->min_aabb[0] = box.mn.x; n->min_aabb[1] = box.mn.y; n->min_aabb[2] = box.mn.z;
	n->max_aabb[0] = box.mx.x; n->max_aabb[1] = box.mx.y; n->max_aabb[2] = box.mx.z;
}

/* FindBestSplitPlane, DataTypes.h:398-483 */
static float bb_best_split(const bvh_build* b, const rto_bvh_node* n, int* axis, float* split_pos)
{
	float best = FLT_MAX;
	for (int a = 0; a < 3; ++a)
	{
		float lo = FLT_MAX, hi = FLT_MIN;
		for (uint32_t k = 0; k < n->idx_count; k += 3)
		{
			const float c = v3_axis(bb_centroid(b, n->first_idx + k), a);
			lo = std_minf(lo, c);
			hi = std_maxf(hi, c);
		}
		const float diff = hi - lo;
		if (fabsf(diff) < FLT_EPSILON) continue;

		aabb bin_box[8];
		uint32_t bin_count[8];
		for (int i = 0; i < 8; ++i) { bin_box[i] = aabb_empty(); bin_count[i] = 0; }
		float scale = 8 / diff;
		for (uint32_t k = 0; k < n->idx_count; k += 3)
		{
			const uint32_t at = n->first_idx + k;
			const float c = v3_axis(bb_centroid(b, at), a);
			int bin = (int)((c - lo) * scale);
			if (7 < bin) bin = 7;                                                     /* std::min(amountOfPlaneBins, ...) */
			bin_count[bin] += 3;
			aabb_grow(&bin_box[bin], bb_position(b, b->indices[at]));
			aabb_grow(&bin_box[bin], bb_position(b, b->indices[at + 1]));
			aabb_grow(&bin_box[bin], bb_position(b, b->indices[at + 2]));
		}

		float left_area[7], right_area[7];
		int left_count[7], right_count[7];
		int left_sum = 0, right_sum = 0;
		aabb left_box = aabb_empty(), right_box = aabb_empty();
		for (int i = 0; i < 7; ++i)
		{
			left_sum += (int)bin_count[i];
			left_count[i] = left_sum;
			aabb_grow_box(&left_box, &bin_box[i]);
			left_area[i] = aabb_area(&left_box);
			right_sum += (int)bin_count[7 - i];
			right_count[7 - i - 1] = right_sum;
			aabb_grow_box(&right_box, &bin_box[7 - i]);
			right_area[7 - i - 1] = aabb_area(&right_box);
		}
		scale = diff / 8;
		for (int i = 0; i < 7; ++i)
		{
			const float cost = left_count[i] * left_area[i] + right_count[i] * right_area[i];
			if (cost < best)
			{
				*axis = a;
				*split_pos = lo + scale * (i + 1);
				best = cost;
			}
		}
	}
	return best;
}

static void swap_v3_at(float* v, uint32_t a, uint32_t b)
{
	for (int k = 0; k < 3; ++k) { const float t = v[3 * a + k]; v[3 * a + k] = v[3 * b + k]; v[3 * b + k] = t; }
}

/* Subdivide, DataTypes.h:323-389 */
static void bb_subdivide(bvh_build* b, uint32_t node_index)
{
	rto_bvh_node* n = &b->nodes[node_index];
	if (n->idx_count <= 8) return;

	int axis = 0;
	float split_pos = 0.f;
	const float split_cost = bb_best_split(b, n, &axis, &split_pos);
	aabb box; box.mn = v3_make(n->min_aabb[0], n->min_aabb[1], n->min_aabb[2]); box.mx = v3_make(n->max_aabb[0], n->max_aabb[1], n->max_aabb[2]);
	const float no_split_cost = (float)n->idx_count * aabb_area(&box);                /* CalculateNodeCost, DataTypes.h:391-396 */
	if (split_cost >= no_split_cost) return;

	int i = (int)n->first_idx;
	int j = i + (int)n->idx_count - 1;
	while (i <= j)
	{
		if (v3_axis(bb_centroid(b, (uint32_t)i), axis) < split_pos) i += 3;
		else
		{
			swap_v3_at(b->normals, (uint32_t)(i / 3), (uint32_t)((j - 2) / 3));
			swap_v3_at(b->tnormals, (uint32_t)(i / 3), (uint32_t)((j - 2) / 3));
			for (int k = 0; k < 3; ++k) { const int32_t t = b->indices[i + k]; b->indices[i + k] = b->indices[j - 2 + k]; b->indices[j - 2 + k] = t; }
			j -= 3;
		}
	}
	const int left_count = i - (int)n->first_idx;
	if (left_count == 0 || left_count == (int)n->idx_count) return;
	if (b->used + 2 > b->capacity) { b->overflow = 1; return; }

	const uint32_t left = (uint32_t)b->used++, right = (uint32_t)b->used++;
	n->left_node = left;
	b->nodes[left].first_idx = n->first_idx;
	b->nodes[left].idx_count = (uint32_t)left_count;
	b->nodes[left].left_node = 0;
	b->nodes[right].first_idx = (uint32_t)i;
	b->nodes[right].idx_count = n->idx_count - (uint32_t)left_count;
	b->nodes[right].left_node = 0;
	n->idx_count = 0;
	bb_update_bounds(b, left);
	bb_update_bounds(b, right);
	bb_subdivide(b, left);
	bb_subdivide(b, right);
}

int32_t rto_update_transforms_bvh(const float* positions, int32_t vertex_count, int32_t* indices, float* normals,
                                  int32_t triangle_count, const float* transform, float* out_positions,
                                  float* out_normals, rto_bvh_node* out_nodes, int32_t node_capacity)
{
	if (node_capacity < 1) return -1;
	rto_transform_mesh(positions, vertex_count, normals, triangle_count, transform, out_positions, out_normals);
	bvh_build b;
	b.tp = out_positions; b.indices = indices; b.normals = normals; b.tnormals = out_normals;
	b.nodes = out_nodes; b.capacity = node_capacity; b.used = 1; b.overflow = 0;
	/* BuildBVH, DataTypes.h:294-308 */
	out_nodes[0].left_node = 0;
	out_nodes[0].first_idx = 0;
	out_nodes[0].idx_count = (uint32_t)(3 * triangle_count);
	bb_update_bounds(&b, 0);
	bb_subdivide(&b, 0);
	return b.overflow ? -1 : b.used;
}

uint64_t rto_fnv1a64(const void* data, uint64_t bytes)
{
	const unsigned char* p = (const unsigned char*)data;
	uint64_t h = 0xcbf29ce484222325ull;
	for (uint64_t i = 0; i < bytes; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
	return h;
}
