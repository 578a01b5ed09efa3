/*
 * TEST INFRASTRUCTURE -- CPU restatement of the reference's per-pixel path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker.  The product
 * (gp1_raytracer_2223_b200/csrc, include/rt_b200.h) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_port.py checks this restatement bit-for-bit
 * against frames rendered by the unmodified reference sources compiled in this repo's
 * build container (oracle/Makefile `ref`, fixtures under tests/golden/).  The reference
 * ships no golden vectors of its own.
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H

#include "../../include/rt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* BVH nodes travel inside rt_mesh_desc (bvh_nodes / bvh_node_count). */
typedef rt_bvh_node rto_bvh_node;

typedef struct rto_mesh
{
	rt_mesh_desc desc;
} rto_mesh;

typedef struct rto_scene
{
	rt_spheres_soa spheres;
	rt_planes_soa planes;
	rt_lights_soa lights;
	const rt_material_desc* materials;
	int32_t material_count;
	const rto_mesh* meshes;
	int32_t mesh_count;
} rto_scene;

enum rto_mesh_path
{
	RTO_MESH_SLAB_LINEAR = 0,    /* reference source/Utils.h:298-325 (#else branch): the north-star algorithm */
	RTO_MESH_BVH = 1             /* reference source/Utils.h:296-297 + 246-288: what the reference ships      */
};

/*
 * Renders rows [row_begin, row_begin + row_count) into dst (tightly packed).  threads <= 0
 * uses the OpenMP default.  counters may be NULL; when given it receives the test histogram
 * of SURVEY.md section 8(d) (enum rt_counter_slot); on the BVH path the slab slots stay 0 and
 * the node box tests are counted in RT_CNT_BVH_P_NODE / RT_CNT_BVH_S_NODE.
 * Returns 0 on success.
 */
int rto_render_rows(const rto_scene* scene, const rt_camera* camera, const rt_frame_desc* frame,
                    int32_t mesh_path, int32_t row_begin, int32_t row_count, uint32_t* dst,
                    int32_t threads, rt_counters* counters);

/* The box the slab-linear path tests: min/max over the indexed vertices, started from
 * +FLT_MAX / +FLT_MIN exactly like BVH root bounds (reference source/DataTypes.h:310-321,
 * source/Vector3.cpp:13-14), so that it equals the shipped build's root-node box. */
void rto_mesh_bounds(const rt_mesh_desc* mesh, float out_min[3], float out_max[3]);

/* TriangleMesh::UpdateTransforms without the BVH build (reference source/DataTypes.h:210-230):
 * out_positions[v] = finalTransform.TransformPoint(positions[v])                (source/Matrix.cpp:49-56)
 * out_normals[t]   = finalTransform.TransformVector(normals[t]).Normalized()    (source/Matrix.cpp:35-42)
 * transform = the 16 floats of Matrix::data[0..3] (x, y, z, w per row). */
void rto_transform_mesh(const float* positions, int32_t vertex_count, const float* normals, int32_t triangle_count,
                        const float* transform, float* out_positions, float* out_normals);

/* TriangleMesh::UpdateTransforms as the reference ships it, BuildBVH included (reference source/DataTypes.h:210-236,
 * 294-483, with BVH and USE_BINS defined, DataTypes.h:8-9).  `indices` (3 * T) and `normals` (3 * T floats) are
 * reordered IN PLACE exactly as the reference reorders TriangleMesh::indices / normals (DataTypes.h:344-363): the
 * next call starts from the order this one leaves, as in the reference.  out_positions (3 * V) and out_normals
 * (3 * T, final order) receive transformedPositions / transformedNormals, out_nodes the BVHNode array (capacity
 * node_capacity >= 2 * T is always enough).  Returns nodesUsed, or -1 when out_nodes is too small. */
int32_t rto_update_transforms_bvh(const float* positions, int32_t vertex_count, int32_t* indices, float* normals,
                                  int32_t triangle_count, const float* transform, float* out_positions,
                                  float* out_normals, rto_bvh_node* out_nodes, int32_t node_capacity);

uint64_t rto_fnv1a64(const void* data, uint64_t bytes);

#ifdef __cplusplus
}
#endif
#endif
