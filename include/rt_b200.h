/*
 * rt_b200.h -- C ABI of the B200-native per-pixel path for GP1_Raytracer_2223.
 *
 * The reference has no plugin / FFI seam: the boundary it offers is the C++
 * call `pRenderer->Render(pScene)` (reference source/main.cpp:91 ->
 * source/Renderer.cpp:34-98).  This header is the C-ABI a drop-in
 * `dae::Renderer` binds instead of running `RenderPixel`
 * (source/Renderer.cpp:100-182) on the host.  Every entry point names the
 * reference interface it replaces.  Plain pointers and sizes only; all
 * pointers are caller-owned and are only read (or, for destinations, written)
 * during the call.  One context per Renderer; calls on one context are
 * serialised by the caller, exactly like the reference's main thread
 * (source/main.cpp:56-111).
 *
 * There is no CPU fallback behind this interface: without a CUDA device
 * `rt_create` fails with RT_ERR_NO_DEVICE.
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_B200_ABI_VERSION 3

typedef struct rt_context rt_context;

/* Every call returns one of these; rt_last_error() gives the text. */
enum rt_status
{
	RT_OK = 0,
	RT_ERR_INVALID_ARGUMENT = 1,
	RT_ERR_CUDA = 2,
	RT_ERR_NO_DEVICE = 3,
	RT_ERR_BAD_STATE = 4,
	RT_ERR_CAPACITY = 5
};

/* Renderer::LightingMode, source/Renderer.h:40-48 (same numeric order, F3 cycles +1 mod 4). */
enum rt_lighting_mode
{
	RT_LIGHTING_OBSERVED_AREA = 0,
	RT_LIGHTING_RADIANCE = 1,
	RT_LIGHTING_BRDF = 2,
	RT_LIGHTING_COMBINED = 3
};

/* TriangleCullMode, source/DataTypes.h:29-34. */
enum rt_cull_mode
{
	RT_CULL_FRONT_FACE = 0,
	RT_CULL_BACK_FACE = 1,
	RT_CULL_NONE = 2
};

/* LightType, source/DataTypes.h:522-526. */
enum rt_light_type
{
	RT_LIGHT_POINT = 0,
	RT_LIGHT_DIRECTIONAL = 1
};

/*
 * Tag of the union that replaces the virtual Material::Shade
 * (source/Material.h:27): one value per concrete class.
 */
enum rt_material_tag
{
	RT_MATERIAL_SOLID_COLOR = 0,   /* Material_SolidColor   source/Material.h:34-48  : color            */
	RT_MATERIAL_LAMBERT = 1,       /* Material_Lambert      source/Material.h:54-68  : color, p0 = kd   */
	RT_MATERIAL_LAMBERT_PHONG = 2, /* Material_LambertPhong source/Material.h:74-94  : color, p0 = kd, p1 = ks, p2 = exponent */
	RT_MATERIAL_COOK_TORRENCE = 3  /* Material_CookTorrence source/Material.h:99-129 : color = albedo, p0 = metalness, p1 = roughness */
};

typedef struct rt_material_desc
{
	int32_t tag;      /* rt_material_tag */
	float color[3];
	float p0;
	float p1;
	float p2;
	float reserved;
} rt_material_desc;

/* Scene::GetSphereGeometries(), source/Scene.h:41; Sphere source/DataTypes.h:13-19, as SoA. */
typedef struct rt_spheres_soa
{
	const float* origin_x;
	const float* origin_y;
	const float* origin_z;
	const float* radius;
	const uint8_t* material_index;
	int32_t count;
} rt_spheres_soa;

/* Scene::GetPlaneGeometries(), source/Scene.h:40; Plane source/DataTypes.h:21-27, as SoA. */
typedef struct rt_planes_soa
{
	const float* origin_x;
	const float* origin_y;
	const float* origin_z;
	const float* normal_x;
	const float* normal_y;
	const float* normal_z;
	const uint8_t* material_index;
	int32_t count;
} rt_planes_soa;

/* Scene::GetLights(), source/Scene.h:42; Light source/DataTypes.h:528-536, as SoA. */
typedef struct rt_lights_soa
{
	const float* origin_x;
	const float* origin_y;
	const float* origin_z;
	const float* direction_x;   /* carried for completeness; the reference never reads it on this path (source/Utils.h:347-348) */
	const float* direction_y;
	const float* direction_z;
	const float* color_r;
	const float* color_g;
	const float* color_b;
	const float* intensity;
	const int32_t* type;        /* rt_light_type */
	int32_t count;
} rt_lights_soa;

/* BVHNode, source/DataTypes.h:43-54, exactly as TriangleMesh::BuildBVH (source/DataTypes.h:294-389)
 * leaves it in pBVHNodes: a leaf has idx_count > 0 and owns indices [first_idx, first_idx + idx_count);
 * an inner node's children are left_node and left_node + 1. */
typedef struct rt_bvh_node
{
	float min_aabb[3];
	float max_aabb[3];
	uint32_t first_idx;
	uint32_t idx_count;
	uint32_t left_node;
} rt_bvh_node;

/* Which of the reference's two HitTest_TriangleMesh bodies runs (source/Utils.h:296-325). */
enum rt_mesh_path
{
	RT_MESH_PATH_AUTO = 0,         /* BVH when every mesh came with nodes, else slab + linear */
	RT_MESH_PATH_SLAB_LINEAR = 1,  /* `#else` branch, source/Utils.h:298-325: mesh AABB, then every triangle */
	RT_MESH_PATH_BVH = 2           /* `#ifdef BVH` branch, source/Utils.h:296-297 + 246-288: what the reference ships */
};

/*
 * One TriangleMesh (source/DataTypes.h:109-156) after UpdateTransforms()
 * (source/DataTypes.h:210-236): world-space `transformedPositions`, `indices`,
 * and per-triangle `transformedNormals`.  Re-upload whenever UpdateTransforms
 * ran.  aabb_min/aabb_max may be NULL: the library then derives a box over the
 * indexed vertices (the shipped BVH build never fills
 * transformedMin/MaxAABB, source/DataTypes.h:231-235).
 */
typedef struct rt_mesh_desc
{
	const float* positions;     /* 3 * vertex_count, xyz interleaved as in std::vector<Vector3> */
	int32_t vertex_count;
	const int32_t* indices;     /* 3 * triangle_count */
	const float* normals;       /* 3 * triangle_count floats: one face normal per triangle */
	int32_t triangle_count;
	int32_t cull_mode;          /* rt_cull_mode */
	uint8_t material_index;
	const float* aabb_min;      /* 3 floats or NULL */
	const float* aabb_max;      /* 3 floats or NULL */
	const rt_bvh_node* bvh_nodes;  /* TriangleMesh::pBVHNodes (source/DataTypes.h:149) or NULL */
	int32_t bvh_node_count;        /* TriangleMesh::nodesUsed (source/DataTypes.h:151); 0 without nodes */
} rt_mesh_desc;

/*
 * What RenderPixel reads from dae::Camera (source/Camera.h:24-40) after
 * CalculateCameraToWorld() (source/Camera.h:43-53): origin, fov (already
 * tan(fovAngle/2), source/Camera.h:55-59) and rows 0..2 of cameraToWorld.
 * These stay host-computed so tanf/cosf/sinf come from the caller's libm.
 */
typedef struct rt_camera
{
	float origin[3];
	float fov;
	float right[3];
	float up[3];
	float forward[3];
} rt_camera;

/*
 * Renderer state read per frame (source/Renderer.h:49-61) plus the surface
 * format shifts SDL_MapRGB uses (source/Renderer.cpp:178-181).
 */
typedef struct rt_frame_desc
{
	int32_t width;              /* m_Width  */
	int32_t height;             /* m_Height */
	float aspect_ratio;         /* m_AspectRatio = width / float(height), source/Renderer.cpp:31 */
	int32_t lighting_mode;      /* rt_lighting_mode, m_CurrentLightingMode */
	int32_t shadows_enabled;    /* m_ShadowsEnabled */
	uint8_t r_shift;            /* SDL_PixelFormat::Rshift (16 for XRGB8888) */
	uint8_t g_shift;            /* 8 */
	uint8_t b_shift;            /* 0 */
	uint8_t reserved;
	uint32_t alpha_mask;        /* SDL_PixelFormat::Amask, OR-ed into every pixel (0 for XRGB8888) */
} rt_frame_desc;

/* Device-side timing of the most recent rt_render* call, from CUDA events. */
typedef struct rt_timing
{
	float kernel_ms;            /* pixel kernel, max over the context's devices */
	float gather_ms;            /* band gather to device 0 (0 with one device) */
	float d2h_ms;               /* device 0 -> host copy (0 for device-only renders) */
	float total_ms;             /* first launch to last completion */
	int32_t kernel_launches;    /* launches of our kernels during the call */
	int32_t reserved;
} rt_timing;

/*
 * Per-frame test histogram reproduced by the counters build of the kernel
 * (SURVEY.md section 8(d)): used only to turn kernel time into algorithmic
 * FLOP/s.  Indices documented in DESIGN.md.
 */
#define RT_COUNTER_SLOTS 40
typedef struct rt_counters
{
	uint64_t slot[RT_COUNTER_SLOTS];
} rt_counters;

/* Slot meanings; the FLOP weight of each event is the table in SURVEY.md section 8(d). */
enum rt_counter_slot
{
	RT_CNT_PIXELS = 0,              /* ray-gen 38, output stage 11 */
	RT_CNT_HIT_PIXELS = 1,          /* +6: offset origin */
	RT_CNT_SPHERE_P_DISC = 2,       /* primary sphere test rejected on the discriminant: 16 */
	RT_CNT_SPHERE_P_TREJ = 3,       /* rejected on t range: 19 */
	RT_CNT_SPHERE_P_HIT = 4,        /* hit, record written: 28 */
	RT_CNT_SPHERE_P_CLOSEST = 5,    /* became closest: +9 (normalise) */
	RT_CNT_SPHERE_S_DISC = 6,       /* shadow-ray sphere test: 16 */
	RT_CNT_SPHERE_S_TREJ = 7,       /* 19 */
	RT_CNT_SPHERE_S_HIT = 8,        /* 19 */
	RT_CNT_PLANE_P_TEST = 9,        /* 14 */
	RT_CNT_PLANE_P_HIT = 10,        /* +6 */
	RT_CNT_PLANE_S_TEST = 11,       /* 14 */
	RT_CNT_SLAB_P_TEST = 12,        /* 22 */
	RT_CNT_SLAB_P_PASS = 13,
	RT_CNT_SLAB_S_TEST = 14,        /* 22 */
	RT_CNT_SLAB_S_PASS = 15,
	RT_CNT_TRI_P_CULLED = 16,       /* parallel or culled: 5 */
	RT_CNT_TRI_P_DEGENERATE = 17,   /* |a| < eps: 25 */
	RT_CNT_TRI_P_UREJ = 18,         /* 35 */
	RT_CNT_TRI_P_VREJ = 19,         /* 51 */
	RT_CNT_TRI_P_TREJ = 20,         /* 57 */
	RT_CNT_TRI_P_HIT = 21,          /* 63 */
	RT_CNT_TRI_S_CULLED = 22,
	RT_CNT_TRI_S_DEGENERATE = 23,
	RT_CNT_TRI_S_UREJ = 24,
	RT_CNT_TRI_S_VREJ = 25,
	RT_CNT_TRI_S_TREJ = 26,
	RT_CNT_TRI_S_HIT = 27,
	RT_CNT_LIGHT_ITERATIONS = 28,   /* 15 each */
	RT_CNT_SHADOW_RAYS = 29,
	RT_CNT_OCCLUDED = 30,
	RT_CNT_LIT = 31,                /* un-shadowed light evaluations: 27 each in Combined mode */
	RT_CNT_SHADE_SOLID = 32,        /* 0 */
	RT_CNT_SHADE_LAMBERT = 33,      /* 6 */
	RT_CNT_SHADE_PHONG = 34,        /* 33 */
	RT_CNT_SHADE_COOK_TORRENCE = 35,/* 112 */
	RT_CNT_BVH_P_NODE = 36,         /* BVH path only: node box tests by primary rays, 22 each */
	RT_CNT_BVH_S_NODE = 37          /* BVH path only: node box tests by shadow rays */
};

/* ---- lifetime ------------------------------------------------------------------------------- */

/* Replaces `new Renderer(pWindow)` (source/Renderer.cpp:24-32) as far as device state goes.
 * device_ids == NULL or n_devices == 0 selects the current CUDA device only. */
int rt_create(const int32_t* device_ids, int32_t n_devices, rt_context** out_ctx);
int rt_destroy(rt_context* ctx);
const char* rt_last_error(const rt_context* ctx);   /* ctx may be NULL: creation errors */
int rt_abi_version(void);
int rt_device_count(const rt_context* ctx);

/* ---- scene upload (replaces the by-reference reads at source/Renderer.cpp:36-38 and
 *      source/Scene.cpp:29-96; call after Scene::Initialize and whenever the data changed) ------ */
int rt_upload_spheres(rt_context* ctx, const rt_spheres_soa* spheres);
int rt_upload_planes(rt_context* ctx, const rt_planes_soa* planes);
int rt_upload_lights(rt_context* ctx, const rt_lights_soa* lights);
int rt_upload_materials(rt_context* ctx, const rt_material_desc* materials, int32_t count);
int rt_set_mesh_count(rt_context* ctx, int32_t mesh_count);
/* Compile-time `#define BVH` of the reference (source/DataTypes.h:8, source/Utils.h:6) as a run-time
 * choice; default RT_MESH_PATH_AUTO.  RT_MESH_PATH_BVH fails at render time if a mesh has no nodes. */
int rt_set_mesh_path(rt_context* ctx, int32_t mesh_path);
int rt_upload_mesh(rt_context* ctx, int32_t mesh_id, const rt_mesh_desc* mesh);

/* ---- N1 of SURVEY.md 8(f): TriangleMesh::UpdateTransforms on the device ---------------------------------
 * Instead of re-uploading transformedPositions / transformedNormals every frame (rt_upload_mesh), upload the
 * UNtransformed mesh once (positions, indices, normals: source/DataTypes.h:133-135) and send only the final
 * transform each frame.  rt_transform_mesh runs source/DataTypes.h:216-230 on the device -
 * TransformPoint (source/Matrix.cpp:49-56) per vertex, TransformVector().Normalized()
 * (source/Matrix.cpp:35-42, source/Vector3.cpp:42-46) per face normal, same operation order - and rebuilds the
 * triangle stream and the mesh box there.  No BVH is built: such a mesh is rendered by the slab + linear body.
 * transform = the 16 floats of Matrix::data[0..3] of finalTransform = scale * rotation * translation. */
typedef struct rt_mesh_source
{
	const float* positions;     /* 3 * vertex_count, untransformed */
	int32_t vertex_count;
	const int32_t* indices;     /* 3 * triangle_count */
	const float* normals;       /* 3 * triangle_count floats, untransformed face normals */
	int32_t triangle_count;
	int32_t cull_mode;          /* rt_cull_mode */
	uint8_t material_index;
} rt_mesh_source;
int rt_upload_mesh_source(rt_context* ctx, int32_t mesh_id, const rt_mesh_source* source);
int rt_transform_mesh(rt_context* ctx, int32_t mesh_id, const float* transform);

/* The reference ships TriangleMesh::UpdateTransforms WITH BuildBVH (source/DataTypes.h:8-9, 231-232, 294-483).
 * rt_set_mesh_device_bvh(ctx, mesh, 1) - after rt_upload_mesh_source, before the mesh's first rt_transform_mesh -
 * moves that build to the device as well: from then on EVERY rt_transform_mesh call is one UpdateTransforms call
 * of the reference (transform + binned-SAH build + the in-place reordering of indices / normals,
 * DataTypes.h:344-363), executed in call order before the next frame, and the mesh is rendered by the BVH body.
 * The build is history-dependent exactly like the reference's (each one starts from the triangle order the
 * previous one left), so the source must be uploaded in the order TriangleMesh::indices / normals have at that
 * moment and calls must mirror the reference's one for one.  Node numbering is the device's own (pairs are
 * handed out in parallel); boxes, leaves, triangle order and the walk order are the reference's.
 * Errors: RT_ERR_BAD_STATE after the first transform, RT_ERR_CAPACITY beyond what node links can address; a leaf
 * wider than the link's triangle field fails the render that runs the build (RT_ERR_CAPACITY). */
int rt_set_mesh_device_bvh(rt_context* ctx, int32_t mesh_id, int32_t enable);

/* What the device-side builds left behind, for the caller that owns the TriangleMesh (its indices / normals
 * vectors go stale while the device runs UpdateTransforms; copy them back before handing the mesh to host code
 * again) and for the parity tests.  Pending rt_transform_mesh calls are executed first.
 *   indices  3 * triangle_count, TriangleMesh::indices after the last build        (may be NULL)
 *   normals  3 * triangle_count, TriangleMesh::normals (untransformed) in that order (may be NULL)
 *   nodes    the tree in the device's numbering (may be NULL); RT_ERR_CAPACITY when node_capacity is too small
 * A node is a leaf iff triangle_count > 0 (BVHNode::IsLeaf, source/DataTypes.h:50-53): then `first` is its first
 * TRIANGLE (BVHNode::firstIdx / 3); otherwise `first` is the left child and first + 1 the right one
 * (BVHNode::leftNode).  `escape` is the node IntersectionTest_BVH (source/Utils.h:246-288) visits after this
 * subtree, -1 at the end of the walk. */
typedef struct rt_built_node
{
	float min_aabb[3];
	float max_aabb[3];
	int32_t first;
	int32_t triangle_count;
	int32_t escape;
} rt_built_node;
int rt_read_mesh_build(rt_context* ctx, int32_t mesh_id, int32_t* indices, float* normals, rt_built_node* nodes, int32_t node_capacity, int32_t* out_node_count);

/* Which build of the pixel kernel renders frames.  All compute the same function, bit for bit.
 *   RT_KERNEL_SCALAR      one pixel per thread, one CTA per 32x8 pixel tile (also what rt_count_frame instruments)
 *   RT_KERNEL_PACKED      two pixels per thread on Blackwell's packed FP32 (FFMA2), one CTA per tile
 *   RT_KERNEL_PERSISTENT  one pixel per thread, persistent warps pulling 8x4 warp tiles off a device queue
 *                         (what RT_KERNEL_AUTO selects for large frames)
 *   RT_KERNEL_WAVEFRONT   rays, not pixels: five launches, one warp per (warp tile, mesh subtree) for the view rays and per
 *                         (warp tile, light, mesh subtree) for the shadow rays (source/Utils.h:246-288 walked subtree by
 *                         subtree, results merged with atomics); needs the BVH body over trees uploaded with rt_upload_mesh.
 *                         RT_KERNEL_AUTO selects it for small frames over deep trees; requested where it cannot run, AUTO's
 *                         choice is used */
enum rt_kernel_variant
{
	RT_KERNEL_AUTO = 0,
	RT_KERNEL_SCALAR = 1,
	RT_KERNEL_PACKED = 2,
	RT_KERNEL_PERSISTENT = 3,
	RT_KERNEL_WAVEFRONT = 4
};
int rt_set_kernel_variant(rt_context* ctx, int32_t variant);

/* ---- render ----------------------------------------------------------------------------------- */

/* Replaces Renderer::Render (source/Renderer.cpp:34-98): blocking; on return host_dst holds
 * height rows of width uint32 pixels, row stride pitch_bytes (>= 4*width), exactly what the
 * reference leaves in m_pBufferPixels.
 *   Every device renders the 8-row strips k, k + n, ... of the frame (k = its position in the context, n = the number of
 *   devices; one device: all of them) into its OWN full-frame buffer (width * height * 4 bytes per device) and copies
 *   exactly those strips to host_dst over its own PCIe link while later strips are still rendering: the kernel counts
 *   finished tiles per band of the frame, one of its CTAs reports complete bands through mapped pinned memory, and the
 *   calling thread issues each band's copy as it is reported ("direct present"; nothing is gathered on device 0).
 *   RT_B200_PRESENT=gather in the environment selects the older multi-device flow (peer stores into device 0's frame,
 *   device 0 presents), RT_B200_SINGLE_PRESENT=stream the older single-device one (copy stream with polled waits).
 * host_dst and pinning: when CUDA knows [host_dst, host_dst + pitch_bytes * (height - 1) + 4 * width) as pinned
 * memory (cudaHostAlloc / cudaHostRegister by the caller, or rt_register_surface below) the copies land in it
 * directly; otherwise they go through a pinned bounce buffer owned by the context and are memcpy'd out before
 * the call returns.  The library never registers caller memory on its own. */
int rt_render(rt_context* ctx, const rt_camera* camera, const rt_frame_desc* frame,
              uint32_t* host_dst, int32_t pitch_bytes);

/* Pin the surface rt_render* writes (the SDL surface's pixels, source/Renderer.cpp:27-29: they live as long as
 * the window) so that device-to-host copies target it directly.  LIFETIME RULE: the range must stay mapped, at
 * this address, until rt_unregister_surface(ctx, host_ptr) or rt_destroy(ctx) - freeing or re-mapping registered
 * memory is undefined behaviour in CUDA.  RT_ERR_CUDA when the range cannot be registered (the bounce buffer is
 * used then), RT_ERR_BAD_STATE for a pointer registered twice / never registered. */
int rt_register_surface(rt_context* ctx, void* host_ptr, size_t bytes);
int rt_unregister_surface(rt_context* ctx, void* host_ptr);

/* Same frame, left in device 0's frame buffer (no host copy): kernel-only measurement. */
int rt_render_device(rt_context* ctx, const rt_camera* camera, const rt_frame_desc* frame);

/* Copy the frame of the last rt_render_device out of device 0. */
int rt_download_frame(rt_context* ctx, uint32_t* host_dst, int32_t pitch_bytes);

/* Fill device 0's frame buffer (the one rt_render_device / rt_frame_export / rt_render_strips_to_frame(NULL) use)
 * with one pixel value; blocking.  Measurement and tests: a frame check must not pass on a stale frame. */
int rt_clear_frame(rt_context* ctx, uint32_t pixel);

/* Stream ordering of the *_device / *_to_frame calls that take a cuda_stream: the kernel is ordered after every
 * earlier scene upload (the caller's stream waits on the context's upload event), and every LATER scene upload /
 * rt_transform_mesh is ordered after that kernel (the context's stream waits on an event recorded behind it), so
 * a caller may enqueue frame k + 1's uploads without synchronising frame k's stream. */

/* One rank's share of a frame, for one-process-per-GPU launches: renders rows
 * [row_begin, row_begin + row_count) of the frame on the context's first device into
 * device_dst (tightly packed, row_count * width uint32) on `cuda_stream` (a cudaStream_t,
 * NULL = the context's own stream).  Asynchronous with respect to the host when a stream
 * is given. */
int rt_render_rows_device(rt_context* ctx, const rt_camera* camera, const rt_frame_desc* frame,
                          int32_t row_begin, int32_t row_count, void* device_dst, void* cuda_stream);

/* Same, for load-balanced interleaving: the frame is cut into strips of RT_STRIP_ROWS rows and
 * this call renders strips strip_first, strip_first + strip_step, ... (rank, world size) into
 * device_dst, packed in that order (every strip occupies RT_STRIP_ROWS * width pixels; the rows
 * of the last strip that fall outside the frame are left untouched). */
#define RT_STRIP_ROWS 8
int rt_render_strips_device(rt_context* ctx, const rt_camera* camera, const rt_frame_desc* frame,
                            int32_t strip_first, int32_t strip_step, void* device_dst, void* cuda_stream);

/* The gather's last step on the root rank: `device_src` holds `world` packed strip bands back to
 * back (each strips_per_rank * RT_STRIP_ROWS * width pixels, band r = strips r, r + world, ...);
 * writes the height x width frame to device_dst. */
int rt_unstripe_device(rt_context* ctx, const void* device_src, void* device_dst, int32_t width, int32_t height,
                       int32_t world, int32_t strips_per_rank, void* cuda_stream);

/* ---- one process per GPU, gather fused into the kernel ------------------------------------------
 * The root rank owns the frame; the other ranks map it (CUDA IPC, NVLink peer access) and their pixel
 * kernels store their strips straight into it: no gather step, no unstripe, only a barrier.
 *   root:   rt_frame_export(ctx, w, h, handle)      -> send the RT_IPC_HANDLE_BYTES to the other ranks
 *   others: rt_frame_import(ctx, handle, &ptr)      -> ptr addresses the root's frame from this process
 *   all:    rt_render_strips_to_frame(ctx, cam, frame, rank, world, ptr_or_NULL, stream)
 *           (NULL = this context's own frame, i.e. the root), then a barrier across ranks
 *   root:   rt_download_frame(ctx, host, pitch)     others: rt_frame_release(ctx, ptr) when done */
#define RT_IPC_HANDLE_BYTES 64
int rt_frame_export(rt_context* ctx, int32_t width, int32_t height, void* out_handle);
int rt_frame_import(rt_context* ctx, const void* handle, void** out_device_ptr);
int rt_frame_release(rt_context* ctx, void* device_ptr);
int rt_render_strips_to_frame(rt_context* ctx, const rt_camera* camera, const rt_frame_desc* frame,
                              int32_t strip_first, int32_t strip_step, void* frame_device_ptr, void* cuda_stream);
/* Completion signalling over the same peer mapping (instead of a collective): the frame allocation carries
 * a signal word behind its width * height pixels.  A non-root rank bumps it after its strips
 * (rt_frame_signal, stream-ordered after rt_render_strips_to_frame); the root makes its stream wait until
 * the word has reached `expected` (rt_frame_wait, cuStreamWaitValue32; the word only ever grows, so
 * expected = frames so far x (world - 1)).  RT_ERR_BAD_STATE if stream memory operations are unavailable. */
int rt_frame_signal(rt_context* ctx, void* frame_device_ptr, int32_t width, int32_t height, void* cuda_stream);
int rt_frame_wait(rt_context* ctx, uint32_t expected, void* cuda_stream);
/* Direct present for one process per GPU: this context's device renders strips strip_first, strip_first +
 * strip_step, ... (rank / world) and copies exactly those strips into the host surface, following its kernel band by
 * band over its OWN PCIe link; blocking.  host_dst is the whole surface (source/Renderer.cpp:27-29: the SDL surface's
 * pixels), e.g. a shared-memory mapping every rank opened: rows of other ranks' strips are not touched.  With N ranks
 * the frame crosses N links at once instead of all of it crossing GPU 0's (rt_frame_present).  In-process contexts
 * over several devices do the same inside rt_render.  Replaces nothing in the reference (it has one device: the CPU);
 * it is the multi-GPU form of the present at source/Renderer.cpp:97. */
int rt_render_strips_to_host(rt_context* ctx, const rt_camera* camera, const rt_frame_desc* frame,
                             int32_t strip_first, int32_t strip_step, uint32_t* host_dst, int32_t pitch_bytes);

/* Progressive present across ranks: with bands > 0 every CTA of rt_render_strips_to_frame_banded also bumps
 * the counter of the band (group of consecutive strips of the FRAME) it belongs to, in the root's trailer.
 * The root's rt_frame_present copies band after band to the host surface as soon as ALL ranks' CTAs of that
 * band are done (cuStreamWaitValue32 on a copy stream) and returns when the surface is complete.  Counters
 * only grow: frame_number counts the banded frames rendered into this frame buffer so far, starting at 1,
 * and must be the same on every rank's call. */
#define RT_MAX_PRESENT_BANDS 32
int rt_render_strips_to_frame_banded(rt_context* ctx, const rt_camera* camera, const rt_frame_desc* frame,
                                     int32_t strip_first, int32_t strip_step, void* frame_device_ptr,
                                     int32_t bands, void* cuda_stream);
int rt_frame_present(rt_context* ctx, uint32_t* host_dst, int32_t pitch_bytes, int32_t bands, uint32_t frame_number);

/* Host-side completion handshake of the direct present across processes: `words` points at one 64-bit arrival word per
 * rank in memory every rank maps (stride_words 64-bit words apart, e.g. 8 = one cache line each).  Stores `frame` into
 * this rank's word (release) and returns once every rank's word has reached `frame` (acquire): all strips of that frame
 * are in the shared surface.  No device work, no context: it only spares the caller an interpreted spin loop.
 * RT_ERR_BAD_STATE after timeout_seconds. */
int rt_host_arrive_and_wait(volatile int64_t* words, int32_t stride_words, int32_t rank, int32_t world, int64_t frame, double timeout_seconds);

int rt_get_timing(const rt_context* ctx, rt_timing* out_timing);

/* Counters build of the same kernel: fills the test histogram for one frame (slow path,
 * measurement only; the frame it renders is identical).  mesh_path selects which mesh body is
 * counted (RT_MESH_PATH_SLAB_LINEAR gives the algorithmic counts of SURVEY.md 8(d)). */
int rt_count_frame(rt_context* ctx, const rt_camera* camera, const rt_frame_desc* frame,
                   int32_t mesh_path, rt_counters* out_counters);

/* FP32 roofline denominator, measured on the device the context owns: a register-resident
 * chain of independent FP32 operations on every SM.  use_fma = 0 issues FMUL + FADD pairs
 * (what this path can use: the reference's arithmetic is unfused), 1 issues FFMA.  Returns
 * TFLOP/s (one FLOP per add or multiply, two per FMA) and the kernel time. */
int rt_measure_fp32_peak(rt_context* ctx, int32_t use_fma, double* out_tflops, float* out_ms);

#ifdef __cplusplus
}
#endif

#endif /* RT_B200_H */
