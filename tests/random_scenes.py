"""Seeded random scenes for parity tests (test infrastructure): spheres, planes, triangle meshes with every
cull mode, every material class, point and directional lights, plus a simple median-split BVH builder that
emits nodes in the reference's BVHNode format (any valid tree exercises the same traversal code)."""
import numpy as np

from gp1_raytracer_2223_b200.scene_file import BVH_NODE_DTYPE, MATERIAL_DTYPE, Camera, FlatScene, Mesh


def build_bvh(positions, indices, normals, leaf_triangles=2):
    """Reorders (indices, normals) in place like TriangleMesh::Subdivide does and returns BVHNode records:
    children at left_node, left_node + 1; leaves own contiguous index ranges."""
    idx = indices.copy()
    nrm = normals.copy()
    nodes = []

    def bounds(first, count):
        p = positions[idx[first:first + count].reshape(-1)]
        return p.min(0), p.max(0)

    def make(first, count):
        lo, hi = bounds(first, count)
        nodes.append([lo, hi, 3 * first, 3 * count, 0])
        return len(nodes) - 1

    def subdivide(n):
        first, count = nodes[n][2] // 3, nodes[n][3] // 3
        if count <= leaf_triangles:
            return
        cent = positions[idx[first:first + count]].mean(1)
        axis = int(np.argmax(cent.max(0) - cent.min(0)))
        order = np.argsort(cent[:, axis], kind="stable")
        idx[first:first + count] = idx[first:first + count][order]
        nrm[first:first + count] = nrm[first:first + count][order]
        half = count // 2
        left = make(first, half)
        right = make(first + half, count - half)
        assert right == left + 1
        nodes[n][3] = 0
        nodes[n][4] = left
        subdivide(left)
        subdivide(right)

    root = make(0, len(idx))
    subdivide(root)
    out = np.zeros(len(nodes), dtype=BVH_NODE_DTYPE)
    for i, (lo, hi, first_idx, idx_count, left) in enumerate(nodes):
        out[i] = (lo, hi, first_idx, idx_count, left)
    return idx, nrm, out


def random_scene(seed, n_spheres=4, n_planes=4, n_meshes=2, n_triangles=40, n_lights=3, pow_materials=True,
                 with_bvh=True):
    rng = np.random.default_rng(seed)
    f32 = np.float32

    def soa(a):
        return np.ascontiguousarray(np.asarray(a, dtype=f32).reshape(-1, 3).T)

    mats = np.zeros(8, dtype=MATERIAL_DTYPE)
    tags = [0, 1, 1, 0, 1, 1, 0, 1] if not pow_materials else [0, 1, 2, 3, 3, 2, 1, 3]
    for i, tag in enumerate(tags):
        mats[i]["tag"] = tag
        mats[i]["color"] = rng.uniform(0.1, 1.0, 3)
        if tag == 1:
            mats[i]["p0"] = rng.uniform(0.3, 1.0)
        elif tag == 2:
            mats[i]["p0"], mats[i]["p1"], mats[i]["p2"] = rng.uniform(0.3, 1.0), rng.uniform(0.1, 1.0), rng.choice([1.0, 3.0, 15.0, 60.0])
        elif tag == 3:
            mats[i]["p0"], mats[i]["p1"] = rng.choice([0.0, 1.0]), rng.choice([0.1, 0.6, 1.0])

    sph_o = rng.uniform([-4, 0, -2], [4, 5, 6], (n_spheres, 3))
    sph_r = rng.uniform(0.3, 1.2, n_spheres)
    # an open box of planes around the action plus random tilted ones
    planes_o = [(0, -0.5, 0), (0, 0, 12), (-7, 0, 0), (7, 0, 0)][:n_planes]
    planes_n = [(0, 1, 0), (0, 0, -1), (1, 0, 0), (-1, 0, 0)][:n_planes]
    for _ in range(max(0, n_planes - 4)):
        planes_o.append(tuple(rng.uniform(-6, 6, 3)))
        planes_n.append(tuple(rng.normal(size=3)))      # un-normalised on purpose: the reference never normalises them

    meshes = []
    for m in range(n_meshes):
        centre = rng.uniform([-3, 0.5, 0], [3, 4, 5], 3)
        verts = (centre + rng.normal(scale=0.9, size=(n_triangles + 2, 3))).astype(f32)
        idx = np.stack([np.arange(n_triangles), np.arange(n_triangles) + 1, np.arange(n_triangles) + 2], 1).astype(np.int32)
        e1, e2 = verts[idx[:, 1]] - verts[idx[:, 0]], verts[idx[:, 2]] - verts[idx[:, 0]]
        nrm = np.cross(e1, e2)
        nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(f32)
        nodes = None
        if with_bvh:
            idx, nrm, nodes = build_bvh(verts, idx, nrm, leaf_triangles=int(rng.integers(1, 5)))
        meshes.append(Mesh(verts, idx, nrm, cull_mode=m % 3, material_index=int(rng.integers(0, 8)), bvh_nodes=nodes))

    light_o = rng.uniform([-5, 3, -6], [5, 8, 6], (n_lights, 3))
    light_c = rng.uniform(0.3, 1.0, (n_lights, 3))
    light_t = np.array([0, 0, 1, 0, 1][:n_lights], dtype=np.int32)            # a directional light too
    light_i = np.where(light_t == 0, rng.uniform(20, 80, n_lights), rng.uniform(0.2, 0.8, n_lights)).astype(f32)

    fwd = np.array([0.1, -0.15, 1.0]); fwd /= np.linalg.norm(fwd)
    right = np.cross([0, 1, 0], fwd); right /= np.linalg.norm(right)
    up = np.cross(fwd, right); up /= np.linalg.norm(up)
    cam = Camera(np.array([0.3, 2.5, -9.0], f32), float(np.tan(np.radians(50.0) / 2)), right.astype(f32), up.astype(f32), fwd.astype(f32))
    return FlatScene(
        sphere_origin=soa(sph_o), sphere_radius=sph_r.astype(f32), sphere_material=rng.integers(0, 8, n_spheres).astype(np.uint8),
        plane_origin=soa(planes_o), plane_normal=soa(planes_n), plane_material=rng.integers(0, 8, len(planes_o)).astype(np.uint8),
        light_origin=soa(light_o), light_direction=soa(np.zeros((n_lights, 3))), light_color=soa(light_c),
        light_intensity=light_i, light_type=light_t, materials=mats, meshes=meshes, camera=cam)
