// Host side of the drop-in: walks a dae::Scene and hands it to the C ABI (include/rt_b200.h) as
// SoA arrays.  Compiled against the reference's UNMODIFIED headers (Scene.h, Material.h,
// DataTypes.h under /root/reference/source).
//
// Two pieces of reference state have no public accessor:
//   * Scene::m_TriangleMeshGeometries is protected            (source/Scene.h:50)
//   * the Material_* parameters are private, no type tag       (source/Material.h:46-47,65-67,89-93,125-128)
// This translation unit is therefore built with -fno-access-control (GCC/Clang) and reads the
// members directly, using RTTI for the class tag.  INTEGRATION.md lists the two additive
// accessors a maintainer would add instead; nothing in the reference's public API changes.
#pragma once

#include "Scene.h"
#include "Material.h"

#include "rt_b200.h"

#include <cstdint>
#include <string>
#include <array>
#include <cstring>
#include <limits>
#include <vector>

namespace rt_host
{
	struct FlatScene
	{
		// spheres / planes / lights, one array per field (what rt_*_soa point at)
		std::vector<float> sphere[4];          // ox, oy, oz, radius
		std::vector<uint8_t> sphere_material;
		std::vector<float> plane[6];           // ox, oy, oz, nx, ny, nz
		std::vector<uint8_t> plane_material;
		std::vector<float> light[10];          // ox, oy, oz, dx, dy, dz, r, g, b, intensity
		std::vector<int32_t> light_type;
		std::vector<rt_material_desc> materials;
		// device-side UpdateTransforms (RT_B200_DEVICE_TRANSFORM): 0 = host (upload its results), 1 = transform on the
		// device, 2 = transform + BuildBVH on the device; what was last sent as mesh source / transform
		int device_transform = 0;
		std::vector<size_t> source_vertices, source_triangles;
		std::vector<std::array<float, 16>> sent_transform;
	};

	inline bool DescribeMaterial(dae::Material* material, rt_material_desc& out)
	{
		out = rt_material_desc{};
		if (auto* m = dynamic_cast<dae::Material_SolidColor*>(material))
		{
			out.tag = RT_MATERIAL_SOLID_COLOR;
			out.color[0] = m->m_Color.r; out.color[1] = m->m_Color.g; out.color[2] = m->m_Color.b;
			return true;
		}
		if (auto* m = dynamic_cast<dae::Material_Lambert*>(material))
		{
			out.tag = RT_MATERIAL_LAMBERT;
			out.color[0] = m->m_DiffuseColor.r; out.color[1] = m->m_DiffuseColor.g; out.color[2] = m->m_DiffuseColor.b;
			out.p0 = m->m_DiffuseReflectance;
			return true;
		}
		if (auto* m = dynamic_cast<dae::Material_LambertPhong*>(material))
		{
			out.tag = RT_MATERIAL_LAMBERT_PHONG;
			out.color[0] = m->m_DiffuseColor.r; out.color[1] = m->m_DiffuseColor.g; out.color[2] = m->m_DiffuseColor.b;
			out.p0 = m->m_DiffuseReflectance; out.p1 = m->m_SpecularReflectance; out.p2 = m->m_PhongExponent;
			return true;
		}
		if (auto* m = dynamic_cast<dae::Material_CookTorrence*>(material))
		{
			out.tag = RT_MATERIAL_COOK_TORRENCE;
			out.color[0] = m->m_Albedo.r; out.color[1] = m->m_Albedo.g; out.color[2] = m->m_Albedo.b;
			out.p0 = m->m_Metalness; out.p1 = m->m_Roughness;
			return true;
		}
		return false;   // a Material subclass this path has no tag for
	}

	// Uploads everything Renderer::Render / RenderPixel read from the scene.  Returns an rt_status;
	// `why` receives the text on failure.
	inline int UploadScene(rt_context* ctx, dae::Scene* pScene, FlatScene& scratch, std::string& why)
	{
		auto fail = [&](int rc, const char* what) { why = std::string(what) + ": " + rt_last_error(ctx); return rc; };
		int rc;

		const auto& spheres = pScene->GetSphereGeometries();
		for (auto& v : scratch.sphere) v.clear();
		scratch.sphere_material.clear();
		for (const dae::Sphere& s : spheres)
		{
			scratch.sphere[0].push_back(s.origin.x); scratch.sphere[1].push_back(s.origin.y); scratch.sphere[2].push_back(s.origin.z);
			scratch.sphere[3].push_back(s.radius);
			scratch.sphere_material.push_back(s.materialIndex);
		}
		rt_spheres_soa ss{ scratch.sphere[0].data(), scratch.sphere[1].data(), scratch.sphere[2].data(), scratch.sphere[3].data(),
		                   scratch.sphere_material.data(), (int32_t)spheres.size() };
		if ((rc = rt_upload_spheres(ctx, &ss)) != RT_OK) return fail(rc, "rt_upload_spheres");

		const auto& planes = pScene->GetPlaneGeometries();
		for (auto& v : scratch.plane) v.clear();
		scratch.plane_material.clear();
		for (const dae::Plane& p : planes)
		{
			scratch.plane[0].push_back(p.origin.x); scratch.plane[1].push_back(p.origin.y); scratch.plane[2].push_back(p.origin.z);
			scratch.plane[3].push_back(p.normal.x); scratch.plane[4].push_back(p.normal.y); scratch.plane[5].push_back(p.normal.z);
			scratch.plane_material.push_back(p.materialIndex);
		}
		rt_planes_soa ps{ scratch.plane[0].data(), scratch.plane[1].data(), scratch.plane[2].data(), scratch.plane[3].data(),
		                  scratch.plane[4].data(), scratch.plane[5].data(), scratch.plane_material.data(), (int32_t)planes.size() };
		if ((rc = rt_upload_planes(ctx, &ps)) != RT_OK) return fail(rc, "rt_upload_planes");

		const auto& lights = pScene->GetLights();
		for (auto& v : scratch.light) v.clear();
		scratch.light_type.clear();
		for (const dae::Light& l : lights)
		{
			scratch.light[0].push_back(l.origin.x); scratch.light[1].push_back(l.origin.y); scratch.light[2].push_back(l.origin.z);
			scratch.light[3].push_back(l.direction.x); scratch.light[4].push_back(l.direction.y); scratch.light[5].push_back(l.direction.z);
			scratch.light[6].push_back(l.color.r); scratch.light[7].push_back(l.color.g); scratch.light[8].push_back(l.color.b);
			scratch.light[9].push_back(l.intensity);
			scratch.light_type.push_back((int32_t)l.type);
		}
		rt_lights_soa ls{ scratch.light[0].data(), scratch.light[1].data(), scratch.light[2].data(), scratch.light[3].data(),
		                  scratch.light[4].data(), scratch.light[5].data(), scratch.light[6].data(), scratch.light[7].data(),
		                  scratch.light[8].data(), scratch.light[9].data(), scratch.light_type.data(), (int32_t)lights.size() };
		if ((rc = rt_upload_lights(ctx, &ls)) != RT_OK) return fail(rc, "rt_upload_lights");

		scratch.materials.clear();
		for (dae::Material* m : pScene->m_Materials)
		{
			rt_material_desc d;
			if (!DescribeMaterial(m, d)) { why = "scene holds a Material subclass without a device tag"; return RT_ERR_INVALID_ARGUMENT; }
			scratch.materials.push_back(d);
		}
		if ((rc = rt_upload_materials(ctx, scratch.materials.data(), (int32_t)scratch.materials.size())) != RT_OK) return fail(rc, "rt_upload_materials");

		// Meshes: the outputs of TriangleMesh::UpdateTransforms (source/DataTypes.h:210-236), including
		// the BVH it rebuilt.  Vector3 is three packed floats, BVHNode matches rt_bvh_node field for field.
		static_assert(sizeof(dae::Vector3) == 3 * sizeof(float), "Vector3 must be three packed floats");
		static_assert(sizeof(dae::BVHNode) == sizeof(rt_bvh_node), "BVHNode layout differs from rt_bvh_node");
		auto& meshes = pScene->m_TriangleMeshGeometries;
		if ((rc = rt_set_mesh_count(ctx, (int32_t)meshes.size())) != RT_OK) return fail(rc, "rt_set_mesh_count");
		if (scratch.device_transform)
		{
			// the untransformed mesh crosses once, each frame only finalTransform = S * R * T
			// (source/DataTypes.h:213); TransformPoint / TransformVector().Normalized() run on the device and the
			// mesh is rendered by the slab + linear body (no BVH upload).  BuildBVH keeps permuting `indices` and
			// `normals` together on the host; any such order is the same triangle set, so the first one is kept.
			//
			// Mode 2 moves BuildBVH along (rt_set_mesh_device_bvh): the mesh is rendered by the BVH body, and every
			// CHANGE of finalTransform is one UpdateTransforms call of the reference, run on the device from the
			// triangle order the previous one left.  That is the reference's own sequence of builds when the host's
			// Scene::Update only sets the pose (RotateY / Translate / Scale) and leaves UpdateTransforms to this path
			// (INTEGRATION.md); the source is taken in the order TriangleMesh::indices has at the first Render.
			scratch.source_vertices.resize(meshes.size(), (size_t)-1);
			scratch.source_triangles.resize(meshes.size(), (size_t)-1);
			scratch.sent_transform.resize(meshes.size());
			for (size_t i = 0; i < meshes.size(); ++i)
			{
				const dae::TriangleMesh& m = meshes[i];
				if (scratch.source_vertices[i] != m.positions.size() || scratch.source_triangles[i] != m.indices.size() / 3)
				{
					rt_mesh_source src{};
					src.positions = m.positions.empty() ? nullptr : &m.positions[0].x;
					src.vertex_count = (int32_t)m.positions.size();
					src.indices = m.indices.data();
					src.normals = m.normals.empty() ? nullptr : &m.normals[0].x;
					src.triangle_count = (int32_t)(m.indices.size() / 3);
					src.cull_mode = (int32_t)m.cullMode;
					src.material_index = m.materialIndex;
					if ((rc = rt_upload_mesh_source(ctx, (int32_t)i, &src)) != RT_OK) return fail(rc, "rt_upload_mesh_source");
					if (scratch.device_transform == 2 && (rc = rt_set_mesh_device_bvh(ctx, (int32_t)i, 1)) != RT_OK) return fail(rc, "rt_set_mesh_device_bvh");
					scratch.source_vertices[i] = m.positions.size();
					scratch.source_triangles[i] = m.indices.size() / 3;
					scratch.sent_transform[i].fill(std::numeric_limits<float>::quiet_NaN());     // differs from every transform
				}
				const dae::Matrix finalTransform = m.scaleTransform * m.rotationTransform * m.translationTransform;
				float t[16];
				for (int r = 0; r < 4; ++r) { const dae::Vector4 row = finalTransform[r]; t[4 * r] = row.x; t[4 * r + 1] = row.y; t[4 * r + 2] = row.z; t[4 * r + 3] = row.w; }
				if (scratch.device_transform == 2)
				{
					if (std::memcmp(t, scratch.sent_transform[i].data(), sizeof t) == 0) continue;    // same pose: no new build
					std::memcpy(scratch.sent_transform[i].data(), t, sizeof t);
				}
				if ((rc = rt_transform_mesh(ctx, (int32_t)i, t)) != RT_OK) return fail(rc, "rt_transform_mesh");
			}
			return RT_OK;
		}
		for (size_t i = 0; i < meshes.size(); ++i)
		{
			const dae::TriangleMesh& m = meshes[i];
			rt_mesh_desc d{};
			d.positions = m.transformedPositions.empty() ? nullptr : &m.transformedPositions[0].x;
			d.vertex_count = (int32_t)m.transformedPositions.size();
			d.indices = m.indices.data();
			d.normals = m.transformedNormals.empty() ? nullptr : &m.transformedNormals[0].x;
			d.triangle_count = (int32_t)(m.indices.size() / 3);
			d.cull_mode = (int32_t)m.cullMode;
			d.material_index = m.materialIndex;
			d.aabb_min = nullptr; d.aabb_max = nullptr;   // transformedMin/MaxAABB are never filled in the shipped build
			d.bvh_nodes = reinterpret_cast<const rt_bvh_node*>(m.pBVHNodes);
			d.bvh_node_count = m.pBVHNodes ? (int32_t)m.nodesUsed : 0;
			if ((rc = rt_upload_mesh(ctx, (int32_t)i, &d)) != RT_OK) return fail(rc, "rt_upload_mesh");
		}
		return RT_OK;
	}
}
