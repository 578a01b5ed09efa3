// TEST INFRASTRUCTURE (oracle build only) -- not part of the shipped product.
//
// Headless driver around the UNMODIFIED reference sources.  It is compiled
// together with /root/reference/source/{Vector3,Vector4,Matrix,Scene,Renderer,
// Timer}.cpp (see oracle/Makefile) and does what the reference's main loop does
// (source/main.cpp:41-49, 88-91): build a window-sized surface, a Renderer and a
// Scene, call Scene::Initialize(), optionally Scene::Update(), then
// Renderer::Render(pScene).  Nothing here computes a pixel: frames come out of
// the reference's own Renderer::RenderPixel (source/Renderer.cpp:100-182).
//
// Outputs:
//   --out FILE         raw little-endian uint32 XRGB8888 frame (W*H*4 bytes)
//   --dump-scene FILE  the scene exactly as RenderPixel sees it, flattened to the
//                      "RTSC0001" layout read by gp1_raytracer_2223_b200/scene_file.py
//                      and oracle/port (fixtures under tests/golden/ come from this)
//   stdout             one JSON line: per-frame Render() times, thread count, FNV-1a-64
//
// Built with -fno-access-control so the dump can read Scene::m_TriangleMeshGeometries
// (protected, source/Scene.h:50) and the private material parameters
// (source/Material.h:46-47,65-67,89-93,125-128) without touching reference headers.
#include "sdl_stub.h"

#include "Renderer.h"
#include "Scene.h"
#include "Material.h"
#include "Timer.h"
#include "Utils.h"

#include <omp.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace dae;

namespace
{
	struct Options
	{
		std::string scene = "W4_Bunny";
		int width = 640;
		int height = 480;
		int mode = 3;          // LightingMode::Combined (source/Renderer.h:49)
		int shadows = 1;       // m_ShadowsEnabled default (source/Renderer.h:50)
		int frames = 1;
		int warmup = 0;
		int threads = 0;       // 0 = leave OpenMP default (all cores)
		bool haveTime = false;
		float time = 0.f;      // Timer::GetTotal() fed to Scene::Update
		bool haveMeshYaw = false;
		float meshYaw = 0.f;   // direct TriangleMesh::RotateY on every mesh
		bool haveCamOrigin = false;
		float camOrigin[3] = { 0, 0, 0 };
		bool haveCamRot = false;
		float camPitch = 0.f, camYaw = 0.f;
		bool haveFov = false;
		float fovAngle = 45.f;
		std::string out;
		std::string dump;
		std::string dumpSource;   // untransformed meshes + their final transform ("RTMS0001")
		std::vector<float> yawSteps;  // --yaw-steps a,b,c: RotateY(a); UpdateTransforms(); RotateY(b); UpdateTransforms(); ...
		std::string dumpSteps;    // the mesh state BEFORE those steps + the final transform of each step ("RTMP0001")
		bool stepsOnDevice = false; // --steps-on-device (drop-in build only): RotateY(a); Render(); RotateY(b); Render(); ... -
		                            // the host never runs UpdateTransforms for the steps, the drop-in's device side does
		std::string resources; // directory that CONTAINS "Resources/"
		float animateDt = 0.f; // --animate DT: every timed frame is Scene::Update(timer advanced by DT) + Render, as in
		                       // the reference's main loop (source/main.cpp:88-91); the two are timed separately
	};

	[[noreturn]] void Usage(const char* why)
	{
		std::fprintf(stderr,
			"ref_render: %s\n"
			"usage: ref_render --scene {W1|W2|W3|W3_Test|W4_Reference|W4_Bunny|W4_Optional}\n"
			"  [--width W --height H] [--mode 0..3] [--shadows 0|1] [--frames N] [--warmup N]\n"
			"  [--threads T] [--time SECONDS] [--mesh-yaw RAD] [--cam-origin X Y Z]\n"
			"  [--cam-rot PITCH YAW] [--fov DEGREES] [--out FILE] [--dump-scene FILE] [--dump-mesh-source FILE]\n"
			"  [--yaw-steps A,B,... [--dump-mesh-steps FILE] [--steps-on-device]]\n"
			"  [--resources DIR] [--animate DT_SECONDS]\n", why);
		std::exit(2);
	}

	Options Parse(int argc, char** argv)
	{
		Options o;
		auto need = [&](int i, int n) { if (i + n >= argc) Usage("missing value"); };
		for (int i = 1; i < argc; ++i)
		{
			const std::string a = argv[i];
			if (a == "--scene") { need(i, 1); o.scene = argv[++i]; }
			else if (a == "--width") { need(i, 1); o.width = std::atoi(argv[++i]); }
			else if (a == "--height") { need(i, 1); o.height = std::atoi(argv[++i]); }
			else if (a == "--mode") { need(i, 1); o.mode = std::atoi(argv[++i]); }
			else if (a == "--shadows") { need(i, 1); o.shadows = std::atoi(argv[++i]); }
			else if (a == "--frames") { need(i, 1); o.frames = std::atoi(argv[++i]); }
			else if (a == "--warmup") { need(i, 1); o.warmup = std::atoi(argv[++i]); }
			else if (a == "--threads") { need(i, 1); o.threads = std::atoi(argv[++i]); }
			else if (a == "--time") { need(i, 1); o.haveTime = true; o.time = std::strtof(argv[++i], nullptr); }
			else if (a == "--mesh-yaw") { need(i, 1); o.haveMeshYaw = true; o.meshYaw = std::strtof(argv[++i], nullptr); }
			else if (a == "--cam-origin") { need(i, 3); o.haveCamOrigin = true; for (int k = 0; k < 3; ++k) o.camOrigin[k] = std::strtof(argv[++i], nullptr); }
			else if (a == "--cam-rot") { need(i, 2); o.haveCamRot = true; o.camPitch = std::strtof(argv[++i], nullptr); o.camYaw = std::strtof(argv[++i], nullptr); }
			else if (a == "--fov") { need(i, 1); o.haveFov = true; o.fovAngle = std::strtof(argv[++i], nullptr); }
			else if (a == "--out") { need(i, 1); o.out = argv[++i]; }
			else if (a == "--dump-scene") { need(i, 1); o.dump = argv[++i]; }
			else if (a == "--dump-mesh-source") { need(i, 1); o.dumpSource = argv[++i]; }
			else if (a == "--dump-mesh-steps") { need(i, 1); o.dumpSteps = argv[++i]; }
			else if (a == "--yaw-steps")
			{
				need(i, 1);
				const std::string list = argv[++i];
				size_t pos = 0;
				while (pos <= list.size())
				{
					const size_t comma = list.find(',', pos);
					const std::string item = list.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
					if (!item.empty()) o.yawSteps.push_back((float)std::atof(item.c_str()));
					if (comma == std::string::npos) break;
					pos = comma + 1;
				}
			}
			else if (a == "--resources") { need(i, 1); o.resources = argv[++i]; }
			else if (a == "--steps-on-device") o.stepsOnDevice = true;
			else if (a == "--animate") { need(i, 1); o.animateDt = std::strtof(argv[++i], nullptr); }
			else Usage(("unknown argument " + a).c_str());
		}
		if (o.width <= 0 || o.height <= 0 || o.frames < 0 || o.mode < 0 || o.mode > 3) Usage("bad value");
		return o;
	}

	Scene* MakeScene(const std::string& name)
	{
		if (name == "W1") return new Scene_W1();
		if (name == "W2") return new Scene_W2();
		if (name == "W3") return new Scene_W3();
		if (name == "W3_Test") return new Scene_W3_TestScene();
		if (name == "W4_Reference") return new Scene_W4_ReferenceScene();
		if (name == "W4_Bunny") return new Scene_W4_BunnyScene();
		if (name == "W4_Optional") return new Scene_W4_OptionalScene();
		// Scene_W4_TestScene is not offered: in the shipped configuration it dereferences a null
		// pBVHNodes (source/Scene.cpp:306-310 vs source/DataTypes.h:231-232,296).
		Usage("unknown scene");
	}

	uint64_t Fnv1a64(const void* data, size_t bytes)
	{
		const unsigned char* p = static_cast<const unsigned char*>(data);
		uint64_t h = 0xcbf29ce484222325ull;
		for (size_t i = 0; i < bytes; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
		return h;
	}

	struct Writer
	{
		FILE* f;
		void I32(int32_t v) { std::fwrite(&v, 4, 1, f); }
		void U32(uint32_t v) { std::fwrite(&v, 4, 1, f); }
		void F32(float v) { std::fwrite(&v, 4, 1, f); }
		void V3(const Vector3& v) { F32(v.x); F32(v.y); F32(v.z); }
		void C3(const ColorRGB& c) { F32(c.r); F32(c.g); F32(c.b); }
	};

	// Flatten what RenderPixel / GetClosestHit / DoesHit read.  Layout: see scene_file.py.
	void DumpScene(const Options& o, Scene* pScene, float aspect, const char* path)
	{
		FILE* f = std::fopen(path, "wb");
		if (!f) { std::perror(path); std::exit(1); }
		Writer w{ f };
		std::fwrite("RTSC0001", 1, 8, f);
		w.I32(o.width); w.I32(o.height); w.I32(o.mode); w.I32(o.shadows);
		w.F32(aspect);

		Camera& cam = pScene->GetCamera();
		cam.CalculateCameraToWorld();
		w.V3(cam.origin);
		w.F32(cam.fov);
		w.V3(Vector3{ cam.cameraToWorld.data[0] });
		w.V3(Vector3{ cam.cameraToWorld.data[1] });
		w.V3(Vector3{ cam.cameraToWorld.data[2] });

		const auto& spheres = pScene->GetSphereGeometries();
		const auto& planes = pScene->GetPlaneGeometries();
		const auto& lights = pScene->GetLights();
		const auto& materials = pScene->m_Materials;
		const auto& meshes = pScene->m_TriangleMeshGeometries;
		w.I32((int32_t)spheres.size()); w.I32((int32_t)planes.size()); w.I32((int32_t)lights.size());
		w.I32((int32_t)materials.size()); w.I32((int32_t)meshes.size());

		for (const Sphere& s : spheres) { w.V3(s.origin); w.F32(s.radius); w.I32(s.materialIndex); }
		for (const Plane& p : planes) { w.V3(p.origin); w.V3(p.normal); w.I32(p.materialIndex); }
		for (const Light& l : lights) { w.V3(l.origin); w.V3(l.direction); w.C3(l.color); w.F32(l.intensity); w.I32((int32_t)l.type); }
		for (Material* m : materials)
		{
			int32_t tag = -1; ColorRGB c{}; float p0 = 0, p1 = 0, p2 = 0;
			if (auto* s = dynamic_cast<Material_SolidColor*>(m)) { tag = 0; c = s->m_Color; }
			else if (auto* l = dynamic_cast<Material_Lambert*>(m)) { tag = 1; c = l->m_DiffuseColor; p0 = l->m_DiffuseReflectance; }
			else if (auto* lp = dynamic_cast<Material_LambertPhong*>(m)) { tag = 2; c = lp->m_DiffuseColor; p0 = lp->m_DiffuseReflectance; p1 = lp->m_SpecularReflectance; p2 = lp->m_PhongExponent; }
			else if (auto* ct = dynamic_cast<Material_CookTorrence*>(m)) { tag = 3; c = ct->m_Albedo; p0 = ct->m_Metalness; p1 = ct->m_Roughness; }
			else { std::fprintf(stderr, "unknown material class\n"); std::exit(1); }
			w.I32(tag); w.C3(c); w.F32(p0); w.F32(p1); w.F32(p2); w.F32(0.f);
		}
		for (const TriangleMesh& m : meshes)
		{
			const int32_t nV = (int32_t)m.transformedPositions.size();
			const int32_t nT = (int32_t)(m.indices.size() / 3);
			const int32_t nNodes = m.pBVHNodes ? (int32_t)m.nodesUsed : 0;
			w.I32(nV); w.I32(nT); w.I32((int32_t)m.cullMode); w.I32(m.materialIndex); w.I32(nNodes);
			for (const Vector3& p : m.transformedPositions) w.V3(p);
			for (int idx : m.indices) w.I32(idx);
			for (const Vector3& n : m.transformedNormals) w.V3(n);
			for (int32_t i = 0; i < nNodes; ++i)
			{
				const BVHNode& n = m.pBVHNodes[i];
				w.V3(n.minAABB); w.V3(n.maxAABB); w.U32(n.firstIdx); w.U32(n.idxCount); w.U32(n.leftNode);
			}
		}
		std::fclose(f);
	}
}

namespace
{
	// What TriangleMesh::UpdateTransforms consumes (source/DataTypes.h:210-230): the untransformed positions
	// and face normals (in their current order), the indices, and finalTransform = S * R * T.
	void DumpMeshSources(Scene* pScene, const char* path)
	{
		FILE* f = std::fopen(path, "wb");
		if (!f) { std::perror(path); std::exit(1); }
		Writer w{ f };
		std::fwrite("RTMS0001", 1, 8, f);
		const auto& meshes = pScene->m_TriangleMeshGeometries;
		w.I32((int32_t)meshes.size());
		for (const TriangleMesh& m : meshes)
		{
			w.I32((int32_t)m.positions.size()); w.I32((int32_t)(m.indices.size() / 3));
			for (const Vector3& p : m.positions) w.V3(p);
			for (int idx : m.indices) w.I32(idx);
			for (const Vector3& n : m.normals) w.V3(n);
			const Matrix finalTransform = m.scaleTransform * m.rotationTransform * m.translationTransform;   // DataTypes.h:213
			for (int r = 0; r < 4; ++r) { w.F32(finalTransform.data[r].x); w.F32(finalTransform.data[r].y); w.F32(finalTransform.data[r].z); w.F32(finalTransform.data[r].w); }
		}
		std::fclose(f);
	}
}

namespace
{
	// Input of a sequence of TriangleMesh::UpdateTransforms calls (source/DataTypes.h:210-236, BuildBVH
	// included): the meshes as they are BEFORE the first call - BuildBVH reorders indices and normals in
	// place (DataTypes.h:344-363), so every build starts from the order the previous one left - and the
	// finalTransform of every call.  The matching output is the --dump-scene of the same run.
	struct StepRecorder
	{
		struct MeshBefore { std::vector<Vector3> positions, normals; std::vector<int> indices; };
		std::vector<MeshBefore> before;
		std::vector<std::vector<Matrix>> transforms;   // [step][mesh]
		void Capture(Scene* pScene)
		{
			for (const TriangleMesh& m : pScene->m_TriangleMeshGeometries) before.push_back({ m.positions, m.normals, m.indices });
		}
		void Step(Scene* pScene)
		{
			transforms.emplace_back();
			for (const TriangleMesh& m : pScene->m_TriangleMeshGeometries)
				transforms.back().push_back(m.scaleTransform * m.rotationTransform * m.translationTransform);   // DataTypes.h:213
		}
		void Write(const char* path) const
		{
			FILE* f = std::fopen(path, "wb");
			if (!f) { std::perror(path); std::exit(1); }
			Writer w{ f };
			std::fwrite("RTMP0001", 1, 8, f);
			w.I32((int32_t)before.size()); w.I32((int32_t)transforms.size());
			for (size_t k = 0; k < before.size(); ++k)
			{
				const MeshBefore& m = before[k];
				w.I32((int32_t)m.positions.size()); w.I32((int32_t)(m.indices.size() / 3));
				for (const Vector3& p : m.positions) w.V3(p);
				for (int idx : m.indices) w.I32(idx);
				for (const Vector3& n : m.normals) w.V3(n);
				for (size_t s = 0; s < transforms.size(); ++s)
				{
					const Matrix& t = transforms[s][k];
					for (int r = 0; r < 4; ++r) { w.F32(t.data[r].x); w.F32(t.data[r].y); w.F32(t.data[r].z); w.F32(t.data[r].w); }
				}
			}
			std::fclose(f);
		}
	};
}

int main(int argc, char** argv)
{
	const Options o = Parse(argc, argv);

	// The scenes open "Resources/<name>.obj" relative to the cwd (source/Scene.cpp:307,413,450).
	std::string res = o.resources;
	if (res.empty())
	{
		char exe[4096];
		const ssize_t n = readlink("/proc/self/exe", exe, sizeof(exe) - 1);
		if (n > 0) { exe[n] = 0; res = exe; res = res.substr(0, res.find_last_of('/')); }
	}
	std::string out = o.out, dump = o.dump, dumpSource = o.dumpSource, dumpSteps = o.dumpSteps;
	auto absolutise = [](std::string& p)
	{
		if (!p.empty() && p[0] != '/') { char cwd[4096]; if (getcwd(cwd, sizeof cwd)) p = std::string(cwd) + "/" + p; }
	};
	absolutise(out); absolutise(dump); absolutise(dumpSource); absolutise(dumpSteps);
	if (!res.empty() && chdir(res.c_str()) != 0) { std::perror(res.c_str()); return 1; }

	if (o.threads > 0) omp_set_num_threads(o.threads);

	SDL_Window* pWindow = GP1_CreateHeadlessWindow(o.width, o.height);
	Timer* pTimer = new Timer();
	Renderer* pRenderer = new Renderer(pWindow);
	Scene* pScene = MakeScene(o.scene);
	pScene->Initialize();

	// Renderer starts at Combined + shadows on (source/Renderer.h:49-50); reach the requested
	// state through the same public toggles F3/F2 drive (source/main.cpp:67-78).
	for (int m = 3; m != o.mode; m = (m + 1) % 4) pRenderer->CycleLightingMode();
	if (!o.shadows) pRenderer->ToggleShadows();

	Camera& cam = pScene->GetCamera();
	if (o.haveFov) cam.SetCameraFOV(o.fovAngle);
	if (o.haveTime)
	{
		pTimer->m_TotalTime = o.time;   // what Timer::GetTotal() returns (source/Timer.h:30)
		pTimer->m_ElapsedTime = 0.f;
		pScene->Update(pTimer);          // mesh yaw from the timer (source/Scene.cpp:391-400, 431-437)
	}
	if (o.haveMeshYaw)
	{
		for (TriangleMesh& m : pScene->m_TriangleMeshGeometries) { m.RotateY(o.meshYaw); m.UpdateTransforms(); }
	}
	if (!o.yawSteps.empty())
	{
		StepRecorder rec;
		rec.Capture(pScene);
		for (float yaw : o.yawSteps)
		{
#ifdef GP1_DROPIN
			if (o.stepsOnDevice)
			{
				// what a host whose Scene::Update no longer calls UpdateTransforms does: set the pose, render
				// (RT_B200_DEVICE_TRANSFORM=2: the drop-in runs UpdateTransforms + BuildBVH on the device)
				for (TriangleMesh& m : pScene->m_TriangleMeshGeometries) m.RotateY(yaw);
				pRenderer->Render(pScene);
				continue;
			}
#endif
			for (TriangleMesh& m : pScene->m_TriangleMeshGeometries) { m.RotateY(yaw); m.UpdateTransforms(); }
			rec.Step(pScene);
		}
		if (!dumpSteps.empty()) rec.Write(dumpSteps.c_str());
	}
	if (o.haveCamOrigin) cam.origin = { o.camOrigin[0], o.camOrigin[1], o.camOrigin[2] };
	if (o.haveCamRot)
	{
		cam.totalPitch = o.camPitch; cam.totalYaw = o.camYaw;
		cam.CalculateForwardVector();    // source/Camera.h:61-66
	}

	std::vector<double> ms, updateMs;
	auto update = [&]()
	{
		// source/main.cpp:88: pScene->Update(pTimer) with the timer one frame further (mesh yaw = PI/2 * total time on the
		// W4 scenes, source/Scene.cpp:391-400, 431-437, 468-474: RotateY + UpdateTransforms incl. BuildBVH)
		if (o.animateDt <= 0.f) return;
		const auto u0 = std::chrono::steady_clock::now();
		pTimer->m_TotalTime += o.animateDt;
		pTimer->m_ElapsedTime = o.animateDt;
		pScene->Update(pTimer);
		updateMs.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - u0).count());
	};
	for (int i = 0; i < o.warmup; ++i) { update(); pRenderer->Render(pScene); }
	updateMs.clear();
	for (int i = 0; i < o.frames; ++i)
	{
		update();
		const auto t0 = std::chrono::steady_clock::now();
		pRenderer->Render(pScene);
		const auto t1 = std::chrono::steady_clock::now();
		ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
	}

	const size_t bytes = (size_t)o.width * (size_t)o.height * 4u;
	const uint64_t hash = Fnv1a64(pWindow->surface.pixels, bytes);
	if (!out.empty())
	{
		FILE* f = std::fopen(out.c_str(), "wb");
		if (!f) { std::perror(out.c_str()); return 1; }
		std::fwrite(pWindow->surface.pixels, 1, bytes, f);
		std::fclose(f);
	}
	if (!dump.empty()) DumpScene(o, pScene, pRenderer->m_AspectRatio, dump.c_str());
	if (!dumpSource.empty()) DumpMeshSources(pScene, dumpSource.c_str());

	std::vector<double> sorted = ms;
	std::sort(sorted.begin(), sorted.end());
	const double median = sorted.empty() ? 0.0 : sorted[sorted.size() / 2];
	std::printf("{\"scene\": \"%s\", \"width\": %d, \"height\": %d, \"mode\": %d, \"shadows\": %d, "
		"\"threads\": %d, \"frames\": %d, \"warmup\": %d, \"ms_median\": %.6f, \"ms_min\": %.6f, \"ms\": [",
		o.scene.c_str(), o.width, o.height, o.mode, o.shadows, omp_get_max_threads(), o.frames, o.warmup,
		median, sorted.empty() ? 0.0 : sorted.front());
	for (size_t i = 0; i < ms.size(); ++i) std::printf("%s%.6f", i ? ", " : "", ms[i]);
	if (!updateMs.empty())
	{
		std::printf("], \"update_ms\": [");
		for (size_t i = 0; i < updateMs.size(); ++i) std::printf("%s%.6f", i ? ", " : "", updateMs[i]);
	}
#ifdef GP1_DROPIN
	const char* path = "drop-in Renderer -> librt_b200.so (B200)";
#else
	const char* path = "reference BVH traversal (source/Utils.h:296-297)";
#endif
	std::printf("], \"fnv1a64\": \"%016llx\", \"path\": \"%s\"}\n", (unsigned long long)hash, path);

	delete pScene;
	delete pRenderer;
	delete pTimer;
	GP1_DestroyHeadlessWindow(pWindow);
	return 0;
}
