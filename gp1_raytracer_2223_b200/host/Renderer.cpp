// Drop-in replacement for the reference's source/Renderer.cpp: the same dae::Renderer class
// (declared by the reference's own, unmodified Renderer.h), with the pixel loop of
// Renderer::Render handed to the B200 through the C ABI (include/rt_b200.h).
//
// What stays as in the reference: the constructor grabs the window surface
// (source/Renderer.cpp:24-32), Render refreshes cameraToWorld and presents the surface
// (source/Renderer.cpp:40,97), CycleLightingMode / ToggleShadows / SaveBufferToImage keep their
// meaning.  What changes: instead of fanning RenderPixel over W*H pixels with
// concurrency::parallel_for (source/Renderer.cpp:79-85), Render uploads what RenderPixel would
// read and calls rt_render, which fills m_pBufferPixels.
//
// Renderer.h fixes the class layout, so the device context lives in a side table keyed by the
// Renderer (created on the first Render, released at process exit).  Devices: environment variable
// RT_B200_DEVICES="0,1,2,3" (default: the current device).  RT_B200_DEVICE_TRANSFORM=1 moves
// TriangleMesh::UpdateTransforms' vertex / normal transform to the device as well (mesh source uploaded once,
// 64 bytes per mesh and frame; rendered by the slab + linear body); RT_B200_DEVICE_TRANSFORM=2 moves BuildBVH there
// too (one build per change of pose; rendered by the BVH body).  Errors have no channel in the reference
// API (Render returns void): they are printed to stderr and the frame is left untouched; there is
// no CPU fallback.
#include "SDL.h"
#include "SDL_surface.h"

#include "Renderer.h"
#include "Math.h"
#include "Matrix.h"
#include "Material.h"
#include "Scene.h"
#include "Utils.h"

#include "SceneFlattener.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

using namespace dae;

namespace
{
	struct DeviceSide
	{
		rt_context* ctx = nullptr;
		rt_host::FlatScene scratch;
		~DeviceSide() { if (ctx) rt_destroy(ctx); }
	};

	std::mutex g_mutex;
	std::unordered_map<const Renderer*, std::unique_ptr<DeviceSide>> g_devices;

	// `surface` / `surface_bytes`: the window surface's pixels (source/Renderer.cpp:27-29).  They live as long as the
	// window, i.e. longer than this context, so they are pinned once for direct device-to-host copies
	// (rt_register_surface's lifetime rule); if that fails rt_render goes through its bounce buffer.
	DeviceSide* DeviceFor(const Renderer* renderer, void* surface, size_t surface_bytes)
	{
		std::lock_guard<std::mutex> lock(g_mutex);
		auto it = g_devices.find(renderer);
		if (it != g_devices.end()) return it->second.get();

		std::vector<int32_t> ids;
		if (const char* env = std::getenv("RT_B200_DEVICES"))
		{
			const char* p = env;
			while (*p)
			{
				char* end = nullptr;
				const long v = std::strtol(p, &end, 10);
				if (end == p) break;
				ids.push_back((int32_t)v);
				p = (*end == ',') ? end + 1 : end;
			}
		}
		auto side = std::make_unique<DeviceSide>();
		const int rc = rt_create(ids.empty() ? nullptr : ids.data(), (int32_t)ids.size(), &side->ctx);
		if (rc != RT_OK)
		{
			std::fprintf(stderr, "rt_b200: rt_create failed (%d): %s\n", rc, rt_last_error(nullptr));
			side->ctx = nullptr;
		}
		else if (surface && surface_bytes && rt_register_surface(side->ctx, surface, surface_bytes) != RT_OK)
			std::fprintf(stderr, "rt_b200: the window surface could not be pinned (%s); frames go through a bounce buffer\n", rt_last_error(side->ctx));
		if (const char* env = std::getenv("RT_B200_DEVICE_TRANSFORM")) side->scratch.device_transform = std::max(0, std::min(2, std::atoi(env)));
		DeviceSide* raw = side.get();
		g_devices.emplace(renderer, std::move(side));
		return raw;
	}

	rt_camera CameraFor(Camera& camera)
	{
		const Matrix& m = camera.CalculateCameraToWorld();     // source/Renderer.cpp:40
		const Vector3 right = m.GetAxisX(), up = m.GetAxisY(), forward = m.GetAxisZ();
		rt_camera c{};
		c.origin[0] = camera.origin.x; c.origin[1] = camera.origin.y; c.origin[2] = camera.origin.z;
		c.fov = camera.fov;
		c.right[0] = right.x; c.right[1] = right.y; c.right[2] = right.z;
		c.up[0] = up.x; c.up[1] = up.y; c.up[2] = up.z;
		c.forward[0] = forward.x; c.forward[1] = forward.y; c.forward[2] = forward.z;
		return c;
	}
}

Renderer::Renderer(SDL_Window* pWindow) :
	m_pWindow(pWindow),
	m_pBuffer(SDL_GetWindowSurface(pWindow))
{
	SDL_GetWindowSize(pWindow, &m_Width, &m_Height);
	m_pBufferPixels = static_cast<uint32_t*>(m_pBuffer->pixels);
	m_AspectRatio = m_Width / static_cast<float>(m_Height);
}

void Renderer::Render(Scene* pScene) const
{
	DeviceSide* side = DeviceFor(this, m_pBufferPixels, (size_t)m_pBuffer->pitch * (size_t)m_Height);
	if (!side->ctx) return;

	std::string why;
	if (rt_host::UploadScene(side->ctx, pScene, side->scratch, why) != RT_OK)
	{
		std::fprintf(stderr, "rt_b200: scene upload failed: %s\n", why.c_str());
		return;
	}

	const rt_camera camera = CameraFor(pScene->GetCamera());
	rt_frame_desc frame{};
	frame.width = m_Width;
	frame.height = m_Height;
	frame.aspect_ratio = m_AspectRatio;
	frame.lighting_mode = static_cast<int32_t>(m_CurrentLightingMode);
	frame.shadows_enabled = m_ShadowsEnabled ? 1 : 0;
	frame.r_shift = m_pBuffer->format->Rshift;
	frame.g_shift = m_pBuffer->format->Gshift;
	frame.b_shift = m_pBuffer->format->Bshift;
	frame.alpha_mask = m_pBuffer->format->Amask;

	const int rc = rt_render(side->ctx, &camera, &frame, m_pBufferPixels, m_pBuffer->pitch);
	if (rc != RT_OK) std::fprintf(stderr, "rt_b200: rt_render failed (%d): %s\n", rc, rt_last_error(side->ctx));

	SDL_UpdateWindowSurface(m_pWindow);                    // source/Renderer.cpp:97
}

// Kept for source compatibility (source/Renderer.h:29).  One pixel through the same device path:
// the row that holds it is rendered on the GPU and the pixel copied into the surface.
void Renderer::RenderPixel(Scene* pScene, uint32_t pixelIndex, float, const Camera&, const std::vector<Light>&, const std::vector<Material*>&) const
{
	DeviceSide* side = DeviceFor(this, m_pBufferPixels, (size_t)m_pBuffer->pitch * (size_t)m_Height);
	if (!side->ctx) return;
	std::string why;
	if (rt_host::UploadScene(side->ctx, pScene, side->scratch, why) != RT_OK) return;
	const rt_camera camera = CameraFor(pScene->GetCamera());
	rt_frame_desc frame{};
	frame.width = m_Width; frame.height = m_Height; frame.aspect_ratio = m_AspectRatio;
	frame.lighting_mode = static_cast<int32_t>(m_CurrentLightingMode);
	frame.shadows_enabled = m_ShadowsEnabled ? 1 : 0;
	frame.r_shift = m_pBuffer->format->Rshift; frame.g_shift = m_pBuffer->format->Gshift; frame.b_shift = m_pBuffer->format->Bshift;
	frame.alpha_mask = m_pBuffer->format->Amask;
	std::vector<uint32_t> whole((size_t)m_Width * (size_t)m_Height);
	if (rt_render(side->ctx, &camera, &frame, whole.data(), m_Width * 4) != RT_OK) return;
	const int px = pixelIndex % m_Width, py = pixelIndex / m_Width;
	m_pBufferPixels[px + (py * m_Width)] = whole[pixelIndex];
}

bool Renderer::SaveBufferToImage() const
{
	return SDL_SaveBMP(m_pBuffer, "RayTracing_Buffer.bmp");
}

void Renderer::CycleLightingMode()
{
	const int count = static_cast<int>(LightingMode::Count);
	m_CurrentLightingMode = static_cast<LightingMode>((static_cast<int>(m_CurrentLightingMode) + 1) % count);
}
