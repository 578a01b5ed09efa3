// TEST INFRASTRUCTURE (oracle build only) -- not part of the shipped product.
// See sdl_stub.h.  Call sites being satisfied:
//   Renderer.cpp:26,29,97,178,186  (surface, size, present, MapRGB, SaveBMP)
//   Camera.h:71,88                 (keyboard / mouse polling)
//   Timer.cpp:14,20,33 ...         (performance counter)
#include "sdl_stub.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

extern "C" {

SDL_Window* GP1_CreateHeadlessWindow(int width, int height)
{
	SDL_Window* w = static_cast<SDL_Window*>(calloc(1, sizeof(SDL_Window)));
	w->width = width;
	w->height = height;

	// SDL_PIXELFORMAT_RGB888 == XRGB8888: what SDL_GetWindowSurface hands back
	// on a desktop; alpha mask 0, so SDL_MapRGB yields 0x00RRGGBB.
	SDL_PixelFormat& f = w->format;
	f.format = SDL_PIXELFORMAT_RGB888;
	f.palette = nullptr;
	f.BitsPerPixel = 32;
	f.BytesPerPixel = 4;
	f.Rmask = 0x00FF0000u; f.Gmask = 0x0000FF00u; f.Bmask = 0x000000FFu; f.Amask = 0u;
	f.Rloss = 0; f.Gloss = 0; f.Bloss = 0; f.Aloss = 8;
	f.Rshift = 16; f.Gshift = 8; f.Bshift = 0; f.Ashift = 0;
	f.refcount = 1;
	f.next = nullptr;

	SDL_Surface& s = w->surface;
	s.flags = 0;
	s.format = &w->format;
	s.w = width;
	s.h = height;
	s.pitch = width * 4;
	s.pixels = calloc(static_cast<size_t>(width) * static_cast<size_t>(height), 4);
	s.refcount = 1;
	return w;
}

void GP1_DestroyHeadlessWindow(SDL_Window* window)
{
	if (!window) return;
	free(window->surface.pixels);
	free(window);
}

SDL_Surface* SDL_GetWindowSurface(SDL_Window* window) { return &window->surface; }

void SDL_GetWindowSize(SDL_Window* window, int* w, int* h)
{
	if (w) *w = window->width;
	if (h) *h = window->height;
}

int SDL_UpdateWindowSurface(SDL_Window*) { return 0; }

Uint32 SDL_MapRGB(const SDL_PixelFormat* format, Uint8 r, Uint8 g, Uint8 b)
{
	return (static_cast<Uint32>(r >> format->Rloss) << format->Rshift)
		| (static_cast<Uint32>(g >> format->Gloss) << format->Gshift)
		| (static_cast<Uint32>(b >> format->Bloss) << format->Bshift)
		| format->Amask;
}

SDL_RWops* SDL_RWFromFile(const char*, const char*) { return nullptr; }
int SDL_SaveBMP_RW(SDL_Surface*, SDL_RWops*, int) { return -1; }

static Uint8 g_keyboard[SDL_NUM_SCANCODES];
const Uint8* SDL_GetKeyboardState(int* numkeys)
{
	if (numkeys) *numkeys = SDL_NUM_SCANCODES;
	return g_keyboard;
}

Uint32 SDL_GetRelativeMouseState(int* x, int* y)
{
	if (x) *x = 0;
	if (y) *y = 0;
	return 0;
}

Uint64 SDL_GetPerformanceFrequency(void) { return 1000000000ull; }
Uint64 SDL_GetPerformanceCounter(void)
{
	using namespace std::chrono;
	return static_cast<Uint64>(duration_cast<nanoseconds>(steady_clock::now().time_since_epoch()).count());
}

} // extern "C"
