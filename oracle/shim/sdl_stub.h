// TEST INFRASTRUCTURE (oracle build only) -- not part of the shipped product.
//
// Headless replacement for the handful of SDL 2.0.9 entry points the reference
// links against.  The vendored SDL headers under
// /root/reference/include/sdl2-2.0.9 supply the struct layouts; this file
// supplies a fake window that owns a W x H XRGB8888 surface in plain memory.
#pragma once
#include "SDL.h"

// SDL_Window is opaque in the public headers, so the stub is free to define it.
struct SDL_Window
{
	SDL_Surface surface;
	SDL_PixelFormat format;
	int width;
	int height;
};

extern "C" SDL_Window* GP1_CreateHeadlessWindow(int width, int height);
extern "C" void GP1_DestroyHeadlessWindow(SDL_Window* window);
