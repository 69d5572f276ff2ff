/* stub: just enough for engine/core/sdl_wrapper.h to parse; nothing here is called by the code under test */
#pragma once
#include <stdint.h>
typedef uint32_t Uint32; typedef uint8_t Uint8; typedef uint16_t Uint16; typedef int32_t Sint32;
#define SDL_VERSION_ATLEAST(a, b, c) 1
typedef struct SDL_Surface SDL_Surface; typedef struct SDL_Window SDL_Window; typedef struct SDL_Renderer SDL_Renderer;
typedef struct SDL_Rect { int x, y, w, h; } SDL_Rect; typedef struct SDL_Cursor SDL_Cursor; typedef void* SDL_GLContext;
typedef int32_t SDL_Keycode; typedef struct SDL_KeyboardEvent { int type; } SDL_KeyboardEvent; typedef union SDL_Event { int type; } SDL_Event;
