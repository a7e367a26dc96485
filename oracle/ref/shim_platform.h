// TEST INFRASTRUCTURE ONLY — never linked into the product library.
//
// Platform stand-ins (Win32 / OpenGL / FastNoiseSIMD / imgui types) so that the
// reference's own hot-path sources under /root/reference/Code compile VERBATIM with
// g++ (the reference is an MSVC/Win32 unity build, Code/Main.cpp:21-53).  Nothing here
// re-implements mesher behaviour; everything the mesher executes comes from the
// reference's files.  The only behavioural stand-in is FastNoiseSIMD (absent from the
// reference tree, Misc/BuildLib.bat:40-48) and it is only reachable from the terrain
// generators, which parity tests use solely for the flat world (no noise involved).
#pragma once

#include <cassert>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>

#include <algorithm>
#include <atomic>
#include <fstream>
#include <functional>
#include <list>
#include <queue>
#include <stack>
#include <string>
#include <unordered_map>
#include <vector>

#include "shim_glm.h"

using namespace glm;
using namespace std;

#define Unused(x) ((void)(x))
#define Print(...) fprintf(stderr, __VA_ARGS__)
#define Error(...) { fprintf(stderr, __VA_ARGS__); abort(); }

// ---- OpenGL: types + no-op calls (only Mesh.cpp's upload/draw code names them) ----
typedef unsigned int GLuint;
typedef int GLint;
typedef unsigned int GLenum;
typedef float GLfloat;
typedef int GLsizei;
typedef unsigned char GLboolean;
#define GL_FALSE 0
#define GL_TRUE 1
#define GL_ARRAY_BUFFER 0
#define GL_ELEMENT_ARRAY_BUFFER 0
#define GL_DYNAMIC_DRAW 0
#define GL_STATIC_DRAW 0
#define GL_UNSIGNED_BYTE 0
#define GL_UNSIGNED_SHORT 0
#define GL_FLOAT 0
#define GL_TRIANGLES 0
#define glGenVertexArrays(...) ((void)0)
#define glBindVertexArray(...) ((void)0)
#define glGenBuffers(...) ((void)0)
#define glBindBuffer(...) ((void)0)
#define glBufferData(...) ((void)0)
#define glBufferSubData(...) ((void)0)
#define glVertexAttribPointer(...) ((void)0)
#define glVertexAttribIPointer(...) ((void)0)
#define glEnableVertexAttribArray(...) ((void)0)
#define glVertexAttribDivisor(...) ((void)0)
#define glDeleteBuffers(...) ((void)0)
#define glDeleteVertexArrays(...) ((void)0)
#define glDrawElements(...) ((void)0)
#define glDrawElementsInstanced(...) ((void)0)
#define glUseProgram(...) ((void)0)
#define glUniform1f(...) ((void)0)
#define glUniform1i(...) ((void)0)
#define glUniform2f(...) ((void)0)
#define glUniform3f(...) ((void)0)
#define glUniform4f(...) ((void)0)
#define glUniformMatrix4fv(...) ((void)0)
#define glActiveTexture(...) ((void)0)
#define glBindTexture(...) ((void)0)
#define GL_TEXTURE0 0
#define GL_TEXTURE_2D 0
#define GL_TEXTURE_2D_ARRAY 0

// ---- Win32 synchronisation types: inert ----
struct CRITICAL_SECTION { int unused; };
struct CONDITION_VARIABLE { int unused; };
struct SRWLOCK { int unused; };
typedef void* HANDLE;
typedef unsigned long DWORD;
#define MAX_PATH 260
static inline void InitializeCriticalSection(CRITICAL_SECTION*) {}
static inline void EnterCriticalSection(CRITICAL_SECTION*) {}
static inline void LeaveCriticalSection(CRITICAL_SECTION*) {}
static inline void InitializeConditionVariable(CONDITION_VARIABLE*) {}
static inline void WakeConditionVariable(CONDITION_VARIABLE*) {}
static inline void SleepConditionVariableCS(CONDITION_VARIABLE*, CRITICAL_SECTION*, DWORD) {}
#define INFINITE 0xFFFFFFFF

// ---- Win32 file API: inert (WorldIO.cpp names it for region files; the harness only drives the in-memory RLE
// encode / decode of SaveGroup / LoadGroupFromDisk, never a file) ----
#define INVALID_HANDLE_VALUE ((HANDLE)(long long)-1)
#define GENERIC_READ 0
#define GENERIC_WRITE 0
#define OPEN_EXISTING 0
#define CREATE_ALWAYS 0
#define FILE_ATTRIBUTE_NORMAL 0
#define FILE_END 0
#define INVALID_SET_FILE_POINTER 0xFFFFFFFFul
static inline bool PathFileExists(const char*) { return false; }
static inline HANDLE CreateFile(const char*, int, int, void*, int, int, void*) { return INVALID_HANDLE_VALUE; }
static inline bool ReadFile(HANDLE, void*, DWORD, DWORD* n, void*) { if (n) *n = 0; return false; }
static inline bool WriteFile(HANDLE, const void*, DWORD, DWORD* n, void*) { if (n) *n = 0; return false; }
static inline void CloseHandle(HANDLE) {}
static inline DWORD SetFilePointer(HANDLE, long, long*, DWORD) { return 0; }
static inline bool CreateDirectory(const char*, void*) { return true; }
static inline std::string GetLastErrorText() { return "no file system in the test harness"; }
static inline void WriteBinary(char*, char*, int) {}
static inline void ReadBinary(char*, char*) {}

// ---- FastNoiseSIMD: API-shaped stand-in (deterministic hash noise, NOT the real library) ----
class FastNoiseSIMD
{
public:
    enum NoiseType { Value, ValueFractal, Perlin, PerlinFractal, Simplex, SimplexFractal, WhiteNoise, Cellular, Cubic, CubicFractal };
    enum FractalType { FBM, Billow, RigidMulti };

    static FastNoiseSIMD* NewFastNoiseSIMD(int seed = 1337) { FastNoiseSIMD* n = new FastNoiseSIMD; n->seed = seed; return n; }
    void SetSeed(int s) { seed = s; }
    void SetNoiseType(NoiseType t) { type = t; }
    void SetFrequency(float f) { freq = f; }
    void SetFractalOctaves(int o) { octaves = o; }
    void SetFractalType(FractalType t) { ftype = t; }
    void SetFractalGain(float) {}
    void SetFractalLacunarity(float) {}

    // Tests may plug a noise field in (gcref_set_noise): the reference's own generator code then runs over it.  The hook gets
    // what the generator configured (seed, noise type, fractal type, octaves), whether the set is a 2-D one (one layer) and the
    // sample position in FastNoiseSIMD's convention: (start + index) * frequency * scale per axis.
    typedef float (*Hook)(int seed, int type, int ftype, int octaves, int dims, float x, float y, float z);
    static Hook& hook() { static Hook h = nullptr; return h; }

    float* GetNoiseSet(int xs, int ys, int zs, int sx, int sy, int sz, float scale = 1.0f)
    {
        float* set = (float*)malloc(sizeof(float) * (size_t)sx * sy * sz);
        size_t i = 0;
        for (int x = 0; x < sx; x++)
            for (int y = 0; y < sy; y++)
                for (int z = 0; z < sz; z++)
                {
                    const float px = (xs + x) * freq * scale, py = (ys + y) * freq * scale, pz = (zs + z) * freq * scale;
                    set[i++] = hook() ? hook()(seed, (int)type, (int)ftype, octaves, sy == 1 ? 2 : 3, px, py, pz) : Smooth(px, py, pz);
                }
        return set;
    }
    static void FreeNoiseSet(float* s) { free(s); }

private:
    int seed = 1337, octaves = 3;
    float freq = 0.01f;
    NoiseType type = Simplex;
    FractalType ftype = FBM;

    float Lattice(int x, int y, int z) const
    {
        uint32_t h = (uint32_t)seed * 0x9E3779B1u ^ (uint32_t)x * 0x85EBCA6Bu ^ (uint32_t)y * 0xC2B2AE35u ^ (uint32_t)z * 0x27D4EB2Fu;
        h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
        return (h & 0xFFFFFF) / (float)0x7FFFFF - 1.0f;
    }
    float Smooth(float x, float y, float z) const
    {
        int xi = (int)floorf(x), yi = (int)floorf(y), zi = (int)floorf(z);
        float fx = x - xi, fy = y - yi, fz = z - zi;
        fx = fx * fx * (3 - 2 * fx); fy = fy * fy * (3 - 2 * fy); fz = fz * fz * (3 - 2 * fz);
        float r = 0;
        for (int dz = 0; dz < 2; dz++) for (int dy = 0; dy < 2; dy++) for (int dx = 0; dx < 2; dx++)
            r += Lattice(xi + dx, yi + dy, zi + dz) * (dx ? fx : 1 - fx) * (dy ? fy : 1 - fy) * (dz ? fz : 1 - fz);
        return r;
    }
};

// ---- imgui / misc forward types named by headers we keep ----
struct GameState;
struct Camera;
struct Renderer;
struct World;
struct Player;
