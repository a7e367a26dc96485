// gcfill — batch arenas -> per-chunk device buffers (see gcfill_kernel.cuh).
#include "gcfill_kernel.cuh"

namespace gc
{
namespace
{
// segments start on 4-byte boundaries in the arenas (48 B per quad of vertices, 12 B per quad of indices) and hold whole
// quads; 16-byte accesses where both sides allow it (mapped graphics buffers are at least 256-byte aligned)
__device__ void copy_bytes(void* dst, const void* src, size_t bytes)
{
    if ((((uintptr_t)dst | (uintptr_t)src | bytes) & 15u) == 0)
    {
        const size_t n = bytes / 16;
        for (size_t i = threadIdx.x; i < n; i += blockDim.x) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
    }
    else if ((((uintptr_t)dst | (uintptr_t)src | bytes) & 3u) == 0)
    {
        const size_t n = bytes / 4;
        for (size_t i = threadIdx.x; i < n; i += blockDim.x) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
    }
    else
        for (size_t i = threadIdx.x; i < bytes; i += blockDim.x) reinterpret_cast<uint8_t*>(dst)[i] = reinterpret_cast<const uint8_t*>(src)[i];
}

__global__ void __launch_bounds__(256) gc_fill_device_kernel(const FillJob* __restrict__ jobs)
{
    const FillJob j = jobs[blockIdx.x];
    for (int k = 0; k < 1 + GC_MESH_TYPE_COUNT; k++)
        if (j.bytes[k]) copy_bytes(j.dst[k], j.src[k], j.bytes[k]);
}
}  // namespace

cudaError_t launch_fill_device(const FillJob* jobs, int n, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    gc_fill_device_kernel<<<n, 256, 0, s>>>(jobs);
    return cudaGetLastError();
}
}  // namespace gc
