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

// ids >= n_types in the block arena: how many, and the first one's voxel index (the reference asserts block < BLOCK_COUNT
// when a block is stored, World.cpp:294; the mesh kernel reads such ids through an 8-bit table index)
__global__ void __launch_bounds__(256) gc_validate_blocks_kernel(const uint4* __restrict__ blocks, size_t n_vec, uint32_t n_types, unsigned long long* out)
{
    unsigned long long bad = 0, first = ~0ull;
    const uint32_t lim2 = n_types | (n_types << 16);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x)
    {
        const uint4 v = blocks[i];
        // per 16-bit half: id >= n_types
        const uint32_t m = __vcmpgeu2(v.x, lim2) | __vcmpgeu2(v.y, lim2) | __vcmpgeu2(v.z, lim2) | __vcmpgeu2(v.w, lim2);
        if (m)
        {
            const uint32_t w[4] = { v.x, v.y, v.z, v.w };
            for (int k = 0; k < 8; k++)
                if (((w[k >> 1] >> (16 * (k & 1))) & 0xFFFFu) >= n_types)
                {
                    bad++;
                    const unsigned long long at = (unsigned long long)i * 8 + k;
                    if (at < first) first = at;
                }
        }
    }
    for (int d = 16; d > 0; d >>= 1)
    {
        bad += __shfl_xor_sync(0xffffffffu, bad, d);
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, first, d);
        first = o < first ? o : first;
    }
    if ((threadIdx.x & 31) == 0 && bad)
    {
        atomicAdd(out, bad);
        atomicMin(out + 1, first);
    }
}

// zero fill as a kernel: cudaMemsetAsync may be served by a copy engine, where it queues behind the transfers that keep the
// engines busy during gcmesh_mesh_host
__global__ void __launch_bounds__(256) gc_zero_kernel(uint4* __restrict__ p, size_t n_vec)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(0u, 0u, 0u, 0u);
}
}  // namespace

cudaError_t launch_zero(void* p, size_t bytes, int num_sms, cudaStream_t s)
{
    if (bytes == 0) return cudaSuccess;
    if ((((uintptr_t)p | bytes) & 15u) != 0) return cudaMemsetAsync(p, 0, bytes, s);
    const size_t n_vec = bytes / 16;
    size_t grid = (n_vec + 255) / 256;
    if (grid > (size_t)num_sms * 8) grid = (size_t)num_sms * 8;
    gc_zero_kernel<<<(unsigned)grid, 256, 0, s>>>(reinterpret_cast<uint4*>(p), n_vec);
    return cudaGetLastError();
}

cudaError_t launch_validate_blocks(const uint16_t* blocks, size_t n_voxels, int n_types, unsigned long long* d_out, int num_sms, cudaStream_t s)
{
    gc_validate_blocks_kernel<<<num_sms * 8, 256, 0, s>>>(reinterpret_cast<const uint4*>(blocks), n_voxels / 8, (uint32_t)n_types, d_out);
    return cudaGetLastError();
}

cudaError_t launch_fill_device(const FillJob* jobs, int n, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    gc_fill_device_kernel<<<n, 256, 0, s>>>(jobs);
    return cudaGetLastError();
}
}  // namespace gc
