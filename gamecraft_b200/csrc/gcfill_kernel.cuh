// gcfill kernel (sm_100a): the hand-off of finished meshes to buffers that live on the device — what FillMeshData
// (Code/Mesh.cpp:71-130) does with glBufferData from host memory, for graphics buffers mapped into CUDA
// (cudaGraphicsGLRegisterBuffer / cudaGraphicsResourceGetMappedPointer): one vertex buffer and one element buffer per mesh
// type per chunk, filled straight from the batch arenas without a trip through the host.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gcmesh.h"

namespace gc
{
// one chunk's streams, resolved on the host from its descriptor: where each lies in the arenas, where it goes, how long it is
struct FillJob
{
    const void* src[1 + GC_MESH_TYPE_COUNT];   // [0] vertices, [1 + t] index stream t
    void* dst[1 + GC_MESH_TYPE_COUNT];
    uint32_t bytes[1 + GC_MESH_TYPE_COUNT];    // 0 = nothing to copy
};
// one CTA per job
cudaError_t launch_fill_device(const FillJob* jobs, int n, cudaStream_t s);
// zero fill by a kernel (16-byte stores; falls back to cudaMemsetAsync for unaligned ranges)
cudaError_t launch_zero(void* p, size_t bytes, int num_sms, cudaStream_t s);
// block ids >= n_types: d_out[0] += their number, d_out[1] = min(d_out[1], voxel index of the first); the caller presets {0, ~0}
cudaError_t launch_validate_blocks(const uint16_t* blocks, size_t n_voxels, int n_types, unsigned long long* d_out, int num_sms, cudaStream_t s);
}  // namespace gc
