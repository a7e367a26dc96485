// gchalo kernels (sm_100a): the slab halo exchange of the multi-GPU path as direct peer-memory stores over NVLink —
// one copy kernel writes this rank's boundary planes straight into the neighbours' window arrays, ordered by
// system-scope flags in the neighbours' memory.  No staging buffer, no NCCL call, no host round trip.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gcmesh.h"

namespace gc
{
// words of a world's flag block (device memory, written by the neighbours)
enum
{
    HF_ARRIVED_FROM_DOWN = 0,   // serial of the last exchange whose planes from the rank below have landed
    HF_ARRIVED_FROM_UP = 1,
    HF_READY_FROM_DOWN = 2,     // serial of the last exchange the rank below has entered: everything it had enqueued before
    HF_READY_FROM_UP = 3,       //   (the mesh kernel reading ITS halo) is done, so its halo rows may be overwritten
    HF_ERROR = 4,               // a wait gave up (wait_ns)
    HF_TICKET = 5,              // CTA counter of the copy kernel
    HF_WORDS = 8
};

struct HaloPeer
{
    uint16_t* blocks; uint8_t* sun; uint8_t* light; uint8_t* surface;   // the neighbour's window arrays (peer mapped)
    uint32_t* flags;                                                    // the neighbour's flag block
    int32_t row;                                                        // its halo row that receives the plane
    int32_t connected;
};

struct HaloPushArgs
{
    const uint16_t* blocks; const uint8_t* sun; const uint8_t* light; const uint8_t* surface;   // this rank's window
    uint32_t* flags;
    int32_t size_x;
    int32_t send_row[2];        // [0]: first owned row (its z = 0 plane goes down), [1]: last owned row (z = 31 plane goes up)
    HaloPeer peer[2];           // [0]: the rank below (z - 1), [1]: the rank above
    uint32_t serial;            // 1, 2, ... per exchange
    unsigned long long wait_ns; // how long a wait for a neighbour may last before it gives up
};

// loads the kernel (call before the first exchange; see gchalo_kernel.cu)
cudaError_t preload_halo_kernels();
// one exchange (ready -> copy -> wait inside one kernel), enqueued on `s`
cudaError_t launch_halo_push(const HaloPushArgs& a, int num_sms, cudaStream_t s);
}  // namespace gc
