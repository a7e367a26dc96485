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
    HF_ERROR = 4,               // a wait gave up (wait_ns) — here or, reported by a neighbour, there: (serial << 8) | (1 + flag word waited for)
    HF_TICKET = 5,              // CTA counter of the copy kernel; in the fused mesh + exchange kernel: push tasks done towards the rank below
    HF_TICKET_UP = 6,           // fused kernel: push tasks done towards the rank above
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

// The exchange fused into the mesh kernel (gcmesh_submit_exchange): the kernel's first work items store this rank's boundary
// planes into the neighbours' halo rows; chunks of the rows that read a halo plane wait for the neighbour's "arrived" flag.
struct HaloFused
{
    int32_t enabled;            // 0: plain submission
    int32_t tasks_per_dir;      // push work items per connected direction
    const uint16_t* blocks; const uint8_t* sun; const uint8_t* light; const uint8_t* surface;   // this rank's window
    int32_t size_x;
    int32_t send_row[2];        // as HaloPushArgs; chunks of row send_row[d] read the plane that peer d sends
    HaloPeer peer[2];
    uint32_t* flags;            // this world's flag block
    uint32_t serial;
    unsigned long long wait_ns;
    int32_t debug;              // GCMESH_HALO_DEBUG (development): 1 no copies, 2 no arrival waits, 4 no ready handshake
};

#if defined(__CUDACC__)
__device__ __forceinline__ void halo_store_release_sys(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t halo_load_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long halo_now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// A wait that gives up records which one it was, here and (when `peer_flags` is given) in the neighbour's flag block:
// HF_ERROR = (serial << 8) | (flag word + 1), first failure only.
__device__ __forceinline__ bool halo_wait_at_least(uint32_t* flags, int word, uint32_t serial, unsigned long long limit_ns, uint32_t* peer_flags = nullptr)
{
    if ((int32_t)(halo_load_acquire_sys(flags + word) - serial) >= 0) return true;
    const unsigned long long t0 = halo_now_ns();
    while ((int32_t)(halo_load_acquire_sys(flags + word) - serial) < 0)
    {
        if (halo_now_ns() - t0 > limit_ns)
        {
            const uint32_t code = (serial << 8) | (uint32_t)(word + 1);
            atomicCAS(flags + HF_ERROR, 0u, code);
            if (peer_flags) atomicCAS_system(peer_flags + HF_ERROR, 0u, code);
            return false;
        }
        __nanosleep(200);
    }
    return true;
}
#endif

// loads the kernel (call before the first exchange; see gchalo_kernel.cu)
cudaError_t preload_halo_kernels();
// one exchange (ready -> copy -> wait inside one kernel), enqueued on `s`
cudaError_t launch_halo_push(const HaloPushArgs& a, int num_sms, cudaStream_t s);
}  // namespace gc
