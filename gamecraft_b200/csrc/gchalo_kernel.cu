// gchalo — slab halo exchange over peer memory (see gchalo_kernel.cuh).
//
// Per exchange k, one kernel on every rank's own stream:
//   ready  tells both neighbours "I have entered exchange k" (whatever read my halo rows before is finished) and waits
//          until both have said the same — only then may their halo rows be overwritten
//   copy   boundary planes (blocks | sunlight | blockLight of every chunk of the row, plus the surface row) stored with
//          16-byte accesses straight into the neighbours' arrays; the last CTA to finish publishes "arrived k" in both
//          neighbours' flag blocks behind a system-scope fence
//   wait   until both neighbours' planes of exchange k have arrived here
// Every wait is bounded (HaloPushArgs::wait_ns): a missing neighbour raises HF_ERROR instead of hanging the stream.
#include "gchalo_kernel.cuh"

namespace gc
{
namespace
{
// One kernel per exchange.  No CTA waits for another CTA of the grid — only for flags the neighbours write — so the grid
// needs no co-residency: CTA 0 tells the neighbours this rank has entered the exchange, every CTA waits until both
// neighbours have said the same, copies its share, and leaves once the neighbours' planes have arrived here.
__global__ void __launch_bounds__(256) gc_halo_exchange_kernel(HaloPushArgs a)
{
    // ---- ready: I am the upper neighbour of the rank below me, the lower neighbour of the rank above ----
    if (blockIdx.x == 0 && threadIdx.x < 2 && a.peer[threadIdx.x].connected)
        halo_store_release_sys(a.peer[threadIdx.x].flags + (threadIdx.x == 0 ? HF_READY_FROM_UP : HF_READY_FROM_DOWN), a.serial);
    // a neighbour that does not enter the exchange in time is left alone: nothing is stored into its halo rows (its mesh
    // kernel may still be reading them), no arrival is published to it, and the failure is recorded in both flag blocks
    __shared__ int ready[2];
    if (threadIdx.x < 2)
        ready[threadIdx.x] = a.peer[threadIdx.x].connected &&
                             halo_wait_at_least(a.flags, threadIdx.x == 0 ? HF_READY_FROM_DOWN : HF_READY_FROM_UP, a.serial, a.wait_ns, a.peer[threadIdx.x].flags);
    __syncthreads();

    // ---- copy ----
    const int per_chunk = 8192 / 16;                      // uint4 per (column, cy): 256 blocks, 128 sunlight, 128 blockLight
    const int total = a.size_x * 4 * per_chunk;
    const int surf_words = a.size_x * 8;
    for (int d = 0; d < 2; d++)
    {
        const HaloPeer& p = a.peer[d];
        if (!ready[d]) continue;
        const int z = d == 0 ? 0 : 31;
        const int src_row = a.send_row[d];
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
        {
            const int plane = i / per_chunk, q = i % per_chunk;
            const int cx = plane >> 2, cy = plane & 3;
            const size_t src = (((size_t)src_row * a.size_x + cx) * 4 + cy) * GC_CHUNK_SIZE_3 + (size_t)z * 2048;
            const size_t dst = (((size_t)p.row * a.size_x + cx) * 4 + cy) * GC_CHUNK_SIZE_3 + (size_t)z * 2048;
            if (q < 256) reinterpret_cast<uint4*>(p.blocks + dst)[q] = reinterpret_cast<const uint4*>(a.blocks + src)[q];
            else if (q < 384) reinterpret_cast<uint4*>(p.sun + dst)[q - 256] = reinterpret_cast<const uint4*>(a.sun + src)[q - 256];
            else reinterpret_cast<uint4*>(p.light + dst)[q - 384] = reinterpret_cast<const uint4*>(a.light + src)[q - 384];
        }
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < surf_words; i += gridDim.x * blockDim.x)
        {
            const int cx = i >> 3, q = i & 7;
            reinterpret_cast<uint32_t*>(p.surface + ((size_t)p.row * a.size_x + cx) * 1024 + z * 32)[q] =
                reinterpret_cast<const uint32_t*>(a.surface + ((size_t)src_row * a.size_x + cx) * 1024 + z * 32)[q];
        }
    }
    // last CTA out publishes the arrival: every CTA's stores are fenced system-wide before its ticket
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(a.flags + HF_TICKET, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x < 2 && ready[threadIdx.x])
    {
        __threadfence_system();
        halo_store_release_sys(a.peer[threadIdx.x].flags + (threadIdx.x == 0 ? HF_ARRIVED_FROM_UP : HF_ARRIVED_FROM_DOWN), a.serial);
    }
    if (last && threadIdx.x == 0) a.flags[HF_TICKET] = 0;

    // ---- wait: the kernel (and with it the stream) goes on once both neighbours' planes are here ----
    if (threadIdx.x < 2 && ready[threadIdx.x])
        halo_wait_at_least(a.flags, threadIdx.x == 0 ? HF_ARRIVED_FROM_DOWN : HF_ARRIVED_FROM_UP, a.serial, a.wait_ns);
}
}  // namespace

// With lazy module loading the first launch of a kernel can synchronise the context — behind an exchange kernel that is
// spinning for a neighbour whose own launch is what got stuck (worlds of one process).  Loading it up front removes that.
cudaError_t preload_halo_kernels()
{
    cudaFuncAttributes attr;
    return cudaFuncGetAttributes(&attr, gc_halo_exchange_kernel);
}

cudaError_t launch_halo_push(const HaloPushArgs& a, int num_sms, cudaStream_t s)
{
    if (!a.peer[0].connected && !a.peer[1].connected) return cudaSuccess;
    const int total = a.size_x * 4 * 512;
    int grid = (total + 255) / 256;
    if (grid > num_sms) grid = num_sms;                    // one CTA per SM: 1 MB per direction is latency, not bandwidth
    gc_halo_exchange_kernel<<<grid, 256, 0, s>>>(a);
    return cudaGetLastError();
}
}  // namespace gc
