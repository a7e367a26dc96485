// gcrle — the region chunk format on the device (see gcrle_kernel.cuh).
#include "gcrle_kernel.cuh"

namespace gc
{
namespace
{
constexpr int kVox = GC_CHUNK_SIZE_3;
constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}

// CTA-wide inclusive scan of one value per thread (256 threads); returns the inclusive prefix, *total = sum of all
__device__ uint32_t cta_incl_scan(uint32_t v, uint32_t* s_warp, uint32_t* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t incl = warp_incl_scan(v, lane);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t base = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++)
    {
        const uint32_t t = s_warp[w];
        if (w < warp) base += t;
    }
    uint32_t all = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) all += s_warp[w];
    *total = all;
    __syncthreads();
    return base + incl;
}

// ---- decode: one CTA per record (LoadGroupFromDisk's inner loop, WorldIO.cpp:196-216) ----
// 1. inclusive scan of the run lengths, 256 runs per step, into run_ends   2. every thread owns 256 consecutive voxels:
// one binary search for the run its first voxel falls in, then a walk along the runs, eight voxels per 16-byte store
__global__ void __launch_bounds__(kThreads) gc_rle_decode_kernel(const RleJob* __restrict__ jobs, const uint16_t* __restrict__ data,
                                                                 uint32_t* __restrict__ run_ends, uint16_t* __restrict__ blocks, uint32_t* bad_records)
{
    __shared__ uint32_t s_warp[kThreads / 32];
    __shared__ uint32_t s_carry;
    const RleJob job = jobs[blockIdx.x];
    const int tid = threadIdx.x;
    uint4* out = reinterpret_cast<uint4*>(blocks + (size_t)job.slot * kVox);
    const uint16_t* rec = data + job.first_word;
    const uint32_t n_runs = job.words >= 4 ? (job.words - 2) / 2 : 0;
    const uint32_t* pairs = reinterpret_cast<const uint32_t*>(rec + 2);   // (count | block << 16); records start on even words
    bool ok = n_runs > 0;
    if (ok && (pairs[0] & 0xFFFFu) == 0)
    {
        // 65 536 equal voxels, the count wrapped to 0: FillChunk (WorldIO.cpp:200-205)
        const uint32_t b = pairs[0] >> 16, bb = b | (b << 16);
        for (int i = tid; i < kVox / 8; i += kThreads) out[i] = make_uint4(bb, bb, bb, bb);
        return;
    }
    uint32_t* ends = run_ends + job.pair_base;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t r0 = 0; r0 < n_runs && ok; r0 += kThreads)
    {
        const uint32_t r = r0 + tid;
        const uint32_t c = r < n_runs ? (pairs[r] & 0xFFFFu) : 0u;
        uint32_t total;
        const uint32_t incl = cta_incl_scan(c, s_warp, &total);
        const uint32_t carry = s_carry;
        if (r < n_runs) ends[r] = carry + incl;
        const bool zero_run = r < n_runs && c == 0;
        ok = !__syncthreads_or(zero_run) && carry + total <= (uint32_t)kVox;
        if (tid == 0) s_carry = carry + total;
        __syncthreads();
    }
    ok = ok && s_carry == (uint32_t)kVox;
    if (!ok)
    {
        // the runs do not add up to one chunk: refuse the record (the reference would write past the array)
        for (int i = tid; i < kVox / 8; i += kThreads) out[i] = make_uint4(0u, 0u, 0u, 0u);
        if (tid == 0) atomicAdd(bad_records, 1u);
        return;
    }
    __threadfence_block();
    const uint32_t v0 = (uint32_t)tid * (kVox / kThreads);
    uint32_t lo = 0, hi = n_runs - 1;                      // first run whose end lies beyond v0
    while (lo < hi)
    {
        const uint32_t mid = (lo + hi) >> 1;
        if (ends[mid] > v0) hi = mid; else lo = mid + 1;
    }
    uint32_t run = lo, end = ends[run], b = pairs[run] >> 16;
#pragma unroll 1
    for (int g = 0; g < kVox / kThreads / 8; g++)
    {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 8; k++)
        {
            const uint32_t v = v0 + (uint32_t)g * 8 + k;
            if (v >= end) { run++; end = ends[run]; b = pairs[run] >> 16; }   // (runs are non-empty: one step always suffices)
            if (k & 1) w[k >> 1] |= b << 16; else w[k >> 1] = b;
        }
        out[(v0 >> 3) + g] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ---- encode: one CTA per brick (SaveGroup's inner loop, WorldIO.cpp:272-303) ----
// A voxel starts a run when it differs from its predecessor.  Per tile of 2048 voxels: flags, CTA scan -> run index;
// every run head writes its block id and (through its index) lets its predecessor compute its length.
__global__ void __launch_bounds__(kThreads) gc_rle_encode_kernel(const uint32_t* __restrict__ slots, const uint16_t* __restrict__ offsets,
                                                                 const uint16_t* __restrict__ blocks, uint16_t* __restrict__ out_all, uint32_t* __restrict__ out_words)
{
    __shared__ uint32_t s_warp[kThreads / 32];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x;
    const uint16_t* src = blocks + (size_t)slots[blockIdx.x] * kVox;
    uint16_t* out = out_all + (size_t)blockIdx.x * kRleMaxWords;
    // pass 1: run heads -> starts[run] kept in the output area itself (as uint32 pairs are written later, the starts go to a
    // side array at the end of the record area: word offset 2 + 2*run is final, so starts are staged in registers per tile)
    if (tid == 0) s_carry = 0;
    __syncthreads();
    constexpr int kPer = 8;                                  // voxels per thread per tile
    for (int base = 0; base < kVox; base += kThreads * kPer)
    {
        const int v0 = base + tid * kPer;
        const uint4 q = *reinterpret_cast<const uint4*>(src + v0);
        const uint32_t wds[4] = { q.x, q.y, q.z, q.w };
        const uint32_t prev = v0 > 0 ? src[v0 - 1] : 0xFFFFFFFFu;
        uint32_t id[kPer], head = 0;
#pragma unroll
        for (int k = 0; k < kPer; k++)
        {
            id[k] = (wds[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
            const uint32_t p = k ? id[k - 1] : prev;
            if (id[k] != p) head |= 1u << k;
        }
        uint32_t total;
        const uint32_t incl = cta_incl_scan((uint32_t)__popc(head), s_warp, &total);
        const uint32_t carry = s_carry;
        uint32_t run = carry + incl - (uint32_t)__popc(head);   // index of this thread's first run head
#pragma unroll
        for (int k = 0; k < kPer; k++)
        {
            if ((head >> k) & 1u)
            {
                // run `run` starts at voxel v0 + k with block id[k]: its block word is final; its start is parked in the
                // count word until the lengths are taken (two 16-bit halves: start < 65536 fits)
                out[2 + 2 * run + 1] = (uint16_t)id[k];
                out[2 + 2 * run] = (uint16_t)(v0 + k);
                run++;
            }
        }
        if (tid == 0) s_carry = carry + total;
        __syncthreads();
    }
    const uint32_t n_runs = s_carry;
    __syncthreads();
    // pass 2: count = next start - own start (the last run ends at 65 536); a uint16, so a one-run chunk stores 0
    for (uint32_t r0 = 0; r0 < n_runs; r0 += kThreads)
    {
        const uint32_t r = r0 + tid;
        uint32_t start = 0, next = 0;
        if (r < n_runs)
        {
            start = out[2 + 2 * r];
            next = r + 1 < n_runs ? out[2 + 2 * (r + 1)] : (uint32_t)kVox;
        }
        __syncthreads();                                      // every start of the tile (and the next tile's first) is read before any is overwritten
        if (r < n_runs) out[2 + 2 * r] = (uint16_t)(next - start);
        __syncthreads();
    }
    if (tid == 0)
    {
        out[0] = offsets[blockIdx.x];
        out[1] = (uint16_t)(2 * n_runs);                       // items, through the reference's uint16 cast
        out_words[blockIdx.x] = 2 + 2 * n_runs;
    }
}
// ---- pack a group's records back to back: record i goes to dense + sum(words[0 .. i)) ----
__global__ void __launch_bounds__(kThreads) gc_rle_pack_kernel(const uint16_t* __restrict__ out_all, const uint32_t* __restrict__ words, int n,
                                                               uint16_t* __restrict__ dense, uint32_t* __restrict__ total)
{
    __shared__ uint32_t s_warp[kThreads / 32];
    const int tid = threadIdx.x, i = blockIdx.x;
    uint32_t before = 0, all = 0;
    for (int j = tid; j < n; j += kThreads)
    {
        const uint32_t wj = words[j];
        all += wj;
        if (j < i) before += wj;
    }
    uint32_t sum_all;
    const uint32_t incl = cta_incl_scan(before, s_warp, &sum_all);   // sum_all = words before record i
    (void)incl;
    if (i == 0)
    {
        uint32_t t;
        cta_incl_scan(all, s_warp, &t);
        if (tid == 0) *total = t;
    }
    const uint32_t* src = reinterpret_cast<const uint32_t*>(out_all + (size_t)i * kRleMaxWords);   // records are even-sized, even-placed
    uint32_t* dst = reinterpret_cast<uint32_t*>(dense + sum_all);
    const uint32_t pairs = words[i] / 2;
    for (uint32_t k = tid; k < pairs; k += kThreads) dst[k] = src[k];
}
}  // namespace

cudaError_t launch_rle_pack(const uint16_t* out_all, const uint32_t* words, int n, uint16_t* dense, uint32_t* total, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    gc_rle_pack_kernel<<<n, kThreads, 0, s>>>(out_all, words, n, dense, total);
    return cudaGetLastError();
}

cudaError_t launch_rle_decode(const RleJob* jobs, int n, const uint16_t* data, uint32_t* run_ends, uint16_t* blocks, uint32_t* bad_records, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    gc_rle_decode_kernel<<<n, kThreads, 0, s>>>(jobs, data, run_ends, blocks, bad_records);
    return cudaGetLastError();
}

cudaError_t launch_rle_encode(const uint32_t* slots, const uint16_t* offsets, int n, const uint16_t* blocks, uint16_t* out, uint32_t* out_words, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    gc_rle_encode_kernel<<<n, kThreads, 0, s>>>(slots, offsets, blocks, out, out_words);
    return cudaGetLastError();
}
}  // namespace gc
