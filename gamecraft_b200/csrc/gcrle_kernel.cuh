// gcrle kernels (sm_100a): the reference's region chunk format — the run-length coding of a chunk's block ids that
// SaveGroup writes and LoadGroupFromDisk reads (Code/WorldIO.cpp:245-307, 167-221) — decoded / encoded on the device,
// so that saved chunks cross PCIe as their few-KB records instead of 128 KB bricks.
//   record = [offset][items][(count, block) x n]   all uint16; count wraps: (0, b) = the whole chunk is block b
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gcmesh.h"

namespace gc
{
struct RleJob
{
    uint32_t slot;        // destination brick
    uint32_t first_word;  // the record's position in the uint16 stream (its [offset] word)
    uint32_t words;       // record length in uint16, header included
    uint32_t pair_base;   // position of the record's first run in the scratch array of run ends
};

// Decode `n` records of `data` (device) into their bricks of `blocks`; run_ends: scratch, one uint32 per (count, block)
// pair of the stream; bad_records: incremented per record whose runs do not add up to one chunk (its brick is zeroed).
cudaError_t launch_rle_decode(const RleJob* jobs, int n, const uint16_t* data, uint32_t* run_ends, uint16_t* blocks, uint32_t* bad_records, cudaStream_t s);

// Encode `n` bricks into records.  Per chunk: out_words[i] = record length; the record itself goes to
// out + i * kRleMaxWords (worst case: every voxel its own run).  offsets[i] = the record's [offset] word.
constexpr int kRleMaxWords = 2 + 2 * GC_CHUNK_SIZE_3;
cudaError_t launch_rle_encode(const uint32_t* slots, const uint16_t* offsets, int n, const uint16_t* blocks, uint16_t* out, uint32_t* out_words, cudaStream_t s);
// Records of one encoded group packed back to back into `dense`; *total = their words.  words[n] as written by the encoder.
cudaError_t launch_rle_pack(const uint16_t* out_all, const uint32_t* words, int n, uint16_t* dense, uint32_t* total, cudaStream_t s);
}  // namespace gc
