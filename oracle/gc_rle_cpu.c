/*
 * TEST INFRASTRUCTURE ONLY.  CPU oracle for the reference's region chunk format — the run-length coding of a chunk's block
 * ids that SaveGroup writes into region->chunks[] and LoadGroupFromDisk reads back (Code/WorldIO.cpp:245-307, 167-221).
 *
 *   record = [offset][items][(count, block) x items/2]        all uint16
 *   offset   RegionIndex of the chunk inside its 8 x 4 x 8 region (WorldIO.cpp:5-8)
 *   items    number of uint16 that follow, stored through a uint16 cast (65 536 single-voxel runs would store 0)
 *   count    run length in voxel index order (x fastest, then y, then z), a uint16 that wraps: a chunk filled with one
 *            block is the single pair (0, block) — 65 536 wrapped — which the loader special-cases (WorldIO.cpp:200-205)
 *
 * Parity status: PINNED against oracle/_ref/libgcref.so (gcref_rle_save_group / gcref_rle_load_group run the reference's
 * own SaveGroup / LoadGroupFromDisk), tests/test_rle_oracle.py.
 */
#include <stdint.h>

#define CVOX 65536

/* SaveGroup's inner loop (WorldIO.cpp:272-303).  out needs room for 2 + 2 * 65 536 words.  Returns the words written. */
int gcor_rle_encode_chunk(const uint16_t* blocks, uint16_t offset, uint16_t* out)
{
    int n = 0;
    out[n++] = offset;
    out[n++] = 0;
    uint16_t current = blocks[0];
    uint16_t count = 1;
    for (int i = 1; i < CVOX; i++)
    {
        const uint16_t b = blocks[i];
        if (b != current)
        {
            out[n++] = count; out[n++] = current;
            count = 1; current = b;
        }
        else count++;                               /* uint16: wraps to 0 after 65 536 equal voxels */
        if (i == CVOX - 1) { out[n++] = count; out[n++] = current; }
    }
    out[1] = (uint16_t)(n - 2);
    return n;
}

/*
 * LoadGroupFromDisk's inner loop (WorldIO.cpp:196-216) on one record of `words` uint16.  Returns 0, or -1 when the
 * runs do not add up to exactly one chunk (the reference would write past the array; the oracle stops instead).
 */
int gcor_rle_decode_chunk(const uint16_t* record, int words, uint16_t* blocks)
{
    if (words < 4) return -1;
    int loc = 2, i = 0;
    if (record[loc] == 0)
    {
        for (int k = 0; k < CVOX; k++) blocks[k] = record[loc + 1];   /* FillChunk */
        return 0;
    }
    while (i < CVOX)
    {
        if (loc + 1 >= words) return -1;
        const int count = record[loc++];
        const uint16_t b = record[loc++];
        if (count == 0 || i + count > CVOX) return -1;
        for (int j = 0; j < count; j++) blocks[i++] = b;
    }
    return 0;
}
