// Constants and kernel parameters of the tcgen05 / TMEM prefilter (scan_tc5.cuh); shared with the host code in api.cu,
// which builds the weight images and column tables.
#pragma once
#include "common.cuh"

namespace smg {

constexpr int kTc5EpiWarps = 16;
constexpr int kTc5Threads = 32 * (2 + kTc5EpiWarps);
constexpr int kTc5ProdWarp = kTc5EpiWarps;              // strip producer
constexpr int kTc5MmaWarp = kTc5EpiWarps + 1;           // MMA issuer (highest warp id: wins the SMSP arbiter)
constexpr int kTc5Tile = 512;                           // window starts per tile (2 accumulators of 256: even / odd)
constexpr int kTc5StripBases = 544;                     // 2 * 255 + 1 + 32 (widest PWM) rounded up to whole words
constexpr int kTc5StripBytes = kTc5StripBases * 8;      // 4 fp16 per base
constexpr int kTc5EntryBytes = 80;                      // 16 accumulator cells + {threshold, column id} of one flagged lane
constexpr int kTc5ScratchBytes = 32 * kTc5EntryBytes;   // per epilogue warp: the flagged lanes of one group of 16 cells
constexpr int kTc5CandBlock = 128;                      // candidate slots an epilogue warp reserves per atomic
constexpr int kTc5MaxBlk = 40;                          // 128-column blocks per column group
// shared-memory carve-up (bytes)
constexpr int kTc5OffStrip = 0;
constexpr int kTc5OffScratch = kTc5OffStrip + 4 * kTc5StripBytes;
constexpr int kTc5OffMisc = kTc5OffScratch + kTc5EpiWarps * kTc5ScratchBytes;
constexpr int kTc5OffW = ((kTc5OffMisc + 1024 + 1023) / 1024) * 1024;  // weight image of the column group
constexpr int kTc5MaxImageBytes = 232448 - kTc5OffW;                  // 227 KB per CTA
constexpr int kTc5UnitBytes = 4096;                                   // 128 columns x one k16 step of fp16
#define SMG_CAND_INVALID 0xffffffffffffffffull                        // padding of a reserved candidate block

struct Tc5Params {
    ScanParams sp;
    const uint8_t* wimg;          // shared-memory images of all column groups
    const int64_t* cg_img_off;    // [n_colgroups] byte offset of a group's image in wimg
    const int32_t* cg_first_blk;  // [n_colgroups + 1] first 128-column block of each group
    const int32_t* blk_ks;        // [n_blocks] k16 steps of a block
    const int32_t* blk_off;       // [n_blocks] byte offset of a block inside its group's image
    const float* thr_pre_col;     // [n_blocks * 128] prefilter threshold per column (+inf: empty column / strand off)
    const int32_t* col_id;        // [n_blocks * 128] motif * 2 + strand, -1 for empty columns
    int32_t n_chunks;
    int32_t cand_shift;           // candidate = row << cand_shift | start << (motif_bits + 1) | motif << 1 | strand
    unsigned long long* cand;
    unsigned long long cand_cap;
};

}  // namespace smg
