// Tile hand-out shared by the persistent tcgen05 kernels (igemm_tc.cu, igemm_halo.cu).
//
// Static striding (tile = blockIdx.x, += gridDim.x) gives every CTA a fixed 1/gridDim share of the tiles: a CTA whose
// SM is still busy with another stream's kernel when the launch begins (the BiLSTM recurrence of the previous batch
// holds 64 SMs for ~1 ms) starts late and finishes its whole share late -- the kernel then lasts "other kernel + own
// time".  With IgemmParams::tile_counter the producer warp draws tile indices from a global counter instead and hands
// them to the other warp roles of its CTA through a 4-deep shared-memory ring (mbarrier full / empty pairs), so late
// CTAs simply draw fewer tiles and the launch is work-conserving.  Without a counter the same ring carries the static
// sequence (one code path).
#pragma once
#include "ptx.cuh"

constexpr int kTileRing = 4;

// Producer side: every lane of the (converged) producer warp calls next(); lane 0 draws and publishes.
struct TileFeed {
    int* counter;
    int total;
    uint64_t* full;
    uint64_t* empty;
    int* slot;
    int idx = 0;
    int static_next;
    __device__ TileFeed(int* counter_, int total_, uint64_t* full_, uint64_t* empty_, int* slot_)
        : counter(counter_), total(total_), full(full_), empty(empty_), slot(slot_), static_next(blockIdx.x) {}
    __device__ __forceinline__ int next(int lane) {
        int t = 0;
        if (counter) {
            if (lane == 0) t = atomicAdd(counter, 1);
            t = __shfl_sync(0xffffffffu, t, 0);
        } else {
            t = static_next;
            static_next += gridDim.x;
        }
        if (t > total) t = total;   // one terminator value
        const int s = idx & (kTileRing - 1);
        ptx::mbar_wait(&empty[s], ((idx / kTileRing) & 1) ^ 1);
        if (lane == 0) {
            *reinterpret_cast<volatile int*>(&slot[s]) = t;
            ptx::mbar_arrive(&full[s]);      // release: the slot write is visible to whoever acquires the barrier
        }
        __syncwarp();
        ++idx;
        return t;
    }
};

// Consumer side: each consumer (a single thread, or a warp with lane 0 arriving for it) sees every tile once.
struct TileTake {
    int total;
    uint64_t* full;
    uint64_t* empty;
    int* slot;
    int idx = 0;
    __device__ TileTake(int* /*counter*/, int total_, uint64_t* full_, uint64_t* empty_, int* slot_)
        : total(total_), full(full_), empty(empty_), slot(slot_) {}
    __device__ __forceinline__ int next() {                 // one thread
        const int s = idx & (kTileRing - 1);
        ptx::mbar_wait(&full[s], (idx / kTileRing) & 1);
        const int t = *reinterpret_cast<volatile int*>(&slot[s]);
        ptx::mbar_arrive(&empty[s]);
        ++idx;
        return t;
    }
    __device__ __forceinline__ int next_warp(int lane) {    // a converged warp
        const int s = idx & (kTileRing - 1);
        ptx::mbar_wait(&full[s], (idx / kTileRing) & 1);
        const int t = *reinterpret_cast<volatile int*>(&slot[s]);
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[s]);
        ++idx;
        return t;
    }
};
