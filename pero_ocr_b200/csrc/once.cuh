// cudaFuncSetAttribute applies to the current device only: a launch wrapper that raises a kernel's dynamic
// shared-memory limit must do so once per device, not once per process (one engine per device may live in one process).
#pragma once
#include <cuda_runtime.h>

struct PerDeviceOnce {
    bool done[64] = {};
    int device() const {
        int d = 0;
        return cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < 64 ? d : -1;
    }
    bool pending() const {
        const int d = device();
        return d < 0 || !done[d];
    }
    void mark() {
        const int d = device();
        if (d >= 0) done[d] = true;
    }
};
