// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, fp16) as a function of N, operand source and accumulator
// reuse.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pero_ocr_b200/csrc tools/mma_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"

__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(lbo >> 4) << 16;
    d |= static_cast<uint64_t>(sbo >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    return d;
}

// mode 0: SS sw128 A/B, one accumulator; 1: SS, two alternating accumulators; 2: TS (A in TMEM), one accumulator;
// 3: TS, alternating accumulators; 4: SS with no-swizzle B
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int iters) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar, 1);
        ptx::fence_mbar_init();
    }
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
    if (threadIdx.x < 32) ptx::tmem_alloc<512>(&slot);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (threadIdx.x < 32) {
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::idesc_f16_f32(128, N);
            const uint32_t a_addr = ptx::smem_u32(smem);
            const uint32_t b_addr = a_addr + 16384;
            long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t d = (MODE == 1 || MODE == 3) ? ((k & 1) * 256) : 0;
                    if (MODE == 2 || MODE == 3)
                        ptx::mma_f16_ts(d % 256, 256 + (MODE == 3 ? 0 : 0) + k * 8 + 128 * ((it & 1)),
                                        ptx::smem_desc_sw128(b_addr) + 2 * k, idesc, 1);
                    else if (MODE == 4)
                        ptx::mma_f16_ss(d, ptx::smem_desc_sw128(a_addr) + 2 * k, desc_nosw(b_addr + k * 1024, 512, 128),
                                        idesc, 1);
                    else
                        ptx::mma_f16_ss(d, ptx::smem_desc_sw128(a_addr) + 2 * k, ptx::smem_desc_sw128(b_addr) + 2 * k,
                                        idesc, 1);
                }
            }
            ptx::mma_commit(&bar);
            ptx::mbar_wait(&bar, 0);
            long long t1 = clock64();
            out[blockIdx.x] = t1 - t0;
        }
        __syncwarp();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc<512>(0);
}

template <int N, int MODE>
void run(const char* name, int grid) {
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    const int iters = 256;
    cudaFuncSetAttribute(bench<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    bench<N, MODE><<<grid, 128, 64 * 1024>>>(d, iters);
    bench<N, MODE><<<grid, 128, 64 * 1024>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-34s N=%3d grid=%3d: %7.1f cycles/MMA (%s)\n", name, N, grid, (double)mx / (iters * 4), cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int grid : {1, 148}) {
        run<32, 0>("SS sw128, one acc", grid);
        run<64, 0>("SS sw128, one acc", grid);
        run<128, 0>("SS sw128, one acc", grid);
        run<256, 0>("SS sw128, one acc", grid);
        run<64, 1>("SS sw128, two accs", grid);
        run<128, 1>("SS sw128, two accs", grid);
        run<32, 2>("TS, one acc", grid);
        run<64, 2>("TS, one acc", grid);
        run<128, 2>("TS, one acc", grid);
        run<256, 2>("TS, one acc", grid);
        run<32, 4>("SS, no-swizzle B", grid);
    }
    return 0;
}
