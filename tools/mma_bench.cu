// Microbenchmark: issue rate of tcgen05.mma.cta_group::1.kind::f16 (M=128, N in {64,128,256}, K=16) from
// SWIZZLE_128B K-major shared-memory operands, as the GEMM mainloop issues it (4 MMAs per 64-wide K slab,
// one tcgen05.commit per slab).  No TMA, no epilogue: isolates the tensor pipe + smem operand fetch.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I streammind_b200/csrc tools/mma_bench.cu -o tools/bin/mma_bench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ptx.cuh"
using namespace smb;

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int BN, int nslab, int nstage, int commit_every,
                                                           long long* out_cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bars[16];
    __shared__ uint32_t tmem_slot;
    const int A_BYTES = 128 * 128, B_BYTES = BN * 128;
    for (int i = threadIdx.x; i < nstage * (A_BYTES + B_BYTES) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    if (threadIdx.x < 32) { tmem_alloc(&tmem_slot, 256); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x < 32) {
        const uint32_t idesc = umma_idesc_f16(128, BN, false);
        long long t0 = 0, t1 = 0;
        if (elect_one_sync()) {
            t0 = clock64();
            int stage = 0, ncommit = 0;
            for (int s = 0; s < nslab; ++s) {
                const uint64_t ad = umma_desc_sw128_kmajor(smem_u32(smem + stage * A_BYTES));
                const uint64_t bd = umma_desc_sw128_kmajor(smem_u32(smem + nstage * A_BYTES + stage * B_BYTES));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(tmem, ad + 2 * k, bd + 2 * k, idesc, (s | k) ? 1u : 0u);
                if (commit_every && (s % commit_every) == commit_every - 1 && s != nslab - 1) {
                    umma_commit(&bars[1 + (ncommit & 7)]);
                    ++ncommit;
                }
                if (++stage == nstage) stage = 0;
            }
            umma_commit(&bars[0]);
        }
        __syncwarp();
        mbar_wait(&bars[0], 0);
        t1 = clock64();
        long long t0b = __shfl_sync(0xffffffffu, t0, 0);   // elected lane is not necessarily lane 0: take the max
        (void)t0b;
        unsigned long long mx = 0;
        for (int l = 0; l < 32; ++l) {
            const unsigned long long v = __shfl_sync(0xffffffffu, static_cast<unsigned long long>(t0), l);
            mx = v > mx ? v : mx;
        }
        if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - static_cast<long long>(mx);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int main() {
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    std::vector<long long> h(148);
    printf("BN nstage commit_every grid : cycles/MMA(K=16)  [floor = BN/2 cycles]\n");
    for (int grid : {1, 148})
        for (int BN : {32, 64, 128, 256})
            for (int commit_every : {0, 1})
                for (int nstage : {1, 4}) {
                    const int nslab = 512;
                    const int smem = nstage * (128 * 128 + BN * 128) + 2048;
                    for (int rep = 0; rep < 2; ++rep) {
                        mma_rate_kernel<<<grid, 128, smem>>>(BN, nslab, nstage, commit_every, d);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                    }
                    cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
                    double mean = 0; long long mx = 0;
                    for (int i = 0; i < grid; ++i) { mean += h[i]; mx = h[i] > mx ? h[i] : mx; }
                    mean /= grid;
                    printf("%3d %d %d %3d : mean %.1f max %.1f\n", BN, nstage, commit_every, grid, mean / (nslab * 4.0),
                           mx / (nslab * 4.0));
                }
    return 0;
}
