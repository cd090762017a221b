// Microbenchmark: L2 -> shared-memory delivery rate of TMA tile loads (cp.async.bulk.tensor.2d, SWIZZLE_128B boxes of
// 64 x 128 fp16 = 16 KB, the GEMM's operand tile) per SM, with 1 .. 148 CTAs pulling from an L2-resident matrix.
// No MMA, no epilogue: a producer thread keeps NSTAGE loads in flight, a consumer thread recycles the stages.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I streammind_b200/csrc tools/tma_bench.cu -o tools/bin/tma_bench -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ptx.cuh"
using namespace smb;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(64, 1) tma_rate_kernel(const __grid_constant__ CUtensorMap tmap, int nstage, int nloads,
                                                        int rows_total, int kblocks, int same_tile, int per_stage, long long* out_cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full_bar[12], empty_bar[12];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 12; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        fence_mbar_init();
    }
    __syncthreads();
    const int tiles_m = rows_total / 128;
    const long long t0 = clock64();
    if (threadIdx.x == 0) {            // producer: coordinates advance incrementally (no div/mod in the issue loop)
        int stage = 0; uint32_t phase = 0;
        int kb = 0, mt = same_tile ? 0 : (blockIdx.x * 3) % tiles_m;
        for (int i = 0; i < nloads; ++i) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], 16384 * per_stage);
            for (int j = 0; j < per_stage; ++j) {
                int m2 = mt + j; if (m2 >= tiles_m) m2 -= tiles_m;
                tma_load_2d(smem + (stage * per_stage + j) * 16384, &tmap, &full_bar[stage], kb * 64, m2 * 128, kEvictLast);
            }
            if (!same_tile) { if (++kb == kblocks) { kb = 0; if (++mt == tiles_m) mt = 0; } }
            if (++stage == nstage) { stage = 0; phase ^= 1; }
        }
    } else if (threadIdx.x == 32) {    // consumer: recycle the stage as soon as it landed
        int stage = 0; uint32_t phase = 0;
        for (int i = 0; i < nloads; ++i) {
            mbar_wait(&full_bar[stage], phase);
            mbar_arrive(&empty_bar[stage]);
            if (++stage == nstage) { stage = 0; phase ^= 1; }
        }
        out_cycles[blockIdx.x] = clock64() - t0;
    }
}

int main() {
    const int rows = 4096, K = 1024;   // 8 MB fp16: L2 resident
    void* d; cudaMalloc(&d, size_t(rows) * K * 2); cudaMemset(d, 0, size_t(rows) * K * 2);
    long long* dc; cudaMalloc(&dc, 148 * sizeof(long long));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap m;
    cuuint64_t gdim[2] = {cuuint64_t(K), cuuint64_t(rows)}, gstr[1] = {cuuint64_t(K) * 2};
    cuuint32_t box[2] = {64, 128}, estr[2] = {1, 1};
    CUresult r = reinterpret_cast<PFN_encodeTiled>(fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", int(r)); return 1; }
    cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    std::vector<long long> h(148);
    printf("grid boxes/stage nstage same_tile : bytes/clk/SM (mean, min over CTAs)   aggregate TB/s @1.965 GHz\n");
    for (int grid : {1, 74, 148})
        for (int per_stage : {1, 2, 3})
          for (int nstage : {2, 4})
            for (int same : {0, 1}) {
                const int nloads = 4096;
                for (int rep = 0; rep < 2; ++rep) {
                    tma_rate_kernel<<<grid, 64, nstage * per_stage * 16384 + 1024>>>(m, nstage, nloads, rows, K / 64, same, per_stage, dc);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                }
                cudaMemcpy(h.data(), dc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
                double mean = 0, mx = 0;
                for (int i = 0; i < grid; ++i) { mean += h[i]; mx = h[i] > mx ? h[i] : mx; }
                mean /= grid;
                const double bpc = nloads * 16384.0 * per_stage / mean, bpc_min = nloads * 16384.0 * per_stage / mx;
                printf("%3d boxes/stage %d nstage %2d same %d : %.1f %.1f   %.2f\n", grid, per_stage, nstage, same, bpc, bpc_min, bpc * grid * 1.965e9 / 1e12);
            }
    return 0;
}
