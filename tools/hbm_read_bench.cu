// Pure-read HBM bandwidth probes for the weight-streaming (decode) design: which access scheme gets closest to the
// measured copy peak (MEASURED_PEAKS.json hbm_gbs)?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o
// tools/bin/hbm_read_bench tools/hbm_read_bench.cu ; run on the GPU box: tools/bin/hbm_read_bench [GiB]
//   ldg   : 16-byte ld.global.nc.L1::no_allocate, U loads in flight per lane, OCC CTAs of 512 threads per SM,
//           per-CTA contiguous regions (pattern 0) or grid-interleaved 8 KB blocks (pattern 1)
//   bulk  : one CTA per SM, cp.async.bulk (1-D) into an NS-stage ring of SB-byte slots, 8 consumer warps read the
//           slot back from shared memory (LDS.128 + FADD) and release it through an mbarrier
#include <array>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int U, int PATTERN>
__global__ void __launch_bounds__(512) ldg_kernel(const uint4* __restrict__ src, size_t n16, float* out) {
    // n16: number of 16-byte elements; each CTA reads n16 / gridDim.x of them
    const size_t per_cta = n16 / gridDim.x;
    float acc = 0.f;
    if (PATTERN == 0) {
        const uint4* p = src + per_cta * blockIdx.x;
        for (size_t i = threadIdx.x; i + (U - 1) * 512 < per_cta; i += U * 512) {
            uint4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ldg_stream(p + i + u * 512);
#pragma unroll
            for (int u = 0; u < U; ++u) acc += __uint_as_float(v[u].x) + __uint_as_float(v[u].y) + __uint_as_float(v[u].z) + __uint_as_float(v[u].w);
        }
    } else {
        // grid-interleaved: block b of U*512 16-byte elements goes to CTA b % grid
        const size_t blk = U * 512;
        const size_t nblk = n16 / blk;
        for (size_t b = blockIdx.x; b < nblk; b += gridDim.x) {
            const uint4* p = src + b * blk + threadIdx.x;
            uint4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ldg_stream(p + u * 512);
#pragma unroll
            for (int u = 0; u < U; ++u) acc += __uint_as_float(v[u].x) + __uint_as_float(v[u].y) + __uint_as_float(v[u].z) + __uint_as_float(v[u].w);
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

// double-buffered register variant: the next batch is issued before the current one is consumed
template <int U>
__global__ void __launch_bounds__(512) ldg2_kernel(const uint4* __restrict__ src, size_t n16, float* out) {
    const size_t per_cta = n16 / gridDim.x;
    const uint4* p = src + per_cta * blockIdx.x;
    float acc = 0.f;
    uint4 a[U], b[U];
    size_t i = threadIdx.x;
    const size_t step = U * 512;
    if (i + (U - 1) * 512 < per_cta) {
#pragma unroll
        for (int u = 0; u < U; ++u) a[u] = ldg_stream(p + i + u * 512);
    }
    for (; i + (U - 1) * 512 < per_cta; i += 2 * step) {
        const bool hb = i + step + (U - 1) * 512 < per_cta;
        if (hb) {
#pragma unroll
            for (int u = 0; u < U; ++u) b[u] = ldg_stream(p + i + step + u * 512);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += __uint_as_float(a[u].x) + __uint_as_float(a[u].y) + __uint_as_float(a[u].z) + __uint_as_float(a[u].w);
        if (i + 2 * step + (U - 1) * 512 < per_cta) {
#pragma unroll
            for (int u = 0; u < U; ++u) a[u] = ldg_stream(p + i + 2 * step + u * 512);
        }
        if (hb) {
#pragma unroll
            for (int u = 0; u < U; ++u) acc += __uint_as_float(b[u].x) + __uint_as_float(b[u].y) + __uint_as_float(b[u].z) + __uint_as_float(b[u].w);
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    uint32_t spins = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (++spins > (1u << 26)) { printf("mbar timeout\n"); __trap(); }
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// one CTA per SM; warp 8 = producer; warps 0..7 consume.  chunk: bytes per bulk copy (a slot is filled by SB / chunk copies)
__global__ void __launch_bounds__(288) bulk_kernel(const uint8_t* __restrict__ src, size_t bytes, int NS, int SB, int chunk, int pattern, float* out, int reps = 1) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full[16], empty[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t nslots_total = bytes / SB;
    size_t my0, mycount, stride;
    if (pattern == 0) { mycount = nslots_total / gridDim.x; my0 = mycount * blockIdx.x; stride = 1; }
    else { mycount = nslots_total / gridDim.x; my0 = blockIdx.x; stride = gridDim.x; }
    const size_t region = mycount;          // reps > 1: the CTA walks its region reps times (L2-resident when the buffer is small)
    mycount *= reps;
    if (warp == 8) {
        if (lane == 0) {
            for (size_t i = 0; i < mycount; ++i) {
                const int s = static_cast<int>(i % NS);
                const uint32_t ph = static_cast<uint32_t>((i / NS) & 1);
                if (i >= static_cast<size_t>(NS)) mbar_wait(&empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&full[s], SB);
                const uint8_t* g = src + (my0 + (i % region) * stride) * SB;
                for (int c = 0; c < SB; c += chunk) bulk_g2s(smem + static_cast<size_t>(s) * SB + c, g + c, chunk, &full[s]);
            }
        }
    } else {
        float acc = 0.f;
        for (size_t i = 0; i < mycount; ++i) {
            const int s = static_cast<int>(i % NS);
            const uint32_t ph = static_cast<uint32_t>((i / NS) & 1);
            mbar_wait(&full[s], ph);
            const uint4* p = reinterpret_cast<const uint4*>(smem + static_cast<size_t>(s) * SB);
            for (int k = threadIdx.x; k < SB / 16; k += 256) {
                const uint4 v = p[k];
                acc += __uint_as_float(v.x) + __uint_as_float(v.y) + __uint_as_float(v.z) + __uint_as_float(v.w);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (acc == 123.456f) out[0] = acc;
    }
}

template <typename F>
float time_ms(F&& f, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char** argv) {
    const double gib = argc > 1 ? atof(argv[1]) : 4.0;
    const size_t bytes = static_cast<size_t>(gib * (1ull << 30)) / (148 * 4 * 1024 * 1024ull) * (148 * 4 * 1024 * 1024ull);   // multiple of 148 x 4 MiB
    uint8_t* buf;
    float* out;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMalloc(&out, 4));
    CK(cudaMemset(buf, 1, bytes));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, buffer %.2f GiB\n", prop.name, sms, bytes / double(1ull << 30));
    const size_t n16 = bytes / 16;
    auto report = [&](const char* name, float ms) { printf("%-44s %8.3f ms  %8.1f GB/s\n", name, ms, bytes / ms * 1e-6); fflush(stdout); };
    char nm[128];
#define LDG(U, P, OCC) { snprintf(nm, sizeof nm, "ldg U=%d pattern=%d occ=%d", U, P, OCC); \
        report(nm, time_ms([&] { ldg_kernel<U, P><<<sms * OCC, 512>>>(reinterpret_cast<const uint4*>(buf), n16, out); }, 5)); }
    LDG(4, 0, 1) LDG(4, 0, 2) LDG(4, 0, 4) LDG(8, 0, 1) LDG(8, 0, 2) LDG(8, 0, 4) LDG(16, 0, 1) LDG(16, 0, 2)
    LDG(4, 1, 2) LDG(4, 1, 4) LDG(8, 1, 2) LDG(8, 1, 4) LDG(16, 1, 2)
#define LDG2(U, OCC) { snprintf(nm, sizeof nm, "ldg2 (double-buffered) U=%d occ=%d", U, OCC); \
        report(nm, time_ms([&] { ldg2_kernel<U><<<sms * OCC, 512>>>(reinterpret_cast<const uint4*>(buf), n16, out); }, 5)); }
    LDG2(2, 2) LDG2(4, 1) LDG2(4, 2) LDG2(4, 4) LDG2(8, 1) LDG2(8, 2)
    CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int cfgs[][4] = {{4, 16384, 16384, 0}, {8, 16384, 16384, 0}, {12, 16384, 16384, 0}, {6, 32768, 32768, 0}, {6, 32768, 8192, 0},
                           {3, 65536, 65536, 0}, {3, 65536, 16384, 0}, {16, 8192, 8192, 0}, {12, 16384, 4096, 0},
                           {8, 16384, 16384, 1}, {6, 32768, 32768, 1}, {12, 16384, 16384, 1}, {6, 32768, 8192, 1}};
    for (auto& c : cfgs) {
        snprintf(nm, sizeof nm, "bulk NS=%d SB=%d chunk=%d pattern=%d", c[0], c[1], c[2], c[3]);
        report(nm, time_ms([&] { bulk_kernel<<<sms, 288, c[0] * c[1]>>>(buf, bytes, c[0], c[1], c[2], c[3], out); }, 5));
    }
    // two CTAs per SM with half the ring each
    const int cfgs2[][4] = {{4, 16384, 16384, 0}, {6, 16384, 16384, 0}, {3, 32768, 32768, 0}, {12, 8192, 8192, 0}, {6, 16384, 16384, 1}};
    for (auto& c : cfgs2) {
        snprintf(nm, sizeof nm, "bulk x2/SM NS=%d SB=%d chunk=%d pattern=%d", c[0], c[1], c[2], c[3]);
        report(nm, time_ms([&] { bulk_kernel<<<sms * 2, 288, c[0] * c[1]>>>(buf, bytes, c[0], c[1], c[2], c[3], out); }, 5));
    }
    // L2-resident: every SM re-reads its own 192 KB / 384 KB region 256 times (28 / 57 MB in total): what a bulk-copy ring can ingest from L2
    for (int kb : {192, 384}) {
        const size_t small = static_cast<size_t>(sms) * kb * 1024;
        const int reps = 256;
        for (auto& c : {std::array<int, 3>{6, 32768, 32768}, std::array<int, 3>{3, 65536, 65536}, std::array<int, 3>{12, 16384, 16384}}) {
            float ms = time_ms([&] { bulk_kernel<<<sms, 288, c[0] * c[1]>>>(buf, small, c[0], c[1], c[2], 0, out, reps); }, 3);
            printf("bulk L2-resident %3d KB/SM NS=%d SB=%d      %8.3f ms  %8.1f GB/s\n", kb, c[0], c[1], ms, double(small) * reps / ms * 1e-6);
        }
    }
    // reference: device-to-device copy (read + write bytes), as MEASURED_PEAKS.json counts it
    uint8_t* dst;
    if (cudaMalloc(&dst, bytes) == cudaSuccess) {
        float ms = time_ms([&] { cudaMemcpyAsync(dst, buf, bytes, cudaMemcpyDeviceToDevice); }, 5);
        printf("%-44s %8.3f ms  %8.1f GB/s (read + write)\n", "cudaMemcpy D2D", ms, 2.0 * bytes / ms * 1e-6);
    }
    return 0;
}
