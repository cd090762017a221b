// tcgen05 / TMEM / TMA GEMM used by every dense contraction on the path (ViT patch-embed, QKV,
// out-proj, MLP; LLM prefill projections; chunked projector / gate).
//
//   acc[a, b] = sum_k A[a, k] * B[b, k]          A: [Ma, K]  B: [Nb, K]  both row-major (K-major)
//
// One CTA computes a 128 x BN tile: A rows live on the 128 TMEM lanes, B rows on BN TMEM columns.
//   swap = 0 : A = activations (tokens), B = weights (features)  -> out[token][feature]
//   swap = 1 : A = weights (features),  B = activations (tokens) -> out[token][feature]  (small-token
//              GEMMs: the 128-wide MMA M dimension is filled by weight rows, tokens ride on N >= 16)
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> global; warp w may only touch TMEM lanes 32*(w%4)..+31).
// Operand tiles are 64-element (128-byte) K slabs in SWIZZLE_128B layout, NSTAGE-deep mbarrier ring.
#pragma once
#include "ptx.cuh"

namespace smb {

enum EpiMode : int {
    EPI_STORE = 0,       // out = T(acc + bias)
    EPI_QUICK_GELU = 1,  // h = T(acc + bias); out = T(h * T(sigmoid(T(1.702 h))))   (HF QuickGELUActivation in T)
    EPI_RESIDUAL = 2,    // out = T(out + T(acc + bias))                              (in-place residual stream)
    EPI_STORE_F32 = 3,   // out(float) = acc + bias
};

struct GemmArgs {
    int Ma, Nb, K;
    const void* bias;  // T[features] or nullptr
    void* out;
    int ldo;     // elements
    int swap;    // see header comment
    int bn;      // tile width on the B side (multiple of 16, 16..256); TMA box of tmap_b has bn rows
    int nstage;  // smem ring depth
    int epi;     // EpiMode
};

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;
constexpr int kGemmThreads = 192;
constexpr int kGemmMaxStages = 8;
constexpr int kGemmSmemBudget = 200 * 1024;

inline int gemm_stage_bytes(int bn) { return kGemmBM * kGemmBK * 2 + bn * kGemmBK * 2; }
inline int gemm_num_stages(int bn) {
    int s = kGemmSmemBudget / gemm_stage_bytes(bn);
    return s > kGemmMaxStages ? kGemmMaxStages : s;
}
inline int gemm_smem_bytes(int bn) { return gemm_num_stages(bn) * gemm_stage_bytes(bn) + 1024 + 256; }

template <typename T> __device__ __forceinline__ float quick_gelu_t(float h) {
    // every intermediate is materialised in T by the reference: 1.702*x, sigmoid(.), x*(.)
    float a = rnd<T>(1.702f * h);
    float s = rnd<T>(1.0f / (1.0f + __expf(-a)));
    return h * s;
}

template <typename T>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmArgs args) {
    const int BN = args.bn;
    const int NSTAGE = args.nstage;
    constexpr int A_BYTES = kGemmBM * kGemmBK * 2;
    const int B_BYTES = BN * kGemmBK * 2;
    const uint32_t TMEM_COLS = BN <= 32 ? 32u : BN <= 64 ? 64u : BN <= 128 ? 128u : 256u;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + NSTAGE * A_BYTES;  // B_BYTES is a multiple of 2048 -> stays 1024-aligned
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NSTAGE * (A_BYTES + B_BYTES));
    uint64_t* empty_bar = full_bar + kGemmMaxStages;
    uint64_t* accum_bar = empty_bar + kGemmMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int a0 = blockIdx.x * kGemmBM;  // first A row of this tile
    const int b0 = blockIdx.y * BN;       // first B row of this tile
    const int num_kb = (args.K + kGemmBK - 1) / kGemmBK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer
            // weights are streamed once; activations are re-read by every CTA column -> keep them in L2
            const uint64_t pol_a = args.swap ? kEvictFirst : kEvictLast;
            const uint64_t pol_b = args.swap ? kEvictLast : kEvictNormal;
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
                tma_load_2d(smem_a + stage * A_BYTES, &tmap_a, &full_bar[stage], kb * kGemmBK, a0, pol_a);
                tma_load_2d(smem_b + stage * B_BYTES, &tmap_b, &full_bar[stage], kb * kGemmBK, b0, pol_b);
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------- MMA issuer (one thread)
            const uint32_t idesc = umma_idesc_f16(kGemmBM, BN, Cvt<T>::kBf16);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(smem_a + stage * A_BYTES));
                const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(smem_b + stage * B_BYTES));
#pragma unroll
                for (int k = 0; k < kGemmBK / 16; ++k) {
                    // advance 16 elements = 32 bytes inside the 128-byte swizzle row: +2 in (addr>>4) units
                    umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
            umma_commit(accum_bar);  // accumulator complete
        }
    } else {
        // ---------------- epilogue warps 2..5
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int lane_base = (warp & 3) * 32;
        const int a_row = a0 + lane_base + lane;  // A row owned by this thread
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_base) << 16);
        T* out_t = reinterpret_cast<T*>(args.out);
        float* out_f = reinterpret_cast<float*>(args.out);
        const T* bias = reinterpret_cast<const T*>(args.bias);
        const int epi = args.epi;
        const bool a_ok = a_row < args.Ma;
        if (!args.swap) {
            // thread = token row; columns = features; 16 contiguous features per step
#pragma unroll 1
            for (int c = 0; c < BN; c += 16) {
                uint32_t r[16];
                __syncwarp();
                tmem_ld_x16(taddr + c, r);
                tmem_wait_ld();
                const int col = b0 + c;
                if (a_ok && col < args.Nb) {
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
                    if (bias != nullptr) {
                        uint4 q0 = *reinterpret_cast<const uint4*>(bias + col);
                        uint4 q1 = *reinterpret_cast<const uint4*>(bias + col + 8);
                        uint32_t bw[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float2 f = Cvt<T>::unpack2(bw[i]);
                            v[2 * i] += f.x;
                            v[2 * i + 1] += f.y;
                        }
                    }
                    const size_t off = static_cast<size_t>(a_row) * args.ldo + col;
                    if (epi == EPI_STORE_F32) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            *reinterpret_cast<float4*>(out_f + off + i) =
                                make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    } else {
                        if (epi == EPI_QUICK_GELU) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = quick_gelu_t<T>(rnd<T>(v[i]));
                        } else if (epi == EPI_RESIDUAL) {
                            uint4 x0 = *reinterpret_cast<const uint4*>(out_t + off);
                            uint4 x1 = *reinterpret_cast<const uint4*>(out_t + off + 8);
                            uint32_t xw[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                float2 f = Cvt<T>::unpack2(xw[i]);
                                v[2 * i] = f.x + rnd<T>(v[2 * i]);
                                v[2 * i + 1] = f.y + rnd<T>(v[2 * i + 1]);
                            }
                        }
                        uint4 o0, o1;
                        o0.x = Cvt<T>::pack2(v[0], v[1]);   o0.y = Cvt<T>::pack2(v[2], v[3]);
                        o0.z = Cvt<T>::pack2(v[4], v[5]);   o0.w = Cvt<T>::pack2(v[6], v[7]);
                        o1.x = Cvt<T>::pack2(v[8], v[9]);   o1.y = Cvt<T>::pack2(v[10], v[11]);
                        o1.z = Cvt<T>::pack2(v[12], v[13]); o1.w = Cvt<T>::pack2(v[14], v[15]);
                        *reinterpret_cast<uint4*>(out_t + off) = o0;
                        *reinterpret_cast<uint4*>(out_t + off + 8) = o1;
                    }
                }
            }
        } else {
            // thread = feature; columns = tokens; a warp writes 32 consecutive features of one token
            const float bv = (bias != nullptr && a_ok) ? Cvt<T>::to_f(bias[a_row]) : 0.0f;
#pragma unroll 1
            for (int c = 0; c < BN; c += 16) {
                uint32_t r[16];
                __syncwarp();
                tmem_ld_x16(taddr + c, r);
                tmem_wait_ld();
                if (a_ok) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int tok = b0 + c + i;
                        if (tok < args.Nb) {
                            float v = __uint_as_float(r[i]) + bv;
                            const size_t off = static_cast<size_t>(tok) * args.ldo + a_row;
                            if (epi == EPI_STORE_F32) {
                                out_f[off] = v;
                            } else {
                                if (epi == EPI_QUICK_GELU) v = quick_gelu_t<T>(rnd<T>(v));
                                if (epi == EPI_RESIDUAL) v = Cvt<T>::to_f(out_t[off]) + rnd<T>(v);
                                out_t[off] = Cvt<T>::from_f(v);
                            }
                        }
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace smb
