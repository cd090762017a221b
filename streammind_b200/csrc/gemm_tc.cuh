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
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..9 = epilogue (TMEM -> registers -> smem/global; warp w may only touch TMEM lanes 32*(w%4)..+31, so
// warps w and w+4 share a lane quarter and take alternate 32-column chunks).
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
    int tma_store;  // non-swapped T outputs with bn % 64 == 0: stage the tile in smem, write it with TMA
    int split_k;    // > 1: blockIdx.z takes k-blocks [z*kb_per, (z+1)*kb_per) and writes an fp32 partial tile to
                    // out + z*split_stride (EPI_STORE_F32, no bias); the consumer sums the partials in fixed order
    long long split_stride;   // elements between partial outputs
    int cluster_n;  // CTAs per cluster along blockIdx.y (1, 2, 4 or 8): they share the A tile, each CTA loads
                    // 128/cluster_n of its rows and TMA-multicasts them to the whole cluster (L2 reads of A / cluster_n)
    int w_tiled;    // the weight operand is stored pre-tiled in HBM: [n_tile][k_block][128 rows][64 cols], each
                    // 16 KiB operand tile contiguous (full-rate DRAM bursts instead of 128-byte strided reads)
    int w_kb;       // k-blocks per n_tile in that layout
    int two_producers;   // second TMA producer warp (warp 10) takes the odd k-blocks
    int bm2;        // 1: the CTA owns a 256-row tile = two 128-row halves that share every B (weight) stage; two
                    // accumulators in TMEM (columns 0 and 256).  Halves the weight bytes an SM ingests per output
                    // element (the L2 -> SM rate, ~48 B/clk/SM measured, bounds the 128-row tile).  Non-swapped only.
    int pre_weights;   // request the first ring of weight tiles before griddepcontrol.wait (see the producer)
    int dbg_mode;   // microbenchmark aid: 1 = no MMA issue (TMA + barriers only), 2 = no TMA (MMA + barriers only)
    long long* dbg; // optional: per-CTA phase timestamps (globaltimer ns), 8 slots per CTA
};

__device__ __forceinline__ long long gtimer() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;
constexpr int kGemmThreads2 = 352;  // gemm_tc_kernel: + a second TMA producer warp (warp 10)
constexpr int kGemmEpiThreads = 256;
constexpr int kGemmMaxStages = 8;
constexpr int kGemmSmemBudget = 200 * 1024;
inline int gemm_stage_bytes(int bn, int bm2 = 0) { return (bm2 ? 2 : 1) * kGemmBM * kGemmBK * 2 + bn * kGemmBK * 2; }
inline int gemm_num_stages(int bn, int bm2 = 0) {
    int s = kGemmSmemBudget / gemm_stage_bytes(bn, bm2);
    return s > kGemmMaxStages ? kGemmMaxStages : s;
}
inline int gemm_smem_bytes(int bn, int bm2 = 0) { return gemm_num_stages(bn, bm2) * gemm_stage_bytes(bn, bm2) + 1024 + 256 + 1024; }

template <typename T> __device__ __forceinline__ float quick_gelu_t(float h) {
    // every intermediate is materialised in T by the reference: 1.702*x, sigmoid(.), x*(.)
    // sigmoid through MUFU.EX2 + MUFU.RCP (approximation error << one T ulp; an IEEE division here made
    // this epilogue 8x slower than the mainloop)
    const float a = rnd<T>(1.702f * h);
    const float s = rnd<T>(__fdividef(1.0f, 1.0f + __expf(-a)));
    return h * s;
}

// One 32-column accumulator chunk of one thread, specialised at compile time so the element loop has no
// mode branches (the runtime-flag version spent ~150 clocks per element on branch resolution).
//   SWAP = false: thread owns a token row, the chunk is 32 consecutive features (4 x 16-byte stores)
//   SWAP = true : thread owns a feature, the chunk is 32 tokens (2-byte stores, coalesced across the warp)
//   TMAST: write into the swizzled smem staging tile (TMA store afterwards) instead of global memory
struct EpiCtx {
    const GemmArgs* args;
    uint8_t* stage;        // smem staging tile (the free operand ring)
    const float* bias_s;   // non-swap: bias of this tile's columns in smem
    float bv;              // swap: bias of this thread's feature
    int a_row, b0, BN, lane_row;   // lane_row = row inside the 128-row tile
    size_t out_off;                // split-K: element offset of this split's partial output
};

template <typename T, bool SWAP, int EPI, bool TMAST>
__device__ __forceinline__ void epi_chunk(const EpiCtx& cx, int cbase, const uint32_t (&rb)[32], const uint4 (&xr)[4]) {
    const GemmArgs& args = *cx.args;
    T* out_t = reinterpret_cast<T*>(args.out);
    float* out_f = reinterpret_cast<float*>(args.out) + cx.out_off;
    if constexpr (!SWAP) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int col = cx.b0 + cbase + q * 8;
            if (col < args.Nb && cbase + q * 8 < cx.BN) {
                const float4 bA = *reinterpret_cast<const float4*>(cx.bias_s + cbase + q * 8);
                const float4 bB = *reinterpret_cast<const float4*>(cx.bias_s + cbase + q * 8 + 4);
                float v[8] = {__uint_as_float(rb[q * 8 + 0]) + bA.x, __uint_as_float(rb[q * 8 + 1]) + bA.y,
                              __uint_as_float(rb[q * 8 + 2]) + bA.z, __uint_as_float(rb[q * 8 + 3]) + bA.w,
                              __uint_as_float(rb[q * 8 + 4]) + bB.x, __uint_as_float(rb[q * 8 + 5]) + bB.y,
                              __uint_as_float(rb[q * 8 + 6]) + bB.z, __uint_as_float(rb[q * 8 + 7]) + bB.w};
                const size_t off = static_cast<size_t>(cx.a_row) * args.ldo + col;
                if constexpr (EPI == EPI_STORE_F32) {
                    *reinterpret_cast<float4*>(out_f + off) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4*>(out_f + off + 4) = make_float4(v[4], v[5], v[6], v[7]);
                } else {
                    if constexpr (EPI == EPI_QUICK_GELU) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = quick_gelu_t<T>(rnd<T>(v[i]));
                    } else if constexpr (EPI == EPI_RESIDUAL) {
                        const uint4 x0 = xr[q];
                        const uint32_t xw[4] = {x0.x, x0.y, x0.z, x0.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 f = Cvt<T>::unpack2(xw[i]);
                            v[2 * i] = f.x + rnd<T>(v[2 * i]);
                            v[2 * i + 1] = f.y + rnd<T>(v[2 * i + 1]);
                        }
                    }
                    uint4 o;
                    o.x = Cvt<T>::pack2(v[0], v[1]); o.y = Cvt<T>::pack2(v[2], v[3]);
                    o.z = Cvt<T>::pack2(v[4], v[5]); o.w = Cvt<T>::pack2(v[6], v[7]);
                    if constexpr (TMAST) {
                        // staging tile: 64-column blocks of [128 rows x 128 B], SWIZZLE_128B
                        const int ct = cbase + q * 8, r = cx.lane_row;
                        const uint32_t dst = smem_u32(cx.stage) + (ct >> 6) * (kGemmBM * 128) + r * 128 +
                                             ((((ct & 63) >> 3) ^ (r & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
                    } else {
                        *reinterpret_cast<uint4*>(out_t + off) = o;
                    }
                }
            }
        }
    } else {
        const int fr = cx.lane_row;   // feature within the tile
        const uint32_t sbase = smem_u32(cx.stage) + (fr >> 6) * (cx.BN * 128) + ((fr & 7) << 1);
        const int fchunk = (fr & 63) >> 3;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int row = cbase + i;            // token within the tile
            const int tok = cx.b0 + row;
            float v = __uint_as_float(rb[i]) + cx.bv;
            if constexpr (EPI == EPI_QUICK_GELU) v = quick_gelu_t<T>(rnd<T>(v));
            if constexpr (TMAST) {
                if (row < cx.BN) {
                    const uint32_t dst = sbase + row * 128 + ((fchunk ^ (row & 7)) << 4);
                    const T hv = Cvt<T>::from_f(v);
                    asm volatile("st.shared.b16 [%0], %1;" ::"r"(dst), "h"(*reinterpret_cast<const uint16_t*>(&hv)) : "memory");
                }
            } else if (tok < args.Nb && row < cx.BN) {
                const size_t off = static_cast<size_t>(tok) * args.ldo + cx.a_row;
                if constexpr (EPI == EPI_STORE_F32) {
                    out_f[off] = v;
                } else {
                    if constexpr (EPI == EPI_RESIDUAL) v = Cvt<T>::to_f(out_t[off]) + rnd<T>(v);
                    out_t[off] = Cvt<T>::from_f(v);
                }
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kGemmThreads2, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_c, const GemmArgs args) {
    const int BN = args.bn;
    const int NSTAGE = args.nstage;
    constexpr int A_HALF = kGemmBM * kGemmBK * 2;
    const int NH = args.bm2 ? 2 : 1;                 // 128-row halves of the A tile
    const int A_BYTES = NH * A_HALF;
    const int B_BYTES = BN * kGemmBK * 2;
    const bool ksplit = args.dbg_mode >= 3 && BN <= 128;   // experiment: one accumulator per K=16 step
    const uint32_t TMEM_COLS = (ksplit || args.bm2) ? 512u : BN <= 32 ? 32u : BN <= 64 ? 64u : BN <= 128 ? 128u : 256u;

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
    long long* dbg = args.dbg;   // per-launch record: [0] = min CTA start, [i] = max over CTAs of phase i end
    if (dbg && threadIdx.x == 0) atomicMin(reinterpret_cast<unsigned long long*>(dbg), static_cast<unsigned long long>(gtimer()));
    const int a0 = blockIdx.x * kGemmBM * (args.bm2 ? 2 : 1);  // first A row of this tile
    const int b0 = blockIdx.y * BN;       // first B row of this tile
    const int total_kb = (args.K + kGemmBK - 1) / kGemmBK;
    const int kb_per = args.split_k > 1 ? (total_kb + args.split_k - 1) / args.split_k : total_kb;
    const int kb_begin = args.split_k > 1 ? blockIdx.z * kb_per : 0;
    const int num_kb = max(0, min(total_kb, kb_begin + kb_per) - kb_begin);
    const int CS = args.cluster_n;
    const uint32_t crank = CS > 1 ? cluster_ctarank() : 0u;
    const uint16_t cmask = static_cast<uint16_t>((1u << CS) - 1u);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        if (args.tma_store) tma_prefetch_desc(&tmap_c);
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CS);   // one tcgen05.commit arrival from every CTA that reads the shared A slices
        }
        mbar_init(accum_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    pdl_trigger();
    tc_fence_before();
    if (CS > 1) cluster_sync_all();   // peers' barriers must be initialised before any multicast / remote arrive
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Two producer warps share the ring: one stage costs a single issuing thread ~370 cycles of barrier round trips
    // plus ~75 cycles per TMA box (tools/tma_bench.cu: 37 / 63 / 83 B/clk per SM with 1 / 2 / 3 boxes per stage,
    // whatever the ring depth), i.e. ~520 cycles per K slab at BN = 128 against 256-300 cycles of MMA -- the "250 ns
    // per slab whatever the tile width" of the first traces.  Warp 0 takes the even k-blocks, warp 10 the odd ones.
    const bool two_prod = args.two_producers != 0 && CS == 1 && args.dbg_mode == 0;
    if (warp == 0 || (warp == 10 && two_prod)) {
        const int pid = warp == 0 ? 0 : 1, pstep = two_prod ? 2 : 1;
        // ---------------- TMA producer: the whole warp walks the ring, one elected lane issues
        // weights are streamed once; activations are re-read by every CTA column -> keep them in L2
        const uint64_t pol_a = args.swap ? kEvictFirst : kEvictLast;
        const uint64_t pol_b = args.swap ? kEvictLast : kEvictNormal;
        const int a_slice = kGemmBM / CS;
        const bool a_is_tiled_w = args.w_tiled && args.swap;
        auto load_a = [&](int stage, int kg) {   // A tile: whole (CS = 1) or this CTA's 128/CS-row slice multicast to the cluster
            uint8_t* a_dst = smem_a + stage * A_BYTES + crank * a_slice * 128;
            const int ac0 = a_is_tiled_w ? 0 : kg * kGemmBK;
            const int ac1 = (a_is_tiled_w ? (blockIdx.x * args.w_kb + kg) * kGemmBM : a0) + crank * a_slice;
            if (CS > 1) tma_load_2d_mc(a_dst, &tmap_a, &full_bar[stage], ac0, ac1, cmask, pol_a);
            else tma_load_2d(a_dst, &tmap_a, &full_bar[stage], ac0, ac1, pol_a);
            if (NH == 2) tma_load_2d(a_dst + A_HALF, &tmap_a, &full_bar[stage], ac0, ac1 + kGemmBM, pol_a);   // rows past Ma: zero fill
        };
        auto load_b = [&](int stage, int kg) {
            if (!args.w_tiled || args.swap) {
                tma_load_2d(smem_b + stage * B_BYTES, &tmap_b, &full_bar[stage], kg * kGemmBK, b0, pol_b);
            } else {                  // B = tiled weights, BN rows = BN/128 whole tiles or a slice of one
                const int nld = BN > kGemmBM ? BN / kGemmBM : 1;
                for (int j = 0; j < nld; ++j)
                    tma_load_2d(smem_b + stage * B_BYTES + j * (kGemmBM * 128), &tmap_b, &full_bar[stage], 0,
                                ((b0 / kGemmBM + j) * args.w_kb + kg) * kGemmBM + (b0 % kGemmBM), pol_b);
            }
        };
        // The weight operand never depends on the previous kernel: with programmatic dependent launch this CTA is
        // often resident while its predecessor still runs, so the first ring of WEIGHT tiles is requested before
        // griddepcontrol.wait and only the activation tiles wait for it (hides the cold-HBM ramp of every GEMM).
        const bool can_pre = CS == 1 && args.dbg_mode == 0 && args.pre_weights != 0 && pid == 0;
        const int npre_all = (CS == 1 && args.dbg_mode == 0 && args.pre_weights != 0) ? min(NSTAGE, num_kb) : 0;
        const int npre = can_pre ? npre_all : 0;
        if (npre > 0 && elect_one_sync()) {
            for (int kb = 0; kb < npre; ++kb) {
                mbar_arrive_expect_tx(&full_bar[kb], A_BYTES + B_BYTES);
                if (args.swap) load_a(kb, kb_begin + kb); else load_b(kb, kb_begin + kb);
            }
        }
        __syncwarp();
        pdl_wait();   // the previous kernel's outputs (activations) are visible from here
        if (dbg && lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(dbg + 1), static_cast<unsigned long long>(gtimer()));
        if (npre > 0 && elect_one_sync()) {
            for (int kb = 0; kb < npre; ++kb) {
                if (args.swap) load_b(kb, kb_begin + kb); else load_a(kb, kb_begin + kb);
            }
        }
        __syncwarp();
        for (int kb = npre_all + pid; kb < num_kb; kb += pstep) {
            const int stage = kb % NSTAGE;
            const uint32_t phase = static_cast<uint32_t>(kb / NSTAGE) & 1u;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (args.dbg_mode == 2 || args.dbg_mode == 3) {
                if (elect_one_sync()) mbar_arrive(&full_bar[stage]);
            } else if (elect_one_sync()) {
                mbar_arrive_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
                const int kg = kb_begin + kb;   // global k-block index
                load_a(stage, kg);
                load_b(stage, kg);
            }
            __syncwarp();
        }
    } else if (warp == 10) {
        // second producer not in use for this launch
    } else if (warp == 1) {
        pdl_wait();
        // ---------------- MMA issuer: whole warp waits, one elected lane issues tcgen05.mma / commit
        const uint32_t idesc = umma_idesc_f16(kGemmBM, BN, Cvt<T>::kBf16);
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(smem_a + stage * A_BYTES));
                const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(smem_b + stage * B_BYTES));
                if (ksplit) {
#pragma unroll
                    for (int k = 0; k < kGemmBK / 16; ++k)
                        umma_f16(tmem_base + k * BN, adesc + 2 * k, bdesc + 2 * k, idesc, kb != 0 ? 1u : 0u);
                } else if (args.dbg_mode != 1) {
#pragma unroll
                    for (int k = 0; k < kGemmBK / 16; ++k) {
                        // advance 16 elements = 32 bytes inside the 128-byte swizzle row: +2 in (addr>>4) units
                        umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    if (NH == 2) {   // second 128-row half against the same B stage -> accumulator at column 256
                        const uint64_t adesc1 = umma_desc_sw128_kmajor(smem_u32(smem_a + stage * A_BYTES + A_HALF));
#pragma unroll
                        for (int k = 0; k < kGemmBK / 16; ++k)
                            umma_f16(tmem_base + 256u, adesc1 + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                }
                // frees the smem slot once these MMAs have read it -- in every CTA that multicasts into it
                if (CS > 1) umma_commit_mc(&empty_bar[stage], cmask);
                else umma_commit(&empty_bar[stage]);
                if (kb == num_kb - 1) umma_commit(accum_bar);  // accumulator complete
            }
            __syncwarp();
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
    } else {
        // ---------------- epilogue warps 2..9
        pdl_wait();
        const int lane_base = (warp & 3) * 32;
        const int chalf = (warp - 2) >> 2;   // 0: even 32-column chunks, 1: odd chunks
        int a_row = a0 + lane_base + lane;  // A row owned by this thread (first 128-row half)
        T* out_t = reinterpret_cast<T*>(args.out);
        float* out_f = reinterpret_cast<float*>(args.out);
        const T* bias = reinterpret_cast<const T*>(args.bias);
        const int epi = args.epi;
        bool a_ok = a_row < args.Ma;
        // While the mainloop runs: stage the bias of this tile's columns in smem (non-swapped layout)
        float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);  // [256] floats, 16-byte aligned
        if (!args.swap) {
            const int et = threadIdx.x - 64;  // 0..255
            for (int c = et; c < BN; c += kGemmEpiThreads) {
                const int col = b0 + c;
                bias_s[c] = (bias != nullptr && col < args.Nb) ? Cvt<T>::to_f(bias[col]) : 0.0f;
            }
            named_bar_sync(1, kGemmEpiThreads);
        }
        const float bv = (args.swap && bias != nullptr && a_ok) ? Cvt<T>::to_f(bias[a_row]) : 0.0f;
        if (dbg && threadIdx.x == 64) atomicMax(reinterpret_cast<unsigned long long*>(dbg + 2), static_cast<unsigned long long>(gtimer()));
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        if (dbg && threadIdx.x == 64) atomicMax(reinterpret_cast<unsigned long long*>(dbg + 3), static_cast<unsigned long long>(gtimer()));
        uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_base) << 16);
        // 32-column chunks, TMEM load of chunk c+1 (and residual read) in flight while chunk c is processed
        const int nchunk = (BN + 31) / 32;
        uint32_t rbA[32], rbB[32];
        uint4 xrA[4], xrB[4];
        auto prefetch_res = [&](int c, uint4 (&xr)[4]) {
            if (epi == EPI_RESIDUAL && !args.swap && a_ok) {
                const int col = b0 + c * 32;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (col + q * 8 < args.Nb && c * 32 + q * 8 < BN)
                        xr[q] = *reinterpret_cast<const uint4*>(out_t + static_cast<size_t>(a_row) * args.ldo + col + q * 8);
            }
        };
        EpiCtx cx;
        cx.args = &args; cx.stage = smem_a; cx.bias_s = bias_s; cx.bv = bv; cx.a_row = a_row; cx.b0 = b0; cx.BN = BN;
        cx.lane_row = lane_base + lane;
        cx.out_off = args.split_k > 1 ? static_cast<size_t>(blockIdx.z) * static_cast<size_t>(args.split_stride) : 0;
        const int mode = (args.swap ? 8 : 0) | (args.tma_store ? 4 : 0) | epi;   // uniform: one branch per chunk
        auto process = [&](int c, const uint32_t (&rb)[32], const uint4 (&xr)[4]) {
            if (!a_ok) return;
            const int cbase = c * 32;
            switch (mode) {
                case 0: epi_chunk<T, false, EPI_STORE, false>(cx, cbase, rb, xr); break;
                case 1: epi_chunk<T, false, EPI_QUICK_GELU, false>(cx, cbase, rb, xr); break;
                case 2: epi_chunk<T, false, EPI_RESIDUAL, false>(cx, cbase, rb, xr); break;
                case 3: epi_chunk<T, false, EPI_STORE_F32, false>(cx, cbase, rb, xr); break;
                case 4: epi_chunk<T, false, EPI_STORE, true>(cx, cbase, rb, xr); break;
                case 5: epi_chunk<T, false, EPI_QUICK_GELU, true>(cx, cbase, rb, xr); break;
                case 6: epi_chunk<T, false, EPI_RESIDUAL, true>(cx, cbase, rb, xr); break;
                case 8: epi_chunk<T, true, EPI_STORE, false>(cx, cbase, rb, xr); break;
                case 9: epi_chunk<T, true, EPI_QUICK_GELU, false>(cx, cbase, rb, xr); break;
                case 10: epi_chunk<T, true, EPI_RESIDUAL, false>(cx, cbase, rb, xr); break;
                case 11: epi_chunk<T, true, EPI_STORE_F32, false>(cx, cbase, rb, xr); break;
                case 12: epi_chunk<T, true, EPI_STORE, true>(cx, cbase, rb, xr); break;
                case 13: epi_chunk<T, true, EPI_QUICK_GELU, true>(cx, cbase, rb, xr); break;
                default: break;
            }
        };
        // BN is a multiple of 16: the last chunk may be half valid; x32 loads stay inside the TMEM
        // allocation because it is a power of two >= 32.  Chunk c+1 (TMEM load + residual read) is in
        // flight while chunk c is processed; ping-pong on two statically indexed register sets.
        // this warp's chunks: chalf, chalf + 2, ...
        const int nmine = nchunk > chalf ? (nchunk - chalf + 1) / 2 : 0;
        auto chunk_of = [&](int k) { return chalf + 2 * k; };
        const int half_stage_bytes = (BN / 64) * (kGemmBM * 128);   // staging tile of one 128-row half
        for (int hh = 0; hh < NH; ++hh) {
        if (hh == 1) {   // second 128-row half: accumulator at TMEM column 256, rows + 128, its own staging tile
            a_row += kGemmBM;
            a_ok = a_row < args.Ma;
            taddr += 256u;
            cx.a_row = a_row;
            cx.stage = smem_a + half_stage_bytes;
        }
        if (nmine > 0) {
            prefetch_res(chunk_of(0), xrA);
            __syncwarp();
            tmem_ld_x32(taddr + chunk_of(0) * 32, rbA);
            tmem_wait_ld();
        }
#pragma unroll 1
        for (int k = 0; k < nmine; k += 2) {
            const bool has1 = k + 1 < nmine, has2 = k + 2 < nmine;
            if (has1) {
                prefetch_res(chunk_of(k + 1), xrB);
                __syncwarp();
                tmem_ld_x32(taddr + chunk_of(k + 1) * 32, rbB);
            }
            process(chunk_of(k), rbA, xrA);
            if (has1) {
                tmem_wait_ld();
                if (has2) {
                    prefetch_res(chunk_of(k + 2), xrA);
                    __syncwarp();
                    tmem_ld_x32(taddr + chunk_of(k + 2) * 32, rbA);
                }
                process(chunk_of(k + 1), rbB, xrB);
                if (has2) tmem_wait_ld();
            }
        }
        }   // hh
        if (dbg && threadIdx.x == 64) atomicMax(reinterpret_cast<unsigned long long*>(dbg + 4), static_cast<unsigned long long>(gtimer()));
        if (args.tma_store) {
            // the mainloop is over (accum_bar): the operand ring is free and doubles as the staging tile
            fence_proxy_async_smem();
            named_bar_sync(1, kGemmEpiThreads);
            if (warp == 2 && elect_one_sync()) {
                if (!args.swap) {
                    for (int hh = 0; hh < NH; ++hh)
                        for (int cb = 0; cb < BN / 64; ++cb)
                            if (b0 + cb * 64 < args.Nb && a0 + hh * kGemmBM < args.Ma)
                                tma_store_2d(&tmap_c, smem_a + hh * half_stage_bytes + cb * (kGemmBM * 128), b0 + cb * 64,
                                             a0 + hh * kGemmBM);
                } else {
                    for (int cb = 0; cb < 2; ++cb)
                        if (a0 + cb * 64 < args.Ma) tma_store_2d(&tmap_c, smem_a + cb * (BN * 128), a0 + cb * 64, b0);
                }
                tma_store_commit();
                tma_store_wait_all();
            }
        }
        tc_fence_before();
    }
    if (dbg && threadIdx.x == 64) atomicMax(reinterpret_cast<unsigned long long*>(dbg + 5), static_cast<unsigned long long>(gtimer()));
    if (CS > 1) cluster_sync_all();   // no CTA may exit while peers can still multicast into it / arrive on its barriers
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
    if (dbg && threadIdx.x == 32) atomicMax(reinterpret_cast<unsigned long long*>(dbg + 6), static_cast<unsigned long long>(gtimer()));
}

}  // namespace smb
