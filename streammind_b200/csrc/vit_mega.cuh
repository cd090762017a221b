// Persistent vision-tower kernel: the whole CLIP-ViT forward of a chunk of frames (patch gather, patch-embed
// GEMM, embedding + LayerNorms, 23 x [QKV GEMM, attention, out-proj, LN, fc1 + quick_gelu, fc2, LN], patch
// mean-pool) runs as ONE launch of one CTA per SM that walks a device-resident op list.  Ops are separated
// by a grid barrier (one release-add / acquire-poll on a global counter) instead of a kernel boundary, so
// TMEM, mbarriers and tensor-map descriptors are set up once per frame, and the weight tiles of op k+1 are
// pulled into L2 (cp.async.bulk.prefetch.L2) while op k is still running.
//
// At streaming B = 1 the tower is a latency chain of ~140 small ops (577 tokens = 5 M-tiles): with one kernel
// per op every launch pays fill + epilogue + drain + launch gap (measured 10-16 us per GEMM whose tensor time is
// 2-4 us, profiles/r01_ncu_full_summary.md); here an op boundary costs one grid barrier.
//
// Warp roles (320 threads): warp 0 = TMA producer (+ L2 weight prefetch), warp 1 = tcgen05.mma issuer,
// warps 2..9 = epilogue / row ops / attention (256 threads).  TMEM: 512 columns = two 256-column accumulator
// buffers, so the epilogue of tile i overlaps the mainloop of tile i+1 when a CTA owns several tiles (chunked B).
#pragma once
#include "attention.cuh"
#include "gemm_tc.cuh"
#include "misc_kernels.cuh"

namespace smb {

enum MegaOpType : int { MOP_GEMM = 0, MOP_SPLITK_LN = 1, MOP_EMBED_LN = 2, MOP_ATTN = 3, MOP_POOL = 4, MOP_IM2COL = 5 };

struct MegaOp {
    int type;
    // ---- MOP_GEMM: out[M, N] = epi(A[M, K] . W[N, K]^T (+ bias)); W pre-tiled ([n/128][k/64][128][64])
    int map_a, map_b;   // indices into the tensor-map array (A: box 64 x 128 rows of [M, K]; B: box 64 x 128 of the tiled view)
    int M, N, K;
    int bn;             // 128 or 256
    int split_k;        // > 1: fp32 partials (EPI_STORE_F32) at out + z * split_stride
    int epi;            // EpiMode (EPI_STORE / EPI_QUICK_GELU / EPI_STORE_F32)
    int w_kb;           // k-blocks per 128-row tile of the tiled weight
    const void* w;      // tiled weight base (for the L2 prefetch)
    const void* bias;
    void* out;
    int ldo;
    long long split_stride;
    // ---- MOP_SPLITK_LN: v = T(sum_z part[z][r][:] + rbias); x[r] = T(x[r] + v); h[r] = LN(x[r]) (ln_w == null: skip)
    const float* part;
    int nsplit;
    long long part_stride;
    const void* rbias;
    void* x;
    const void* ln_w;
    const void* ln_b;
    void* h;
    int rows, C;
    float eps;
    // ---- MOP_EMBED_LN: x = pre_LN(T(cls|patch_emb + pos)); h = LN(x)   (ln_w / ln_b = layer-0 LN1)
    const void* pemb;
    const void* cls;
    const void* pos;
    const void* pre_w;
    const void* pre_b;
    int S;
    // ---- MOP_ATTN
    AttnArgs attn;
    int heads, batch;
    // ---- MOP_POOL: pooled[f, :] = T(mean_p x[f*S + 1 + p, :])
    void* pooled;
    // ---- MOP_IM2COL
    const void* pixels;
    void* im;
    int img, patch, kpad;
};

struct MegaParams {
    const MegaOp* ops;
    int op_begin, op_end;
    const CUtensorMap* maps;
    unsigned int* sync;   // [0] = barrier arrivals of this launch, [1] = exited CTAs (the last one zeroes both)
    int single;           // 1: one stand-alone GEMM op per launch (persistent GEMM): no grid barrier, no exit bookkeeping
    long long* dbg;       // optional trace, 4 slots per op: max over CTAs of [barrier passed, work done, arrived] (globaltimer ns)
};

constexpr int kMegaRingBytes = 192 * 1024;
constexpr int kMegaEpiBar = 4;    // named barrier of the 256 epilogue threads (1..3 belong to attention_body)
inline int mega_smem_bytes() { return kMegaRingBytes + 1024 /*align*/ + 512 /*barriers*/ + 2 * 256 * 4 /*bias*/; }

__device__ __forceinline__ void grid_arrive(unsigned int* ctr) {
    asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy writes of this CTA vs. TMA reads elsewhere
    __threadfence();
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}
__device__ __forceinline__ void grid_wait(const unsigned int* ctr, unsigned int target) {
    unsigned int v;
    const long long t0 = clock64();
    while (true) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        if (v >= target) break;
        __nanosleep(20);
        if (clock64() - t0 > 4000000000ll) {   // ~2 s: a lost arrival must trap, not hang the GPU
            printf("smb: grid barrier timeout block %d thread %d have %u want %u\n", blockIdx.x, threadIdx.x, v, target);
            __trap();
        }
    }
    asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

struct MegaTile { int m, n, z, kb_begin, num_kb; };
struct MegaShape { int M, N, K, bn, split_k, w_kb; };   // register copy of the GEMM fields (ops live in global memory)
__device__ __forceinline__ MegaShape mega_shape(const MegaOp& op) {
    MegaShape g;
    g.M = op.M; g.N = op.N; g.K = op.K; g.bn = op.bn; g.split_k = op.split_k; g.w_kb = op.w_kb;
    return g;
}
__device__ __forceinline__ MegaTile mega_tile(const MegaShape& op, int t) {
    const int mt = (op.M + kGemmBM - 1) / kGemmBM, nt = op.N / op.bn;
    MegaTile r;
    r.m = t % mt; r.n = (t / mt) % nt; r.z = t / (mt * nt);
    const int total_kb = (op.K + kGemmBK - 1) / kGemmBK;
    const int kb_per = op.split_k > 1 ? (total_kb + op.split_k - 1) / op.split_k : total_kb;
    r.kb_begin = op.split_k > 1 ? r.z * kb_per : 0;
    r.num_kb = max(0, min(total_kb, r.kb_begin + kb_per) - r.kb_begin);
    return r;
}
__device__ __forceinline__ int mega_ntiles(const MegaShape& op) {
    return ((op.M + kGemmBM - 1) / kGemmBM) * (op.N / op.bn) * max(1, op.split_k);
}

// ---------------------------------------------------------------------------------------------- row ops
// one warp per row; lane owns chunks j = 0..C/256-1 of 8 contiguous elements at column j*256 + lane*8
template <typename T>
__device__ __forceinline__ void mega_splitk_ln_row(const MegaOp& op, int row, int lane) {
    const int C = op.C, nch = C >> 8;
    T* xr = reinterpret_cast<T*>(op.x) + static_cast<long long>(row) * C;
    float v[4][8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j < nch) {
            const int c0 = j * 256 + lane * 8;
            const float* p = op.part + static_cast<long long>(row) * C + c0;
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int z = 0; z < op.nsplit; ++z) {
                const float4 a = *reinterpret_cast<const float4*>(p + z * op.part_stride);
                const float4 b = *reinterpret_cast<const float4*>(p + z * op.part_stride + 4);
                acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
                acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
            }
            const uint4 bu = *reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(op.rbias) + c0);
            const uint4 xu = *reinterpret_cast<const uint4*>(xr + c0);
            const uint32_t bw[4] = {bu.x, bu.y, bu.z, bu.w}, xw[4] = {xu.x, xu.y, xu.z, xu.w};
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 bf = Cvt<T>::unpack2(bw[i]), xf = Cvt<T>::unpack2(xw[i]);
                v[j][2 * i] = rnd<T>(xf.x + rnd<T>(acc[2 * i] + bf.x));
                v[j][2 * i + 1] = rnd<T>(xf.y + rnd<T>(acc[2 * i + 1] + bf.y));
                s += v[j][2 * i] + v[j][2 * i + 1];
                o[i] = Cvt<T>::pack2(v[j][2 * i], v[j][2 * i + 1]);
            }
            *reinterpret_cast<uint4*>(xr + c0) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
    if (op.ln_w == nullptr) return;
    s = warp_sum(s);
    const float mean = s / C;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (j < nch)
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float d = v[j][i] - mean; sq += d * d; }
    sq = warp_sum(sq);
    const float r = rsqrtf(sq / C + op.eps);
    T* hr = reinterpret_cast<T*>(op.h) + static_cast<long long>(row) * C;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j < nch) {
            const int c0 = j * 256 + lane * 8;
            const uint4 wu = *reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(op.ln_w) + c0);
            const uint4 b2 = *reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(op.ln_b) + c0);
            const uint32_t ww[4] = {wu.x, wu.y, wu.z, wu.w}, b2w[4] = {b2.x, b2.y, b2.z, b2.w};
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 wf = Cvt<T>::unpack2(ww[i]), bf = Cvt<T>::unpack2(b2w[i]);
                o[i] = Cvt<T>::pack2((v[j][2 * i] - mean) * r * wf.x + bf.x, (v[j][2 * i + 1] - mean) * r * wf.y + bf.y);
            }
            *reinterpret_cast<uint4*>(hr + c0) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

template <typename T>
__device__ __forceinline__ void mega_embed_ln_row(const MegaOp& op, int row, int lane) {
    constexpr int VPL = 32;
    const int C = op.C, S = op.S;
    const int f = row / S, p = row % S;
    const T* cls = reinterpret_cast<const T*>(op.cls);
    const T* pemb = reinterpret_cast<const T*>(op.pemb);
    const T* pos = reinterpret_cast<const T*>(op.pos);
    float v[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = lane + i * 32;
        if (c < C) {
            const float e = p == 0 ? Cvt<T>::to_f(cls[c])
                                   : Cvt<T>::to_f(pemb[(static_cast<long long>(f) * (S - 1) + p - 1) * C + c]);
            v[i] = rnd<T>(e + Cvt<T>::to_f(pos[static_cast<long long>(p) * C + c]));
        } else {
            v[i] = 0.f;
        }
    }
    warp_layernorm<T, VPL>(v, C, lane, reinterpret_cast<const T*>(op.pre_w), reinterpret_cast<const T*>(op.pre_b), op.eps);
    T* x = reinterpret_cast<T*>(op.x);
    T* h = reinterpret_cast<T*>(op.h);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = lane + i * 32;
        if (c < C) x[static_cast<long long>(row) * C + c] = Cvt<T>::from_f(v[i]);
    }
    warp_layernorm<T, VPL>(v, C, lane, reinterpret_cast<const T*>(op.ln_w), reinterpret_cast<const T*>(op.ln_b), op.eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = lane + i * 32;
        if (c < C) h[static_cast<long long>(row) * C + c] = Cvt<T>::from_f(v[i]);
    }
}

// item = (frame f, 8-column chunk): lanes take patches lane, lane+32, ...; fp32 sum; T(mean)
template <typename T>
__device__ __forceinline__ void mega_pool_item(const MegaOp& op, int item, int lane) {
    const int C = op.C, S = op.S, P = S - 1, nchunk = C >> 3;
    const int f = item / nchunk, c0 = (item % nchunk) * 8;
    const T* src = reinterpret_cast<const T*>(op.x) + (static_cast<long long>(f) * S + 1) * C + c0;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int p = lane; p < P; p += 32) {
        const uint4 u = *reinterpret_cast<const uint4*>(src + static_cast<long long>(p) * C);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 fv = Cvt<T>::unpack2(w[i]);
            acc[2 * i] += fv.x; acc[2 * i + 1] += fv.y;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = warp_sum(acc[i]);
    if (lane == 0) {
        uint4 o;
        o.x = Cvt<T>::pack2(acc[0] / P, acc[1] / P); o.y = Cvt<T>::pack2(acc[2] / P, acc[3] / P);
        o.z = Cvt<T>::pack2(acc[4] / P, acc[5] / P); o.w = Cvt<T>::pack2(acc[6] / P, acc[7] / P);
        *reinterpret_cast<uint4*>(reinterpret_cast<T*>(op.pooled) + static_cast<long long>(f) * C + c0) = o;
    }
}

// ---------------------------------------------------------------------------------------------- the kernel
template <typename T>
__global__ void __launch_bounds__(kGemmThreads, 1) vit_mega_kernel(const MegaParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + kMegaRingBytes);
    uint64_t* empty_bar = full_bar + kGemmMaxStages;
    uint64_t* tfull_bar = empty_bar + kGemmMaxStages;    // [2] accumulator buffer ready
    uint64_t* tempty_bar = tfull_bar + 2;                // [2] accumulator buffer drained (256 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* bias_s = reinterpret_cast<float*>(ring + kMegaRingBytes + 512);   // [2][256]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int G = gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGemmMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], kGemmEpiThreads); }
        fence_mbar_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        // =============================================================== TMA producer
        uint32_t ebits = 0;   // per-stage parity of the next empty-barrier wait
        for (int oi = P.op_begin; oi < P.op_end; ++oi) {
            const MegaOp& gop = P.ops[oi];
            const int type = gop.type;
            const MegaShape op = mega_shape(gop);
            // pull this op's weights towards L2 before waiting for the previous op to finish everywhere
            if (type == MOP_GEMM && !P.single) {
                const char* wbase = reinterpret_cast<const char*>(gop.w);
                const int ntl = mega_ntiles(op);
                for (int t = blockIdx.x; t < ntl; t += G) {
                    const MegaTile tl = mega_tile(op, t);
                    if (tl.m != 0) continue;   // one CTA per (n, z) weight block
                    const int per_n = op.bn / kGemmBM;
                    for (int j = 0; j < per_n; ++j) {
                        const char* base = wbase +
                                           (static_cast<size_t>(tl.n * per_n + j) * op.w_kb + tl.kb_begin) * (kGemmBM * kGemmBK * 2);
                        for (int kb = lane; kb < tl.num_kb; kb += 32)
                            l2_prefetch_bulk(base + static_cast<size_t>(kb) * (kGemmBM * kGemmBK * 2), kGemmBM * kGemmBK * 2);
                    }
                }
            }
            const int p = oi - P.op_begin;
            if (p > 0) {
                if (lane == 0) grid_wait(P.sync, p * G);
                __syncwarp();
                asm volatile("fence.proxy.async;" ::: "memory");
            }
            if (type != MOP_GEMM) continue;
            const CUtensorMap* ta = P.maps + gop.map_a;
            const CUtensorMap* tb = P.maps + gop.map_b;
            const int BN = op.bn;
            constexpr int A_BYTES = kGemmBM * kGemmBK * 2;
            const int B_BYTES = BN * kGemmBK * 2;
            const int NSTAGE = min(kGemmMaxStages, kMegaRingBytes / (A_BYTES + B_BYTES));
            uint8_t* smem_a = ring;
            uint8_t* smem_b = ring + NSTAGE * A_BYTES;
            int stage = 0;
            const int nt = mega_ntiles(op);
            for (int t = blockIdx.x; t < nt; t += G) {
                const MegaTile tl = mega_tile(op, t);
                for (int kb = 0; kb < tl.num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], ((ebits >> stage) & 1u) ^ 1u);
                    ebits ^= 1u << stage;
                    if (elect_one_sync()) {
                        mbar_arrive_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
                        const int kg = tl.kb_begin + kb;
                        tma_load_2d(smem_a + stage * A_BYTES, ta, &full_bar[stage], kg * kGemmBK, tl.m * kGemmBM, kEvictLast);
                        const int per_n = BN / kGemmBM;
                        for (int j = 0; j < per_n; ++j)
                            tma_load_2d(smem_b + stage * B_BYTES + j * (kGemmBM * 128), tb, &full_bar[stage], 0,
                                        ((tl.n * per_n + j) * op.w_kb + kg) * kGemmBM, kEvictNormal);
                    }
                    __syncwarp();
                    if (++stage == NSTAGE) stage = 0;
                }
            }
        }
    } else if (warp == 1) {
        // =============================================================== MMA issuer
        uint32_t fbits = 0;      // per-stage parity of the next full-barrier wait
        uint32_t acc_cnt = 0;    // accumulator tiles issued so far (buffer = acc_cnt & 1)
        for (int oi = P.op_begin; oi < P.op_end; ++oi) {
            const MegaOp& gop = P.ops[oi];
            if (gop.type != MOP_GEMM) continue;
            const MegaShape op = mega_shape(gop);
            const int BN = op.bn;
            constexpr int A_BYTES = kGemmBM * kGemmBK * 2;
            const int B_BYTES = BN * kGemmBK * 2;
            const int NSTAGE = min(kGemmMaxStages, kMegaRingBytes / (A_BYTES + B_BYTES));
            uint8_t* smem_a = ring;
            uint8_t* smem_b = ring + NSTAGE * A_BYTES;
            const uint32_t idesc = umma_idesc_f16(kGemmBM, BN, Cvt<T>::kBf16);
            int stage = 0;
            const int nt = mega_ntiles(op);
            for (int t = blockIdx.x; t < nt; t += G) {
                const MegaTile tl = mega_tile(op, t);
                const uint32_t buf = acc_cnt & 1u;
                mbar_wait(&tempty_bar[buf], ((acc_cnt >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 256u;
                for (int kb = 0; kb < tl.num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], (fbits >> stage) & 1u);
                    fbits ^= 1u << stage;
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(smem_a + stage * A_BYTES));
                        const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(smem_b + stage * B_BYTES));
#pragma unroll
                        for (int k = 0; k < kGemmBK / 16; ++k)
                            umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_commit(&empty_bar[stage]);
                        if (kb == tl.num_kb - 1) umma_commit(&tfull_bar[buf]);
                    }
                    __syncwarp();
                    if (++stage == NSTAGE) stage = 0;
                }
                ++acc_cnt;
            }
        }
    } else {
        // =============================================================== epilogue / row ops / attention (256 threads)
        const int et = threadIdx.x - 64;            // 0..255
        const int ew = warp - 2;                    // 0..7
        const int lane_base = (warp & 3) * 32;      // TMEM lane quarter this warp may read
        const int chalf = ew >> 2;                  // 0: even 32-column chunks, 1: odd
        uint32_t acc_cnt = 0;
        for (int oi = P.op_begin; oi < P.op_end; ++oi) {
            const MegaOp& gop = P.ops[oi];
            const int type = gop.type;
            const int p = oi - P.op_begin;
            if (p > 0) {
                if (et == 0) grid_wait(P.sync, p * G);
                named_bar_sync(kMegaEpiBar, kGemmEpiThreads);
            }
            if (P.dbg && et == 0) atomicMax(reinterpret_cast<unsigned long long*>(P.dbg + 4 * oi), static_cast<unsigned long long>(gtimer()));
            if (type == MOP_GEMM) {
                const MegaShape op = mega_shape(gop);
                const int BN = op.bn;
                const int nt = mega_ntiles(op);
                const T* bias = reinterpret_cast<const T*>(gop.bias);
                const int epi = gop.epi;
                const long long split_stride = gop.split_stride;
                GemmArgs ga{};
                ga.Ma = op.M; ga.Nb = op.N; ga.K = op.K; ga.out = gop.out; ga.ldo = gop.ldo; ga.bn = BN; ga.epi = epi;
                for (int t = blockIdx.x; t < nt; t += G) {
                    const MegaTile tl = mega_tile(op, t);
                    const uint32_t buf = acc_cnt & 1u;
                    float* bs = bias_s + buf * 256;
                    const int b0 = tl.n * BN;
                    for (int c = et; c < BN; c += kGemmEpiThreads)
                        bs[c] = (bias != nullptr && op.split_k <= 1) ? Cvt<T>::to_f(bias[b0 + c]) : 0.0f;
                    named_bar_sync(kMegaEpiBar, kGemmEpiThreads);
                    mbar_wait(&tfull_bar[buf], (acc_cnt >> 1) & 1u);
                    tc_fence_after();
                    const int a_row = tl.m * kGemmBM + lane_base + lane;
                    const bool a_ok = a_row < op.M;
                    EpiCtx cx;
                    cx.args = &ga; cx.stage = nullptr; cx.bias_s = bs; cx.bv = 0.f; cx.a_row = a_row; cx.b0 = b0; cx.BN = BN;
                    cx.lane_row = lane_base + lane;
                    cx.out_off = op.split_k > 1 ? static_cast<size_t>(tl.z) * static_cast<size_t>(split_stride) : 0;
                    const uint32_t taddr = tmem_base + buf * 256u + (static_cast<uint32_t>(lane_base) << 16);
                    const int nchunk = BN / 32;
                    const int nmine = nchunk > chalf ? (nchunk - chalf + 1) / 2 : 0;
                    uint32_t rbA[32], rbB[32];
                    const T* out_t = reinterpret_cast<const T*>(ga.out);
                    auto process = [&](int c, const uint32_t (&rb)[32]) {
                        if (!a_ok) return;
                        uint4 xr[4] = {};
                        if (epi == EPI_RESIDUAL) {   // in-place residual stream: this thread's 32 columns of its row
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (b0 + c * 32 + q * 8 < op.N && c * 32 + q * 8 < BN)
                                    xr[q] = *reinterpret_cast<const uint4*>(out_t + static_cast<size_t>(a_row) * ga.ldo + b0 + c * 32 + q * 8);
                        }
                        switch (epi) {
                            case EPI_STORE: epi_chunk<T, false, EPI_STORE, false>(cx, c * 32, rb, xr); break;
                            case EPI_QUICK_GELU: epi_chunk<T, false, EPI_QUICK_GELU, false>(cx, c * 32, rb, xr); break;
                            case EPI_RESIDUAL: epi_chunk<T, false, EPI_RESIDUAL, false>(cx, c * 32, rb, xr); break;
                            default: epi_chunk<T, false, EPI_STORE_F32, false>(cx, c * 32, rb, xr); break;
                        }
                    };
                    if (nmine > 0) { tmem_ld_x32(taddr + chalf * 32, rbA); tmem_wait_ld(); }
#pragma unroll 1
                    for (int k = 0; k < nmine; k += 2) {
                        const bool has1 = k + 1 < nmine, has2 = k + 2 < nmine;
                        if (has1) tmem_ld_x32(taddr + (chalf + 2 * (k + 1)) * 32, rbB);
                        process(chalf + 2 * k, rbA);
                        if (has1) {
                            tmem_wait_ld();
                            if (has2) tmem_ld_x32(taddr + (chalf + 2 * (k + 2)) * 32, rbA);
                            process(chalf + 2 * (k + 1), rbB);
                            if (has2) tmem_wait_ld();
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(&tempty_bar[buf]);
                    ++acc_cnt;
                }
            } else if (type == MOP_SPLITK_LN) {
                const MegaOp& op = gop;
                const int rows = op.rows;
                for (int row = ew * G + blockIdx.x; row < rows; row += 8 * G) mega_splitk_ln_row<T>(op, row, lane);
            } else if (type == MOP_EMBED_LN) {
                const MegaOp& op = gop;
                const int rows = op.rows;
                for (int row = ew * G + blockIdx.x; row < rows; row += 8 * G) mega_embed_ln_row<T>(op, row, lane);
            } else if (type == MOP_POOL) {
                const MegaOp& op = gop;
                const int items = op.batch * (op.C >> 3);
                for (int it = ew * G + blockIdx.x; it < items; it += 8 * G) mega_pool_item<T>(op, it, lane);
            } else if (type == MOP_IM2COL) {
                const MegaOp& op = gop;
                const int gw = op.img / op.patch, PP = gw * gw, kreal = 3 * op.patch * op.patch;
                const long long total = static_cast<long long>(op.batch) * PP * op.kpad;
                const T* px = reinterpret_cast<const T*>(op.pixels);
                T* im = reinterpret_cast<T*>(op.im);
                for (long long i = static_cast<long long>(blockIdx.x) * kGemmEpiThreads + et; i < total;
                     i += static_cast<long long>(G) * kGemmEpiThreads) {
                    const int k = static_cast<int>(i % op.kpad);
                    const long long row = i / op.kpad;
                    T v = Cvt<T>::from_f(0.f);
                    if (k < kreal) {
                        const int b = static_cast<int>(row / PP), pp = static_cast<int>(row % PP);
                        const int py = pp / gw, pxx = pp % gw;
                        const int c = k / (op.patch * op.patch), rem = k % (op.patch * op.patch);
                        const int ii = rem / op.patch, jj = rem % op.patch;
                        v = px[((static_cast<long long>(b) * 3 + c) * op.img + (py * op.patch + ii)) * op.img + pxx * op.patch + jj];
                    }
                    im[i] = v;
                }
            } else if (type == MOP_ATTN) {
                const AttnArgs aa = gop.attn;
                const int heads = gop.heads;
                const int q_tiles = (aa.q_len + kAttnBQ - 1) / kAttnBQ;
                const int items = q_tiles * heads * gop.batch;
                for (int it = blockIdx.x; it < items; it += G) {
                    const int qt = it % q_tiles, hh = (it / q_tiles) % heads, bb = it / (q_tiles * heads);
                    attention_body<T, 64>(aa, qt, hh, bb, ring, et);
                    named_bar_sync(kMegaEpiBar, kGemmEpiThreads);   // smem of this item is free again
                }
            }
            if (P.single) continue;
            // ---- this CTA's part of the op is done and visible: arrive on the grid barrier
            if (P.dbg && et == 0) atomicMax(reinterpret_cast<unsigned long long*>(P.dbg + 4 * oi + 1), static_cast<unsigned long long>(gtimer()));
            asm volatile("fence.proxy.async;" ::: "memory");
            __threadfence();
            named_bar_sync(kMegaEpiBar, kGemmEpiThreads);
            if (et == 0) {
                grid_arrive(P.sync);
                if (P.dbg) atomicMax(reinterpret_cast<unsigned long long*>(P.dbg + 4 * oi + 2), static_cast<unsigned long long>(gtimer()));
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
    if (threadIdx.x == 0 && !P.single) {
        const unsigned int old = atomicAdd(P.sync + 1, 1u);
        if (old == G - 1) {   // every CTA is past its last barrier wait: clean up for the next launch
            P.sync[0] = 0;
            P.sync[1] = 0;
            __threadfence();
        }
    }
}

}  // namespace smb
