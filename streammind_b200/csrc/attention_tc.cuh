// Vision-tower attention on the 5th-generation tensor cores (d = 64, non-causal, S = 577): one CTA = 128 query rows of
// one head, S = Q K^T and O += P V as tcgen05.mma with the accumulators in TMEM, Q / K / V tiles by TMA straight out of
// the packed qkv activation (SWIZZLE_128B boxes of 128 rows x 64 columns), softmax by 128 threads that each own one
// query row (TMEM lane) -- no shuffles.  Per 64-key half tile g:
//   S_g = Q K_g^T (4 MMAs, N = 64)  ->  P_g = T(exp2(s c - m_ref c)) -> smem (K-major, swizzled)  ->  O += P_g V_g (4 MMAs)
// S is double-buffered in TMEM (2 x 64 columns) and P in shared memory: the tensor core computes S of half g + 1 and
// P V of half g - 1 while the softmax threads work on half g.  The exp2 rate (MUFU, 16 / clk / SM) is the theoretical
// bound; measured, the kernel sits 4-5x above it, limited by the per-half dependency chain (profiles/r01_ncu_attention_tc.md).
// Online softmax with a LAZY reference maximum: m_ref only moves when the row maximum outgrew it by more than 2^8, then
// the row's O (TMEM) and sum are rescaled by the owning thread (tcgen05.ld / st) -- a handful of times per row at most.
// V is consumed as the MN-major B operand (rows = keys = K, 64 contiguous d = N): exactly the tile TMA delivers.
// P is rounded to the model dtype before it multiplies V; the row sum adds the fp32 exponentials before rounding (the
// normaliser of an fp32 softmax, which is what the reference computes).
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = softmax.
// Shared memory 112 KB, TMEM 256 columns -> two CTAs per SM overlap each other's softmax and MMA phases.
#pragma once
#include "ptx.cuh"

namespace smb {

struct AttnTcArgs {
    void* o;                 // [rows_total, o_ss] T; head h at column h * 64
    long long o_ss;
    int S;                   // tokens per batch item (queries = keys)
    int col_q, col_k, col_v; // column of head 0 of q / k / v in the packed matrix (elements)
    float scale_log2e;
    long long* dbg;          // optional trace (sm_test_gemm_trace): clock64 stamps of one mid-grid CTA, 8 rows of 64 slots
};

constexpr int kAtcThreads = 192;
constexpr int kAtcTile = 128;                 // query rows per CTA and keys per TMA box
constexpr int kAtcTileBytes = kAtcTile * 128; // 128 rows x 64 halfs
#ifndef SMB_ATC_POLY_EVERY
#define SMB_ATC_POLY_EVERY 0x7fffffff
#endif
// every n-th key pair takes its exp2 on the FMA pipe (atc_exp2_poly2).  Measured: n = 2, 3, 4 and "none" give the same
// class time (the kernel is not MUFU-bound) and the polynomial costs ~95 extra instructions per warp and half: off.
constexpr int kAtcPolyEvery = SMB_ATC_POLY_EVERY;
// -DSMB_ATC_TRACE compiles the clock64 trace (AttnTcArgs::dbg, tools/attn_trace.py) into attention_tc_kernel
#ifdef SMB_ATC_TRACE
#define ATC_STAMP(cond, slot) do { if (dbg && (cond)) dbg[slot] = clock64(); } while (0)
#else
#define ATC_STAMP(cond, slot) do { } while (0)
#endif
constexpr float kAtcLazy = 8.f;               // rescale O only when the row maximum grew by more than 2^8 (log2 domain)
inline int attn_tc_smem_bytes() { return 7 * kAtcTileBytes + 256; }   // Q, 2 K, 2 V, 2 P buffers + barriers (base 1024-aligned)

// K-major SW128 descriptor = umma_desc_sw128_kmajor; MN-major SW128 (B = V tile [keys x 64 d]): same fields, the 8-row
// (8 K values) groups are 1024 B apart (SBO); LBO (stride between 64-wide N groups) is unused for N = 64.
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_f16_bmn(int m, int n, bool bf16) {   // B operand MN-major (bit 16)
    return umma_idesc_f16(m, n, bf16) | (1u << 16);
}
__device__ __forceinline__ float atc_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// packed fp32 pairs (FFMA2 / FADD2 / FMUL2): half the issue slots of the exponent arguments and of the row sums
__device__ __forceinline__ uint64_t atc_pack(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void atc_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t atc_fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t atc_add2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// 2^x on the FMA pipe for a share of the elements (the MUFU unit, 16 exp2 / clk / SM, is this kernel's bound):
// x = floor(x) + fr, 2^fr by a degree-4 polynomial (max relative error 2.7e-6 on [0, 1), fitted for this kernel), the
// integer part added into the exponent field.  x is clamped at -126 (the result then rounds to 0 in T).
__device__ __forceinline__ void atc_exp2_poly2(float x0, float x1, float& p0, float& p1) {
    const uint64_t x = atc_pack(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
    uint64_t t;                                       // 1.5 * 2^23 + floor(x): the low mantissa bits hold floor(x)
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(t) : "l"(x), "l"(atc_pack(12582912.f, 12582912.f)));
    const uint64_t j = atc_add2(t, atc_pack(-12582912.f, -12582912.f));
    const uint64_t fr = atc_fma2(j, atc_pack(-1.f, -1.f), x);
    uint64_t p = atc_fma2(fr, atc_pack(0x1.bb7cd4p-7f, 0x1.bb7cd4p-7f), atc_pack(0x1.aa13f0p-5f, 0x1.aa13f0p-5f));
    p = atc_fma2(p, fr, atc_pack(0x1.ee798ap-3f, 0x1.ee798ap-3f));
    p = atc_fma2(p, fr, atc_pack(0x1.62d166p-1f, 0x1.62d166p-1f));
    p = atc_fma2(p, fr, atc_pack(0x1.00002cp+0f, 0x1.00002cp+0f));
    float t0, t1, q0, q1;
    atc_unpack(t, t0, t1);
    atc_unpack(p, q0, q1);
    p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
    p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

template <typename T>
__global__ void __launch_bounds__(kAtcThreads, 2) attention_tc_kernel(const __grid_constant__ CUtensorMap tmap, const AttnTcArgs a) {
    extern __shared__ __align__(1024) uint8_t atc_raw[];
    uint8_t* smem = atc_raw;
    if ((smem_u32(smem) & 1023u) != 0) __trap();      // SWIZZLE_128B tiles need 1024-byte alignment; no slack is budgeted
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + kAtcTileBytes;        // 2 stages of 128 keys
    uint8_t* sV = sK + 2 * kAtcTileBytes;    // 2 stages of 128 keys
    uint8_t* sP = sV + 2 * kAtcTileBytes;    // 2 buffers of 128 rows x 64 keys (one K-major SW128 atom each)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kAtcTileBytes);
    uint64_t *q_full = bars, *k_full = bars + 1, *k_empty = bars + 3, *v_full = bars + 5, *v_empty = bars + 7,
             *s_full = bars + 9, *s_empty = bars + 11, *p_full = bars + 13, *p_empty = bars + 15, *o_full = bars + 17;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kAtcTile, h = blockIdx.y, b = blockIdx.z;
    const int row_base = b * a.S;                       // first row of this batch item in the packed matrix
    const int T_tiles = (a.S + kAtcTile - 1) / kAtcTile;   // 128-key boxes
    const int G = 2 * T_tiles;                          // 64-key halves

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
            mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 128); mbar_init(&p_full[s], 128); mbar_init(&p_empty[s], 1);
        }
        mbar_init(o_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tS = tmem_base, tO = tmem_base + 128u;    // S buffers at columns 0 and 64, O at 128
    pdl_wait();
#ifdef SMB_ATC_TRACE
    long long* dbg = (a.dbg != nullptr && blockIdx.x == min(2u, gridDim.x - 1) && blockIdx.y == gridDim.y / 2 && blockIdx.z == gridDim.z / 2)
                         ? a.dbg : nullptr;
#endif
    ATC_STAMP(threadIdx.x == 0, 7 * 64);

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one_sync()) {
            mbar_arrive_expect_tx(q_full, kAtcTileBytes);
            tma_load_2d(sQ, &tmap, q_full, a.col_q + h * 64, row_base + q0, kEvictNormal);
        }
        __syncwarp();
        for (int it = 0; it < T_tiles; ++it) {
            const int st = it & 1;
            const uint32_t par = (static_cast<uint32_t>(it >> 1) & 1u) ^ 1u;
            mbar_wait(&k_empty[st], par);
            ATC_STAMP(lane == 0, 6 * 64 + it);
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(&k_full[st], kAtcTileBytes);
                tma_load_2d(sK + st * kAtcTileBytes, &tmap, &k_full[st], a.col_k + h * 64, row_base + it * kAtcTile, kEvictLast);
            }
            __syncwarp();
            mbar_wait(&v_empty[st], par);
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(&v_full[st], kAtcTileBytes);
                tma_load_2d(sV + st * kAtcTileBytes, &tmap, &v_full[st], a.col_v + h * 64, row_base + it * kAtcTile, kEvictLast);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer: S runs one half ahead of the softmax
        const uint32_t idesc_s = umma_idesc_f16(128, 64, Cvt<T>::kBf16);
        const uint32_t idesc_o = umma_idesc_f16_bmn(128, 64, Cvt<T>::kBf16);
        auto issue_s = [&](int g) {
            const int it = g >> 1, hh = g & 1, st = it & 1, sb = g & 1;
            ATC_STAMP(lane == 0, 3 * 64 + g);
            if (hh == 0) mbar_wait(&k_full[st], static_cast<uint32_t>(it >> 1) & 1u);
            ATC_STAMP(lane == 0, 4 * 64 + g);
            mbar_wait(&s_empty[sb], (static_cast<uint32_t>(g >> 1) & 1u) ^ 1u);    // softmax has read S of half g - 2
            tc_fence_after();
            ATC_STAMP(lane == 0, 5 * 64 + g);
            if (elect_one_sync()) {
                const uint64_t qd = umma_desc_sw128_kmajor(smem_u32(sQ));
                const uint64_t kd = umma_desc_sw128_kmajor(smem_u32(sK + st * kAtcTileBytes + hh * 8192));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(tS + sb * 64, qd + 2 * k, kd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                if (hh == 1) umma_commit(&k_empty[st]);
                umma_commit(&s_full[sb]);
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        issue_s(0);
        for (int g = 0; g < G; ++g) {
            if (g + 1 < G) issue_s(g + 1);
            const int it = g >> 1, hh = g & 1, st = it & 1, pb = g & 1;
            ATC_STAMP(lane == 0, 7 * 64 + 1 + g);
            mbar_wait(&p_full[pb], static_cast<uint32_t>(g >> 1) & 1u);     // P of half g written, O rescaled if it had to be
            if (hh == 0) mbar_wait(&v_full[st], static_cast<uint32_t>(it >> 1) & 1u);
            tc_fence_after();
            ATC_STAMP(lane == 0, 7 * 64 + 24 + g);
            if (elect_one_sync()) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {    // 16 keys per MMA: P buffer pb (K-major), V rows 64 hh + 16 t .. (MN-major)
                    const uint64_t pd = umma_desc_sw128_kmajor(smem_u32(sP + pb * kAtcTileBytes)) + 2 * t;
                    const uint64_t vd = umma_desc_sw128_mnmajor(smem_u32(sV + st * kAtcTileBytes + hh * 8192 + t * 2048));
                    umma_f16(tO, pd, vd, idesc_o, (g | t) != 0 ? 1u : 0u);
                }
                umma_commit(&p_empty[pb]);
                if (hh == 1) umma_commit(&v_empty[st]);
                if (g == G - 1) umma_commit(o_full);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ softmax: thread = query row = TMEM lane
        const int lane_base = (warp & 3) * 32;          // TMEM lane quarter this warp may access
        const int r = lane_base + lane;                 // row inside the tile
        const uint32_t lane_addr = static_cast<uint32_t>(lane_base) << 16;
        const float c = a.scale_log2e;
        float m_ref = -INFINITY, nmc = 0.f;             // m_ref: the (possibly stale, never too large) maximum P is taken against
        uint64_t l2 = atc_pack(0.f, 0.f);               // row sum of the exponentials (fp32, before rounding P), even / odd keys
        const uint64_t c2 = atc_pack(c, c);
        for (int g = 0; g < G; ++g) {
            const int sb = g & 1, key0 = g * 64;
            const bool partial = key0 + 64 > a.S;
            ATC_STAMP(threadIdx.x == 64, g);
            mbar_wait(&s_full[sb], static_cast<uint32_t>(g >> 1) & 1u);
            tc_fence_after();
            ATC_STAMP(threadIdx.x == 64, 64 + g);
            uint32_t ra[32], rb[32];
            tmem_ld_x32(tS + sb * 64 + lane_addr, ra);
            tmem_ld_x32(tS + sb * 64 + lane_addr + 32, rb);
            tmem_wait_ld();
            float mh4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};     // four chains: the scan is latency-, not issue-bound
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if (!partial || key0 + i < a.S) mh4[i & 3] = fmaxf(mh4[i & 3], __uint_as_float(ra[i]));
                if (!partial || key0 + 32 + i < a.S) mh4[i & 3] = fmaxf(mh4[i & 3], __uint_as_float(rb[i]));
            }
            const float mh = fmaxf(fmaxf(mh4[0], mh4[1]), fmaxf(mh4[2], mh4[3]));
            // Lazy online softmax: the reference maximum moves only when the row maximum outgrew it by more than 2^kAtcLazy,
            // so P stays <= 2^kAtcLazy (exact in T and in the fp32 sums) and O is rescaled a few times per row at most.
            float f = 1.f;
            bool resc = false;
            if (m_ref == -INFINITY) {
                if (mh != -INFINITY) { m_ref = mh; nmc = -mh * c; }     // first valid keys (half 0): nothing accumulated yet
            } else if ((mh - m_ref) * c > kAtcLazy) {
                f = atc_ex2((m_ref - mh) * c);
                m_ref = mh; nmc = -mh * c;
                float l_lo, l_hi;
                atc_unpack(l2, l_lo, l_hi);
                l2 = atc_pack(l_lo * f, l_hi * f);
                resc = true;
            }
            if (__any_sync(0xffffffffu, resc)) {         // warp-uniform; g >= 1 here
                mbar_wait(&p_empty[(g - 1) & 1], static_cast<uint32_t>((g - 1) >> 1) & 1u);   // P V of half g - 1 has landed in O
                tc_fence_after();
#pragma unroll 1
                for (int oc = 0; oc < 2; ++oc) {
                    uint32_t ob[32];
                    tmem_ld_x32(tO + lane_addr + oc * 32, ob);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) ob[i] = __float_as_uint(__uint_as_float(ob[i]) * f);
                    tmem_st_x32(tO + lane_addr + oc * 32, ob);
                }
                tmem_wait_st();
            }
            mbar_wait(&p_empty[sb], (static_cast<uint32_t>(g >> 1) & 1u) ^ 1u);   // P buffer free: P V of half g - 2 done
            const uint32_t dst = smem_u32(sP) + sb * kAtcTileBytes + r * 128;
            const uint64_t nmc2 = atc_pack(nmc, nmc);
#pragma unroll
            for (int ci = 0; ci < 2; ++ci) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float s0 = __uint_as_float(ci == 0 ? ra[2 * i] : rb[2 * i]);
                    const float s1 = __uint_as_float(ci == 0 ? ra[2 * i + 1] : rb[2 * i + 1]);
                    float x0, x1;
                    atc_unpack(atc_fma2(atc_pack(s0, s1), c2, nmc2), x0, x1);
                    float p0, p1;
                    if ((i % kAtcPolyEvery) == kAtcPolyEvery - 1) atc_exp2_poly2(x0, x1, p0, p1);
                    else { p0 = atc_ex2(x0); p1 = atc_ex2(x1); }
                    if (partial) {
                        if (key0 + ci * 32 + 2 * i >= a.S) p0 = 0.f;
                        if (key0 + ci * 32 + 2 * i + 1 >= a.S) p1 = 0.f;
                    }
                    pk[i] = Cvt<T>::pack2(p0, p1);
                    l2 = atc_add2(l2, atc_pack(p0, p1));
                }
                // columns ci*32 .. +31 of this row: 16-byte chunks ci * 4 .. + 3 of P buffer sb, swizzled by row
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int chunk = ci * 4 + q;
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((chunk ^ (r & 7)) << 4)),
                                 "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();      // P (generic-proxy stores) -> tensor-core reads (async proxy)
            mbar_arrive(&p_full[sb]);
            mbar_arrive(&s_empty[sb]);
            ATC_STAMP(threadIdx.x == 64, 2 * 64 + g);
        }
        // ---- O / l -> global
        mbar_wait(o_full, 0);
        tc_fence_after();
        const int qrow = q0 + r;
        float l_lo, l_hi;
        atc_unpack(l2, l_lo, l_hi);
        const float l = l_lo + l_hi;
        const float inv = l > 0.f ? 1.0f / l : 0.f;
        T* orow = reinterpret_cast<T*>(a.o) + static_cast<long long>(row_base + qrow) * a.o_ss + h * 64;
#pragma unroll 1
        for (int ci = 0; ci < 2; ++ci) {
            uint32_t rb[32];
            tmem_ld_x32(tO + lane_addr + ci * 32, rb);
            tmem_wait_ld();
            if (qrow < a.S) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 o;
                    o.x = Cvt<T>::pack2(__uint_as_float(rb[8 * q + 0]) * inv, __uint_as_float(rb[8 * q + 1]) * inv);
                    o.y = Cvt<T>::pack2(__uint_as_float(rb[8 * q + 2]) * inv, __uint_as_float(rb[8 * q + 3]) * inv);
                    o.z = Cvt<T>::pack2(__uint_as_float(rb[8 * q + 4]) * inv, __uint_as_float(rb[8 * q + 5]) * inv);
                    o.w = Cvt<T>::pack2(__uint_as_float(rb[8 * q + 6]) * inv, __uint_as_float(rb[8 * q + 7]) * inv);
                    *reinterpret_cast<uint4*>(orow + ci * 32 + q * 8) = o;
                }
            }
        }
        tc_fence_before();
        ATC_STAMP(threadIdx.x == 64, 7 * 64 + 60);
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace smb
