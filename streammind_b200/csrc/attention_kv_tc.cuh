// LLM prefill attention on the 5th-generation tensor cores: causal, GQA, head_dim 128, queries of the new positions against
// the persistent KV cache (hf MistralForCausalLM SDPA as reached from videollama2_mistral.py:234-243,426-431).
//
// One CTA = one 128-row tile of ONE kv head x one split of the key range.  The 128 rows are (query head of the GQA group,
// token): TB = 128 / group consecutive tokens x group heads, so the K / V blocks of a kv head are read once for the whole
// group (the mma.sync kernel this replaces read them once per query head and ran one CTA per (64 tokens, query head): 32 CTAs
// for an 11-token dialogue suffix at any context length).  A fire prefills 11-74 positions against up to 8k cached ones, so
// the key range is split over the CTAs (split-KV): grid = (token tiles, kv heads, splits) ~ one wave of 148; every split
// writes an un-normalised partial (reference maximum, row sum, O) and a small second kernel merges the splits in fixed order.
// With one split the tile is normalised and written directly.
//
// Per 64-key half g of a 128-key block (same pipeline as attention_tc.cuh, d = 128 instead of 64):
//   S_g = Q K_g^T (8 MMAs 128 x 64 x 16)  ->  causal mask, P_g = T(exp2(s c - m_ref c)) -> smem  ->  O += P_g V_g (2 x 4 MMAs, N = 64 each)
// S double-buffered in TMEM (2 x 64 columns), O in 128 columns; Q / K / V tiles by TMA (SWIZZLE_128B boxes of 64 columns):
// Q straight out of the packed, already rotated qkv activation (one box per (head of the group, d half)), K / V straight out
// of the cache [kv head][max_ctx][128].  Lazy online softmax as in attention_tc.cuh.  P is rounded to the model dtype before
// it multiplies V; the row sum adds the fp32 exponentials.  Rows of the cache beyond the causal limit are masked (their P is
// exactly 0; the cache is zero-initialised, so no NaN can enter through 0 x V).
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = softmax (thread = row).
// Shared memory 192 KB (Q 32, K 2 x 32, V 2 x 32, P 2 x 16), TMEM 256 columns, one CTA per SM.
#pragma once
#include "attention_tc.cuh"

namespace smb {

struct AttnKvArgs {
    void* o;                 // [P, o_ss] T; query head h at column h * 128
    long long o_ss;
    int P;                   // new positions (query rows)
    int pos0;                // position of query row 0 = cached positions before this call
    int group;               // query heads per kv head (power of two <= 16)
    int Hq;
    int max_ctx;             // rows per kv head in the cache matrix
    int col_q;               // column of query head 0 in the packed activation
    int n_splits;
    float scale_log2e;
    float* ws_o;             // [n_splits][P][Hq][128] fp32 partial O   (n_splits > 1)
    float* ws_ml;            // [n_splits][P][Hq][2]   (reference maximum in the scaled log2 domain, row sum)
};

constexpr int kAkvThreads = 192;
constexpr int kAkvBlockBytes = 128 * 256;      // 128 rows x 128 halfs (two SWIZZLE_128B atoms of 16 KB: d 0..63 | d 64..127)
constexpr int kAkvAtomBytes = 128 * 128;
inline int attn_kv_smem_bytes() { return 5 * kAkvBlockBytes + 2 * kAkvAtomBytes + 256; }   // Q, 2 K, 2 V, 2 P + barriers

template <typename T>
__global__ void __launch_bounds__(kAkvThreads, 1) attention_kv_tc_kernel(const __grid_constant__ CUtensorMap tmq,
                                                                         const __grid_constant__ CUtensorMap tmk,
                                                                         const __grid_constant__ CUtensorMap tmv, const AttnKvArgs a) {
    extern __shared__ __align__(1024) uint8_t akv_raw[];
    uint8_t* smem = akv_raw;
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + kAkvBlockBytes;           // 2 stages
    uint8_t* sV = sK + 2 * kAkvBlockBytes;       // 2 stages
    uint8_t* sP = sV + 2 * kAkvBlockBytes;       // 2 buffers of 128 rows x 64 keys
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kAkvAtomBytes);
    uint64_t *q_full = bars, *k_full = bars + 1, *k_empty = bars + 3, *v_full = bars + 5, *v_empty = bars + 7,
             *s_full = bars + 9, *s_empty = bars + 11, *p_full = bars + 13, *p_empty = bars + 15, *o_full = bars + 17;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TB = 128 / a.group;                       // tokens per tile
    const int t0 = blockIdx.x * TB, hk = blockIdx.y, z = blockIdx.z;
    const int kv_hi = a.pos0 + min(a.P, t0 + TB);       // keys [0, kv_hi) can be visible to this tile (causal)
    const int nb = (kv_hi + 127) / 128;                 // 128-key blocks of the tile, shared out over the splits
    const int b0 = static_cast<int>(static_cast<long long>(z) * nb / a.n_splits);
    const int b1 = static_cast<int>(static_cast<long long>(z + 1) * nb / a.n_splits);
    const int NB = b1 - b0;                             // blocks of this CTA (may be 0)
    const int G = 2 * NB;                               // 64-key halves

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmq); tma_prefetch_desc(&tmk); tma_prefetch_desc(&tmv);
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
    const uint32_t tS = tmem_base, tO = tmem_base + 128u;    // S buffers at columns 0 and 64, O at 128..255
    pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (NB > 0) {
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(q_full, kAkvBlockBytes);
                for (int g = 0; g < a.group; ++g)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        tma_load_2d(sQ + j * kAkvAtomBytes + g * TB * 128, &tmq, q_full, a.col_q + (hk * a.group + g) * 128 + j * 64, t0, kEvictNormal);
            }
            __syncwarp();
            for (int it = 0; it < NB; ++it) {
                const int st = it & 1;
                const uint32_t par = (static_cast<uint32_t>(it >> 1) & 1u) ^ 1u;
                const int row = hk * a.max_ctx + (b0 + it) * 128;
                mbar_wait(&k_empty[st], par);
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(&k_full[st], kAkvBlockBytes);
#pragma unroll
                    for (int j = 0; j < 2; ++j) tma_load_2d(sK + st * kAkvBlockBytes + j * kAkvAtomBytes, &tmk, &k_full[st], j * 64, row, kEvictNormal);
                }
                __syncwarp();
                mbar_wait(&v_empty[st], par);
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(&v_full[st], kAkvBlockBytes);
#pragma unroll
                    for (int j = 0; j < 2; ++j) tma_load_2d(sV + st * kAkvBlockBytes + j * kAkvAtomBytes, &tmv, &v_full[st], j * 64, row, kEvictNormal);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer: S runs one half ahead of the softmax
        if (NB > 0) {
            const uint32_t idesc_s = umma_idesc_f16(128, 64, Cvt<T>::kBf16);
            const uint32_t idesc_o = umma_idesc_f16_bmn(128, 64, Cvt<T>::kBf16);
            auto issue_s = [&](int g) {
                const int it = g >> 1, hh = g & 1, st = it & 1, sb = g & 1;
                if (hh == 0) mbar_wait(&k_full[st], static_cast<uint32_t>(it >> 1) & 1u);
                mbar_wait(&s_empty[sb], (static_cast<uint32_t>(g >> 1) & 1u) ^ 1u);    // softmax has read S of half g - 2
                tc_fence_after();
                if (elect_one_sync()) {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {        // d = 128: two atoms of four 16-wide slabs
                        const uint64_t qd = umma_desc_sw128_kmajor(smem_u32(sQ + (kk >> 2) * kAkvAtomBytes)) + 2 * (kk & 3);
                        const uint64_t kd = umma_desc_sw128_kmajor(smem_u32(sK + st * kAkvBlockBytes + (kk >> 2) * kAkvAtomBytes + hh * 8192)) + 2 * (kk & 3);
                        umma_f16(tS + sb * 64, qd, kd, idesc_s, kk != 0 ? 1u : 0u);
                    }
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
                mbar_wait(&p_full[pb], static_cast<uint32_t>(g >> 1) & 1u);     // P of half g written, O rescaled if it had to be
                if (hh == 0) mbar_wait(&v_full[st], static_cast<uint32_t>(it >> 1) & 1u);
                tc_fence_after();
                if (elect_one_sync()) {
#pragma unroll
                    for (int j = 0; j < 2; ++j)              // output d halves: O columns 64 j .. 64 j + 63
#pragma unroll
                        for (int t = 0; t < 4; ++t) {        // 16 keys per MMA: P buffer pb (K-major), V rows 64 hh + 16 t .. (MN-major)
                            const uint64_t pd = umma_desc_sw128_kmajor(smem_u32(sP + pb * kAkvAtomBytes)) + 2 * t;
                            const uint64_t vd = umma_desc_sw128_mnmajor(smem_u32(sV + st * kAkvBlockBytes + j * kAkvAtomBytes + hh * 8192 + t * 2048));
                            umma_f16(tO + j * 64, pd, vd, idesc_o, (g | t) != 0 ? 1u : 0u);
                        }
                    umma_commit(&p_empty[pb]);
                    if (hh == 1) umma_commit(&v_empty[st]);
                    if (g == G - 1) umma_commit(o_full);
                }
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax: thread = tile row = TMEM lane
        const int lane_base = (warp & 3) * 32;          // TMEM lane quarter this warp may access
        const int r = lane_base + lane;                 // row inside the tile: head r / TB of the group, token t0 + r % TB
        const uint32_t lane_addr = static_cast<uint32_t>(lane_base) << 16;
        const int tok = t0 + (r % TB), head = hk * a.group + r / TB;
        const bool row_ok = tok < a.P;
        const int qpos = a.pos0 + tok;
        const float c = a.scale_log2e;
        float m_ref = -INFINITY, nmc = 0.f;
        uint64_t l2 = atc_pack(0.f, 0.f);
        const uint64_t c2 = atc_pack(c, c);
        for (int g = 0; g < G; ++g) {
            const int sb = g & 1, key0 = b0 * 128 + g * 64;
            const int lim = row_ok ? qpos - key0 : -1;          // keys key0 + i with i <= lim are visible to this row
            const bool partial = lim < 63;
            mbar_wait(&s_full[sb], static_cast<uint32_t>(g >> 1) & 1u);
            tc_fence_after();
            uint32_t ra[32], rb[32];
            tmem_ld_x32(tS + sb * 64 + lane_addr, ra);
            tmem_ld_x32(tS + sb * 64 + lane_addr + 32, rb);
            tmem_wait_ld();
            float mh4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if (!partial || i <= lim) mh4[i & 3] = fmaxf(mh4[i & 3], __uint_as_float(ra[i]));
                if (!partial || 32 + i <= lim) mh4[i & 3] = fmaxf(mh4[i & 3], __uint_as_float(rb[i]));
            }
            const float mh = fmaxf(fmaxf(mh4[0], mh4[1]), fmaxf(mh4[2], mh4[3]));
            float f = 1.f;
            bool resc = false;
            if (m_ref == -INFINITY) {
                if (mh != -INFINITY) { m_ref = mh; nmc = -mh * c; }
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
                for (int oc = 0; oc < 4; ++oc) {
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
            const uint32_t dst = smem_u32(sP) + sb * kAkvAtomBytes + r * 128;
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
                    float p0 = atc_ex2(x0), p1 = atc_ex2(x1);
                    if (partial) {
                        if (ci * 32 + 2 * i > lim) p0 = 0.f;
                        if (ci * 32 + 2 * i + 1 > lim) p1 = 0.f;
                    }
                    pk[i] = Cvt<T>::pack2(p0, p1);
                    l2 = atc_add2(l2, atc_pack(p0, p1));
                }
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
        }
        // ---- O / l -> global (one split: normalised, model dtype; several: fp32 partial for the merge kernel)
        if (NB > 0) {
            mbar_wait(o_full, 0);
            tc_fence_after();
        }
        float l_lo, l_hi;
        atc_unpack(l2, l_lo, l_hi);
        const float l = l_lo + l_hi;
        if (a.n_splits == 1) {
            const float inv = l > 0.f ? 1.0f / l : 0.f;
            T* orow = reinterpret_cast<T*>(a.o) + static_cast<long long>(tok) * a.o_ss + head * 128;
#pragma unroll 1
            for (int ci = 0; ci < 4; ++ci) {
                uint32_t ob[32];
                if (NB > 0) {
                    tmem_ld_x32(tO + lane_addr + ci * 32, ob);
                    tmem_wait_ld();
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) ob[i] = 0u;
                }
                if (row_ok) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 o;
                        o.x = Cvt<T>::pack2(__uint_as_float(ob[8 * q + 0]) * inv, __uint_as_float(ob[8 * q + 1]) * inv);
                        o.y = Cvt<T>::pack2(__uint_as_float(ob[8 * q + 2]) * inv, __uint_as_float(ob[8 * q + 3]) * inv);
                        o.z = Cvt<T>::pack2(__uint_as_float(ob[8 * q + 4]) * inv, __uint_as_float(ob[8 * q + 5]) * inv);
                        o.w = Cvt<T>::pack2(__uint_as_float(ob[8 * q + 6]) * inv, __uint_as_float(ob[8 * q + 7]) * inv);
                        *reinterpret_cast<uint4*>(orow + ci * 32 + q * 8) = o;
                    }
                }
            }
        } else {
            const long long prow = (static_cast<long long>(z) * a.P + tok) * a.Hq + head;
            if (row_ok) {
                a.ws_ml[prow * 2] = l > 0.f ? m_ref * c : -INFINITY;
                a.ws_ml[prow * 2 + 1] = l;
            }
            if (NB > 0) {
                float4* orow = reinterpret_cast<float4*>(a.ws_o + prow * 128);
#pragma unroll 1
                for (int ci = 0; ci < 4; ++ci) {
                    uint32_t ob[32];
                    tmem_ld_x32(tO + lane_addr + ci * 32, ob);
                    tmem_wait_ld();
                    if (row_ok) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            orow[ci * 8 + q] = make_float4(__uint_as_float(ob[4 * q]), __uint_as_float(ob[4 * q + 1]), __uint_as_float(ob[4 * q + 2]),
                                                           __uint_as_float(ob[4 * q + 3]));
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// Merge of the key-range splits: one CTA per (token, query head); warp w folds the splits z = w, w + 4, ... (lane = 4 output
// dims, 16-byte loads, the splits of a warp independent of each other), then the four warps are combined in fixed order.
template <typename T>
__global__ void __launch_bounds__(128) attention_kv_merge_kernel(const AttnKvArgs a) {
    __shared__ float sm_m[4], sm_l[4];
    __shared__ float4 sm_o[4][32];
    pdl_trigger();
    pdl_wait();
    const int tok = blockIdx.x, head = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float m = -INFINITY, l = 0.f;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = warp; z < a.n_splits; z += 4) {
        const long long prow = (static_cast<long long>(z) * a.P + tok) * a.Hq + head;
        const float2 ml = *reinterpret_cast<const float2*>(a.ws_ml + prow * 2);
        if (ml.x == -INFINITY) continue;             // the split saw no visible key of this row (its O was not written)
        const float4 oz = reinterpret_cast<const float4*>(a.ws_o + prow * 128)[lane];
        const float mn = fmaxf(m, ml.x);
        const float c0 = atc_ex2(m - mn), c1 = atc_ex2(ml.x - mn);       // c0 = 0 while m = -inf
        l = l * c0 + ml.y * c1;
        o.x = o.x * c0 + oz.x * c1; o.y = o.y * c0 + oz.y * c1; o.z = o.z * c0 + oz.z * c1; o.w = o.w * c0 + oz.w * c1;
        m = mn;
    }
    if (lane == 0) { sm_m[warp] = m; sm_l[warp] = l; }
    sm_o[warp][lane] = o;
    __syncthreads();
    if (warp == 0) {
        float M = fmaxf(fmaxf(sm_m[0], sm_m[1]), fmaxf(sm_m[2], sm_m[3]));
        float L = 0.f;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            if (sm_m[w] == -INFINITY) continue;
            const float c = atc_ex2(sm_m[w] - M);
            const float4 ow = sm_o[w][lane];
            L += sm_l[w] * c;
            r.x += ow.x * c; r.y += ow.y * c; r.z += ow.z * c; r.w += ow.w * c;
        }
        const float inv = L > 0.f ? 1.0f / L : 0.f;
        uint2 pk;
        pk.x = Cvt<T>::pack2(r.x * inv, r.y * inv);
        pk.y = Cvt<T>::pack2(r.z * inv, r.w * inv);
        *reinterpret_cast<uint2*>(reinterpret_cast<T*>(a.o) + static_cast<long long>(tok) * a.o_ss + head * 128 + lane * 4) = pk;
    }
}

}  // namespace smb
