// Vision-tower attention on the 5th-generation tensor cores (d = 64, non-causal, S = 577): one CTA = 128 query rows of
// one head, S = Q K^T and O += P V as tcgen05.mma with the accumulators in TMEM, Q / K / V tiles by TMA straight out of
// the packed qkv activation (SWIZZLE_128B boxes of 128 rows x 64 columns), softmax by 128 threads that each own one
// query row (TMEM lane) -- no shuffles, no online rescaling:
//   pass A: for every 64-key half tile  S = Q K^T  ->  row maximum
//   pass B: for every half tile         S = Q K^T  ->  P = T(exp2(s c - m c))  -> smem (K-major, swizzled)  ->  O += P V
// S is double-buffered in TMEM (2 x 64 columns) and P in shared memory, so the tensor core computes S of half g + 1 and
// P V of half g - 1 while the softmax threads work on half g; the exp2 (MUFU, 16 / clk / SM) is the bound.
// Recomputing S in pass B costs 20 MMAs per CTA; it removes the TMEM round trips an online rescale of O would need.
// V is consumed as the MN-major B operand (rows = keys = K, 64 contiguous d = N): exactly the tile TMA delivers.
// P is rounded to the model dtype before it multiplies V and the row sum adds the ROUNDED values, as in attention.cuh.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = softmax.
// Shared memory 97 KB, TMEM 256 columns -> two CTAs per SM overlap each other's softmax and MMA phases.
#pragma once
#include "ptx.cuh"

namespace smb {

struct AttnTcArgs {
    void* o;                 // [rows_total, o_ss] T; head h at column h * 64
    long long o_ss;
    int S;                   // tokens per batch item (queries = keys)
    int col_q, col_k, col_v; // column of head 0 of q / k / v in the packed matrix (elements)
    float scale_log2e;
};

constexpr int kAtcThreads = 192;
constexpr int kAtcTile = 128;                 // query rows per CTA and keys per tile
constexpr int kAtcTileBytes = kAtcTile * 128; // 128 rows x 64 halfs
inline int attn_tc_smem_bytes() { return 6 * kAtcTileBytes + 1024 + 256; }   // Q, 2 K, V, 2 P atoms + align + barriers

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

template <typename T>
__global__ void __launch_bounds__(kAtcThreads, 2) attention_tc_kernel(const __grid_constant__ CUtensorMap tmap, const AttnTcArgs a) {
    extern __shared__ uint8_t atc_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(atc_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + kAtcTileBytes;        // 2 stages of 128 keys
    uint8_t* sV = sK + 2 * kAtcTileBytes;    // 1 stage: V_j is requested when P V_{j-1} has completed, under the softmax of tile j
    uint8_t* sP = sV + kAtcTileBytes;        // 2 buffers of 128 rows x 64 keys (one K-major SW128 atom each)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kAtcTileBytes);
    uint64_t *q_full = bars, *k_full = bars + 1, *k_empty = bars + 3, *v_full = bars + 5, *v_empty = bars + 6,
             *s_full = bars + 7, *s_empty = bars + 9, *p_full = bars + 11, *p_empty = bars + 13, *o_full = bars + 15;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kAtcTile, h = blockIdx.y, b = blockIdx.z;
    const int row_base = b * a.S;                       // first row of this batch item in the packed matrix
    const int T_tiles = (a.S + kAtcTile - 1) / kAtcTile;
    const int n_it = 2 * T_tiles;                       // 128-key boxes over both passes
    const int GA = 2 * T_tiles, G = 2 * GA;             // 64-key halves: pass A = [0, GA), pass B = [GA, G)

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
            mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 128); mbar_init(&p_full[s], 128); mbar_init(&p_empty[s], 1);
        }
        mbar_init(v_full, 1); mbar_init(v_empty, 1); mbar_init(o_full, 1);
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

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one_sync()) {
            mbar_arrive_expect_tx(q_full, kAtcTileBytes);
            tma_load_2d(sQ, &tmap, q_full, a.col_q + h * 64, row_base + q0, kEvictNormal);
        }
        __syncwarp();
        for (int it = 0; it < n_it; ++it) {
            const int j = it % T_tiles, st = it & 1;
            mbar_wait(&k_empty[st], (static_cast<uint32_t>(it >> 1) & 1u) ^ 1u);
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(&k_full[st], kAtcTileBytes);
                tma_load_2d(sK + st * kAtcTileBytes, &tmap, &k_full[st], a.col_k + h * 64, row_base + j * kAtcTile, kEvictLast);
            }
            __syncwarp();
            if (it >= T_tiles) {
                const int jj = it - T_tiles;
                mbar_wait(v_empty, (static_cast<uint32_t>(jj) & 1u) ^ 1u);
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(v_full, kAtcTileBytes);
                    tma_load_2d(sV, &tmap, v_full, a.col_v + h * 64, row_base + jj * kAtcTile, kEvictLast);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer: S runs one half ahead of the softmax
        const uint32_t idesc_s = umma_idesc_f16(128, 64, Cvt<T>::kBf16);
        const uint32_t idesc_o = umma_idesc_f16_bmn(128, 64, Cvt<T>::kBf16);
        auto issue_s = [&](int g) {
            const int it = g >> 1, hh = g & 1, st = it & 1, sb = g & 1;
            if (hh == 0) mbar_wait(&k_full[st], static_cast<uint32_t>(it >> 1) & 1u);
            mbar_wait(&s_empty[sb], (static_cast<uint32_t>(g >> 1) & 1u) ^ 1u);    // softmax has read S of half g - 2
            tc_fence_after();
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
            if (g >= GA) {
                const int gb = g - GA, jj = gb >> 1, hh = gb & 1, pb = gb & 1;
                mbar_wait(&p_full[pb], static_cast<uint32_t>(gb >> 1) & 1u);
                if (hh == 0) mbar_wait(v_full, static_cast<uint32_t>(jj) & 1u);
                tc_fence_after();
                if (elect_one_sync()) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) {    // 16 keys per MMA: P buffer pb (K-major), V rows 64 hh + 16 t .. (MN-major)
                        const uint64_t pd = umma_desc_sw128_kmajor(smem_u32(sP + pb * kAtcTileBytes)) + 2 * t;
                        const uint64_t vd = umma_desc_sw128_mnmajor(smem_u32(sV + hh * 8192 + t * 2048));
                        umma_f16(tO, pd, vd, idesc_o, (gb | t) != 0 ? 1u : 0u);
                    }
                    umma_commit(&p_empty[pb]);
                    if (hh == 1) umma_commit(v_empty);
                    if (gb == GA - 1) umma_commit(o_full);
                }
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax: thread = query row = TMEM lane
        const int lane_base = (warp & 3) * 32;          // TMEM lane quarter this warp may access
        const int r = lane_base + lane;                 // row inside the tile
        const uint32_t lane_addr = static_cast<uint32_t>(lane_base) << 16;
        const float c = a.scale_log2e;
        float m = -INFINITY, l = 0.f, nmc = 0.f;
        for (int g = 0; g < G; ++g) {
            const bool pass_b = g >= GA;
            const int gb = g - GA, sb = g & 1;
            const int key0 = (pass_b ? gb : g) * 64;
            const bool partial = key0 + 64 > a.S;
            if (g == GA) nmc = (m == -INFINITY) ? 0.f : -m * c;
            mbar_wait(&s_full[sb], static_cast<uint32_t>(g >> 1) & 1u);
            tc_fence_after();
            if (pass_b) mbar_wait(&p_empty[sb], (static_cast<uint32_t>(gb >> 1) & 1u) ^ 1u);   // P of half gb - 2 has been consumed
#pragma unroll
            for (int ci = 0; ci < 2; ++ci) {            // 32 score columns at a time
                uint32_t rb[32];
                tmem_ld_x32(tS + sb * 64 + lane_addr + ci * 32, rb);
                tmem_wait_ld();
                if (!pass_b) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float s = __uint_as_float(rb[i]);
                        if (!partial || key0 + ci * 32 + i < a.S) m = fmaxf(m, s);
                    }
                } else {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float p0 = atc_ex2(fmaf(__uint_as_float(rb[2 * i]), c, nmc));
                        float p1 = atc_ex2(fmaf(__uint_as_float(rb[2 * i + 1]), c, nmc));
                        if (partial) {
                            if (key0 + ci * 32 + 2 * i >= a.S) p0 = 0.f;
                            if (key0 + ci * 32 + 2 * i + 1 >= a.S) p1 = 0.f;
                        }
                        pk[i] = Cvt<T>::pack2(p0, p1);
                        const float2 f = Cvt<T>::unpack2(pk[i]);
                        l += f.x + f.y;
                    }
                    // columns ci*32 .. +31 of this row: 16-byte chunks ci * 4 .. + 3 of P buffer sb, swizzled by row
                    const uint32_t dst = smem_u32(sP) + sb * kAtcTileBytes + r * 128;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int chunk = ci * 4 + q;
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((chunk ^ (r & 7)) << 4)),
                                     "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
                    }
                }
            }
            tc_fence_before();
            if (pass_b) {
                fence_proxy_async_smem();      // P (generic-proxy stores) -> tensor-core reads (async proxy)
                mbar_arrive(&p_full[sb]);
            }
            mbar_arrive(&s_empty[sb]);
        }
        // ---- O / l -> global
        mbar_wait(o_full, 0);
        tc_fence_after();
        const int qrow = q0 + r;
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
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace smb
