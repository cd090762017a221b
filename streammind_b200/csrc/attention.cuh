// Fused attention (QK^T -> online softmax -> PV) for (a) the ViT encoder (S = 577, d = 64, 16 heads,
// non-causal) and (b) LLM prefill against the persistent KV cache (d = 128, GQA, causal with a
// position offset).  One CTA = 64 query rows of one head; K/V stream through a 2-stage cp.async ring
// in XOR-swizzled shared memory; S and O live in registers (fp32), P is rounded to the model dtype
// before PV exactly like flash / SDPA kernels do.
//
// Round-1 implementation uses warp-level mma.sync.m16n8k16 tiles (attention is 9 % of the ViT FLOPs);
// the tcgen05/TMEM version is the planned replacement ("what comes next").
#pragma once
#include "ptx.cuh"

namespace smb {

struct AttnArgs {
    const void* q;  // element (b, s, h, :) at q + b*q_bs + s*q_ss + h*D
    const void* k;  // element (b, s, hk, :) at k + b*k_bs + hk*k_hs + s*k_ss
    const void* v;
    void* o;        // element (b, s, h, :) at o + b*o_bs + s*o_ss + h*D
    long long q_bs, q_ss, k_bs, k_hs, k_ss, v_bs, v_hs, v_ss, o_bs, o_ss;
    int q_len;      // query rows per batch item
    int kv_len;     // keys per batch item
    int q_pos0;     // causal: query row i sits at absolute position q_pos0 + i
    int causal;
    int group;      // query heads per kv head (GQA); 1 for the ViT
    float scale_log2e;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const uint32_t s = smem_u32(smem);
    const int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
template <typename T>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma16816<__half>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                                        uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float ex2_approx(float x) {   // MUFU.EX2: 2^-inf = +0, rel. error 2^-22
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

#ifndef SMB_ATTN_MINBLOCKS
#define SMB_ATTN_MINBLOCKS 2   // 128 registers: two CTAs per SM (3 -> 80 registers + spills: measured slower)
#endif
constexpr int kAttnBQ = 64;   // query rows per CTA (4 warps x 16 per key group)
constexpr int kAttnBK = 64;   // keys per pipeline stage
constexpr int kAttnGroups = 2;  // key groups per CTA: group g walks key tiles g, g+2, ... (halves the serial chain)
constexpr int kAttnGroupThreads = 128;
constexpr int kAttnThreads = kAttnGroups * kAttnGroupThreads;

template <int D> __host__ __device__ constexpr int attn_smem_bytes() {
    return (kAttnBQ + kAttnGroups * 4 * kAttnBK) * D * 2;
}
// byte offset of element (row, col) in a [rows][D] tile with 16-byte chunks XOR-swizzled by row
template <int D> __device__ __forceinline__ uint32_t swz(int row, int col) {
    const int chunk = col >> 3;
    return static_cast<uint32_t>(row * (D * 2) + ((((chunk & 7) ^ (row & 7)) | (chunk & ~7)) << 4) + ((col & 7) << 1));
}

template <typename T, int D, int NT>
__device__ __forceinline__ void attn_load_tile(uint8_t* smem_tile, const T* base, long long row_stride, int row0,
                                               int nrows_valid_total, int tid) {
    // tile = kAttnBK (== kAttnBQ) rows x D; rows >= nrows_valid_total are zero-filled; NT cooperating threads
    constexpr int CHUNKS = D / 8;
    for (int i = tid; i < kAttnBK * CHUNKS; i += NT) {
        const int r = i / CHUNKS, c = i % CHUNKS;
        const bool ok = (row0 + r) < nrows_valid_total;
        const T* src = base + static_cast<long long>(ok ? row0 + r : 0) * row_stride + c * 8;
        cp_async16(smem_tile + swz<D>(r, c * 8), src, ok);
    }
}

// Barrier ids used by the body: 1, 2 = key groups; kAttnBarAll = all kAttnThreads threads of the work item
// (bar.sync id, count -- so the body also runs on a 256-thread subset of a larger CTA: vit_mega.cuh).
constexpr int kAttnBarAll = 3;

// One work item = 64 query rows (tile qt) of head h of batch item b; tid in [0, kAttnThreads).
template <typename T, int D>
__device__ __forceinline__ void attention_body(const AttnArgs& a, int qt, int h, int b, uint8_t* attn_smem, int tid) {
    constexpr int TILE = kAttnBK * D * 2;
    const int grp = tid / kAttnGroupThreads, gtid = tid % kAttnGroupThreads;
    const int warp = gtid >> 5, lane = tid & 31;
    uint8_t* sQ = attn_smem;
    uint8_t* sK = sQ + kAttnBQ * D * 2 + grp * 4 * TILE;   // 2 stages
    uint8_t* sV = sK + 2 * TILE;                            // 2 stages
    const int q0 = qt * kAttnBQ;
    const int hk = h / a.group;
    const T* qp = reinterpret_cast<const T*>(a.q) + b * a.q_bs + h * D;
    const T* kp = reinterpret_cast<const T*>(a.k) + b * a.k_bs + hk * a.k_hs;
    const T* vp = reinterpret_cast<const T*>(a.v) + b * a.v_bs + hk * a.v_hs;

    // keys this query tile can see
    int kv_end = a.kv_len;
    if (a.causal) kv_end = min(kv_end, a.q_pos0 + min(q0 + kAttnBQ, a.q_len));
    const int ntiles = (kv_end + kAttnBK - 1) / kAttnBK;

    attn_load_tile<T, D, kAttnThreads>(sQ, qp, a.q_ss, q0, a.q_len, tid);
    if (grp < ntiles) {
        attn_load_tile<T, D, kAttnGroupThreads>(sK, kp, a.k_ss, grp * kAttnBK, kv_end, gtid);
        attn_load_tile<T, D, kAttnGroupThreads>(sV, vp, a.v_ss, grp * kAttnBK, kv_end, gtid);
    }
    cp_async_commit();
    cp_async_wait<0>();
    named_bar_sync(kAttnBarAll, kAttnThreads);

    constexpr int KD = D / 16;   // k-steps over the head dim for QK^T
    constexpr int NB = kAttnBK / 8;  // 8-wide key blocks per stage
    constexpr int OD = D / 8;    // 8-wide output column blocks
    float o_acc[OD][4];
#pragma unroll
    for (int i = 0; i < OD; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    uint32_t qf[KD][4];
    const int g = lane >> 2, t = lane & 3;
    const int qrow_a = q0 + warp * 16 + g;  // rows owned: qrow_a and qrow_a + 8
    const bool warp_active = q0 + warp * 16 < a.q_len;   // warps whose 16 rows are all padding only load and sync
#pragma unroll
    for (int kk = 0; kk < KD; ++kk) {
        const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int col = kk * 16 + (lane >> 4) * 8;
        ldmatrix_x4(qf[kk], smem_u32(sQ + swz<D>(row, col)));
    }

    int st = 0;
    for (int it = grp; it < ntiles; it += kAttnGroups) {
        if (it + kAttnGroups < ntiles) {
            attn_load_tile<T, D, kAttnGroupThreads>(sK + (st ^ 1) * TILE, kp, a.k_ss, (it + kAttnGroups) * kAttnBK, kv_end, gtid);
            attn_load_tile<T, D, kAttnGroupThreads>(sV + (st ^ 1) * TILE, vp, a.v_ss, (it + kAttnGroups) * kAttnBK, kv_end, gtid);
        }
        cp_async_commit();
        const uint8_t* tK = sK + st * TILE;
        const uint8_t* tV = sV + st * TILE;
        if (warp_active) {
        // ---- S = Q K^T  (16 x 64 per warp)
        float s[NB][4];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) { s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < KD; ++kk) {
#pragma unroll
            for (int nb2 = 0; nb2 < NB / 2; ++nb2) {
                uint32_t bf[4];
                const int row = nb2 * 16 + (lane & 7) + (lane >> 4) * 8;
                const int col = kk * 16 + ((lane >> 3) & 1) * 8;
                ldmatrix_x4(bf, smem_u32(tK + swz<D>(row, col)));
                mma16816<T>(s[2 * nb2], qf[kk], bf[0], bf[1]);
                mma16816<T>(s[2 * nb2 + 1], qf[kk], bf[2], bf[3]);
            }
        }
        // ---- mask + online softmax (rows g and g+8; 4 lanes share a row).  Scores stay unscaled: the scale is
        // folded into the exponent (one FFMA per element), and only tiles that need it are masked.
        const int kbase = it * kAttnBK;
        const bool need_mask = a.causal || (kbase + kAttnBK > kv_end);
        float mx[2] = {-INFINITY, -INFINITY};
        if (need_mask) {
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int key = kbase + nb * 8 + 2 * t + (e & 1);
                    const int qrow = qrow_a + (e >> 1) * 8;
                    bool ok = key < kv_end;
                    if (a.causal) ok = ok && (key <= a.q_pos0 + qrow);
                    const float v = ok ? s[nb][e] : -INFINITY;
                    s[nb][e] = v;
                    mx[e >> 1] = fmaxf(mx[e >> 1], v);
                }
            }
        } else {
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                mx[0] = fmaxf(mx[0], fmaxf(s[nb][0], s[nb][1]));
                mx[1] = fmaxf(mx[1], fmaxf(s[nb][2], s[nb][3]));
            }
        }
        float corr[2], nmsafe[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float mnew = fmaxf(m_run[r], mx[r] * a.scale_log2e);   // running max in scaled (log2) units
            const float msafe = (mnew == -INFINITY) ? 0.f : mnew;
            corr[r] = ex2_approx(m_run[r] - msafe);   // m_run = -inf -> 0
            nmsafe[r] = -msafe;
            m_run[r] = mnew;
        }
        float rs[2] = {0.f, 0.f};
        uint32_t pf[NB / 2][4];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
            const float p0 = ex2_approx(fmaf(s[nb][0], a.scale_log2e, nmsafe[0]));
            const float p1 = ex2_approx(fmaf(s[nb][1], a.scale_log2e, nmsafe[0]));
            const float p2 = ex2_approx(fmaf(s[nb][2], a.scale_log2e, nmsafe[1]));
            const float p3 = ex2_approx(fmaf(s[nb][3], a.scale_log2e, nmsafe[1]));
            // the row sum uses the T-rounded probabilities that actually multiply V
            const uint32_t lo = Cvt<T>::pack2(p0, p1), hi = Cvt<T>::pack2(p2, p3);
            const float2 flo = Cvt<T>::unpack2(lo), fhi = Cvt<T>::unpack2(hi);
            rs[0] += flo.x + flo.y;
            rs[1] += fhi.x + fhi.y;
            pf[nb >> 1][(nb & 1) * 2 + 0] = lo;
            pf[nb >> 1][(nb & 1) * 2 + 1] = hi;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
        if (__any_sync(0xffffffffu, corr[0] != 1.0f || corr[1] != 1.0f)) {   // the running max moved somewhere in the warp
#pragma unroll
            for (int i = 0; i < OD; ++i) {
                o_acc[i][0] *= corr[0]; o_acc[i][1] *= corr[0];
                o_acc[i][2] *= corr[1]; o_acc[i][3] *= corr[1];
            }
        }
        // ---- O += P V
#pragma unroll
        for (int kk = 0; kk < kAttnBK / 16; ++kk) {
#pragma unroll
            for (int nb2 = 0; nb2 < OD / 2; ++nb2) {
                uint32_t bf[4];
                const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int col = nb2 * 16 + (lane >> 4) * 8;
                ldmatrix_x4_trans(bf, smem_u32(tV + swz<D>(row, col)));
                mma16816<T>(o_acc[2 * nb2], pf[kk], bf[0], bf[1]);
                mma16816<T>(o_acc[2 * nb2 + 1], pf[kk], bf[2], bf[3]);
            }
        }
        }   // warp_active
        cp_async_wait<0>();                       // next tile landed (issued before this tile's math)
        named_bar_sync(1 + grp, kAttnGroupThreads);   // group-local: everyone is done with stage st
        st ^= 1;
    }
    // ---- quad-reduce the row sums, then merge the key groups (group 1 -> smem -> group 0)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    named_bar_sync(kAttnBarAll, kAttnThreads);      // all key tiles consumed: stage memory is free
    float* mb = reinterpret_cast<float*>(attn_smem + kAttnBQ * D * 2) + gtid * (OD * 4 + 4);
    if (grp == 1) {
#pragma unroll
        for (int i = 0; i < OD; ++i) {
            mb[i * 4 + 0] = o_acc[i][0]; mb[i * 4 + 1] = o_acc[i][1];
            mb[i * 4 + 2] = o_acc[i][2]; mb[i * 4 + 3] = o_acc[i][3];
        }
        mb[OD * 4 + 0] = m_run[0]; mb[OD * 4 + 1] = m_run[1];
        mb[OD * 4 + 2] = l_run[0]; mb[OD * 4 + 3] = l_run[1];
    }
    named_bar_sync(kAttnBarAll, kAttnThreads);
    if (grp == 0) {
    float sc0[2], sc1[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const float m1 = mb[OD * 4 + r], l1 = mb[OD * 4 + 2 + r];
        const float mm = fmaxf(m_run[r], m1);
        sc0[r] = m_run[r] == -INFINITY ? 0.f : ex2_approx(m_run[r] - mm);
        sc1[r] = m1 == -INFINITY ? 0.f : ex2_approx(m1 - mm);
        l_run[r] = l_run[r] * sc0[r] + l1 * sc1[r];
    }
    T* op = reinterpret_cast<T*>(a.o) + b * a.o_bs + h * D;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int qrow = qrow_a + r * 8;
        if (qrow < a.q_len) {
            const float inv = l_run[r] > 0.f ? 1.0f / l_run[r] : 0.f;
#pragma unroll
            for (int i = 0; i < OD; ++i) {
                const float v0 = o_acc[i][2 * r] * sc0[r] + mb[i * 4 + 2 * r] * sc1[r];
                const float v1 = o_acc[i][2 * r + 1] * sc0[r] + mb[i * 4 + 2 * r + 1] * sc1[r];
                *reinterpret_cast<uint32_t*>(op + qrow * a.o_ss + i * 8 + 2 * t) = Cvt<T>::pack2(v0 * inv, v1 * inv);
            }
        }
    }
    }
}

template <typename T, int D>
__global__ void __launch_bounds__(kAttnThreads, D == 64 ? SMB_ATTN_MINBLOCKS : 1) attention_kernel(const AttnArgs a) {
    extern __shared__ __align__(128) uint8_t attn_smem[];
    pdl_trigger();
    pdl_wait();
    // Launch order = cost order: when the last query tile is partial (ViT: 577 = 9 x 64 + 1) its CTAs come last, so
    // the full tiles each get an SM of their own and the nearly empty ones share (160 CTAs on 148 SMs).
    int qt = blockIdx.x, h = blockIdx.y;
    const int q_tiles = gridDim.x, heads = gridDim.y;
    if (q_tiles > 1 && (a.q_len % kAttnBQ) != 0) {
        const int item = blockIdx.x + q_tiles * blockIdx.y, n_full = (q_tiles - 1) * heads;
        if (item < n_full) { qt = item % (q_tiles - 1); h = item / (q_tiles - 1); }
        else { qt = q_tiles - 1; h = item - n_full; }
    }
    attention_body<T, D>(a, qt, h, blockIdx.z, attn_smem, threadIdx.x);
}

}  // namespace smb
