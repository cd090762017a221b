// Weight-streaming GEMV, bulk-copy + tensor-core edition: same contract, prologues and epilogues as gemv.cuh.
//
//   y[n] = epilogue( sum_k W[n, k] * pro(x)[k] )           W: [N, K] row-major, model dtype T
//
// * The weight rows travel HBM -> shared memory through cp.async.bulk (the TMA engine) into a ring of 32 KB
//   stages filled by one producer warp, so a CTA keeps ~160 KB of weight bytes in flight without holding them
//   in registers, and the ring starts to fill before the previous kernel has finished (weights never depend on
//   it; programmatic dependent launch).
// * The dot products run on the tensor pipe: a stage is a 16-row x 1024-column slab; each of the 8 consumer
//   warps takes a 128-column slice and issues mma.sync.m16n8k16 with the weights as the A operand (ldmatrix
//   from padded rows) and pro(x) as column 0 of the B operand.  fp32 accumulation, exact fp16/bf16 products --
//   the arithmetic of the scalar kernel -- but 3 instructions per 512 weight bytes instead of ~26 per 16: the
//   scalar version was issue-bound at ~46 GB/s per SM (profiles/r01_ncu_full_summary.md), which caps a small
//   ("background") grid at ~1 TB/s and a full grid at ~4.5 TB/s.
//
// Decomposition: the N rows are cut into row blocks (multiple of 16, <= 128 rows); block b goes to CTA
// b % gridDim.x.  Stage sequence inside a block: for each 16-row tile, for each 1024-column chunk, for each
// matrix (NMAT = 2: SwiGLU gate / up).  Warp w accumulates its slice over the chunks of a tile, then writes one
// partial per row -> part[row][w]; after the last stage of the block thread r sums the 8 partials of row r in a
// FIXED order (results do not depend on the grid size) and applies the epilogue.
#pragma once
#include "attention.cuh"   // ldmatrix_x4, mma16816
#include "gemv.cuh"

namespace smb {

constexpr int kGtTileRows = 16;
constexpr int kGtChunk = 1024;                                   // columns per stage (K <= 1024: the whole row)
constexpr int kGtRowPitch = kGtChunk * 2 + 16;                   // padded smem row: conflict-free ldmatrix
constexpr int kGtStageBytes = kGtTileRows * kGtRowPitch;         // 33 024 B
constexpr int kGtStages = 5;
constexpr int kGtConsumerWarps = 8;
constexpr int kGtConsumers = kGtConsumerWarps * 32;
constexpr int kGtThreads = kGtConsumers + 32;   // + 1 producer warp
constexpr int kGtMaxBlockRows = 128;
constexpr int kGtBarAll = 1;                    // named barrier of the 256 consumer threads

inline bool gemv_tma_supported(int K) { return K % 128 == 0 && (K <= kGtChunk || K % kGtChunk == 0) && K <= 16384; }
inline int gemv_tma_smem_bytes(int K, int nmat) {
    return kGtStages * kGtStageBytes + ((K * 2 + 15) & ~15) + 16 + 2 * nmat * kGtMaxBlockRows * kGtConsumerWarps * 4 + 256 + 128;
}

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// position of element k of pro(x) in the fragment-ordered staging vector: k-step s = k / 16 holds, for lane
// pair t = 0..3, the four values {x[16s + 2t], x[16s + 2t + 1], x[16s + 8 + 2t], x[16s + 8 + 2t + 1]} = the
// (b0, b1) registers of the m16n8k16 B fragment of lane t (column n = 0)
__device__ __forceinline__ int gt_xpos(int k) {
    const int w = k & 15;
    return (k & ~15) + ((w & 7) >> 1) * 4 + (w >> 3) * 2 + (w & 1);
}

template <typename T, int NMAT>
__global__ void __launch_bounds__(kGtThreads, 1) gemv_tma_kernel(const GemvArgs a, int rows_per_block, int nblocks) {
    extern __shared__ __align__(128) uint8_t gt_smem[];
    uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gt_smem) + 127) & ~uintptr_t(127));
    T* xs = reinterpret_cast<T*>(ring + kGtStages * kGtStageBytes);                       // [K] fragment order + 16 B zeros
    const int xs_bytes = ((a.K * 2 + 15) & ~15) + 16;
    float* part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(xs) + xs_bytes);   // [2][NMAT][rows][8]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(part + 2 * NMAT * kGtMaxBlockRows * kGtConsumerWarps);
    uint64_t* empty_bar = full_bar + kGtStages;
    __shared__ float red[kGtConsumerWarps * 2];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = a.K;
    const int KC = min(K, kGtChunk);
    const int nchunk = K / KC;
    const uint32_t seg_bytes = static_cast<uint32_t>(KC) * 2u;
    const size_t row_bytes = static_cast<size_t>(K) * 2u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGtStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kGtConsumerWarps); }
        fence_mbar_init();
    }
    pdl_trigger();
    __syncthreads();

    if (warp == kGtConsumerWarps) {
        // ---------------------------------------------------------------- producer (no griddepcontrol.wait: weights only)
        uint32_t sc = 0;   // global stage counter
        for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
            const int r0 = b * rows_per_block, r1 = min(a.N, r0 + rows_per_block);
            for (int r = r0; r < r1; r += kGtTileRows) {
                const int nr = min(kGtTileRows, r1 - r);
                for (int c = 0; c < nchunk; ++c) {
#pragma unroll
                    for (int m = 0; m < NMAT; ++m) {
                        const int st = sc % kGtStages;
                        mbar_wait(&empty_bar[st], ((sc / kGtStages) & 1u) ^ 1u);
                        if (lane == 0) mbar_arrive_expect_tx(&full_bar[st], nr * seg_bytes);
                        __syncwarp();
                        if (lane < nr) {
                            const char* src = reinterpret_cast<const char*>(m == 0 ? a.W0 : a.W1) +
                                              static_cast<size_t>(r + lane) * row_bytes + static_cast<size_t>(c) * seg_bytes;
                            bulk_load(ring + st * kGtStageBytes + lane * kGtRowPitch, src, seg_bytes, &full_bar[st], kEvictFirst);
                        }
                        __syncwarp();
                        ++sc;
                    }
                }
            }
        }
        return;
    }

    // -------------------------------------------------------------------- consumers (256 threads)
    const int ct = threadIdx.x;   // 0..255
    pdl_wait();
    {
        const T* x0 = reinterpret_cast<const T*>(a.x0);
        const T* nw = reinterpret_cast<const T*>(a.nw);
        const T* nb = reinterpret_cast<const T*>(a.nb);
        if (ct < 8) xs[K + ct] = Cvt<T>::from_f(0.f);   // zero pad read by the lanes that carry no B data
        if (a.pro == PRO_PLAIN) {
            for (int k = ct; k < K; k += kGtConsumers) xs[gt_xpos(k)] = x0[k];
        } else if (a.pro == PRO_GQA_EXPAND) {
            for (int k = ct; k < K; k += kGtConsumers) {
                const int hh = k / a.head_dim, d = k - hh * a.head_dim;
                xs[gt_xpos(k)] = x0[(hh / a.gqa_rep) * a.head_dim + d];
            }
        } else {
            float s1 = 0.f, s2 = 0.f;
            for (int k = ct; k < K; k += kGtConsumers) {
                const float v = Cvt<T>::to_f(x0[k]);
                s1 += v;
                s2 += v * v;
            }
            s1 = warp_sum(s1);
            s2 = warp_sum(s2);
            if (lane == 0) { red[warp] = s1; red[kGtConsumerWarps + warp] = s2; }
            named_bar_sync(kGtBarAll, kGtConsumers);
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int w = 0; w < kGtConsumerWarps; ++w) { t1 += red[w]; t2 += red[kGtConsumerWarps + w]; }
            const float inv_k = 1.0f / static_cast<float>(K);
            if (a.pro == PRO_RMSNORM) {
                const float r = rsqrtf(t2 * inv_k + a.eps);
                for (int k = ct; k < K; k += kGtConsumers) {
                    const float v = rnd<T>(Cvt<T>::to_f(x0[k]) * r);
                    xs[gt_xpos(k)] = Cvt<T>::from_f(Cvt<T>::to_f(nw[k]) * v);
                }
            } else {
                const float mean = t1 * inv_k;
                float sq = 0.f;
                for (int k = ct; k < K; k += kGtConsumers) {
                    const float d = Cvt<T>::to_f(x0[k]) - mean;
                    sq += d * d;
                }
                sq = warp_sum(sq);
                named_bar_sync(kGtBarAll, kGtConsumers);
                if (lane == 0) red[warp] = sq;
                named_bar_sync(kGtBarAll, kGtConsumers);
                float var = 0.f;
#pragma unroll
                for (int w = 0; w < kGtConsumerWarps; ++w) var += red[w];
                const float r = rsqrtf(var * inv_k + a.eps);
                for (int k = ct; k < K; k += kGtConsumers) {
                    float v = (Cvt<T>::to_f(x0[k]) - mean) * r * Cvt<T>::to_f(nw[k]) + Cvt<T>::to_f(nb[k]);
                    v = rnd<T>(v);
                    if (a.pro == PRO_LN_LEAKY) v = leaky_relu_f(v);
                    xs[gt_xpos(k)] = Cvt<T>::from_f(v);
                }
            }
        }
    }
    named_bar_sync(kGtBarAll, kGtConsumers);

    // this warp's slice of a chunk: columns [warp * KW, (warp + 1) * KW), KW = KC / 8 (multiple of 16)
    const int KW = KC / kGtConsumerWarps;
    const int ksteps = KW / 16;
    // ldmatrix.x4 source of this lane: row (lane & 7) + 8 * ((lane >> 3) & 1), column offset 8 * (lane >> 4)
    const uint32_t a_lane_off = static_cast<uint32_t>(((lane & 7) + ((lane >> 3) & 1) * 8) * kGtRowPitch + (warp * KW + (lane >> 4) * 8) * 2);
    const uint32_t ring_u32 = smem_u32(ring);
    const uint32_t xs_u32 = smem_u32(xs);
    const uint32_t zero_u32 = xs_u32 + static_cast<uint32_t>(K) * 2u;
    const bool has_b = lane < 4;

    uint32_t sc = 0;
    int pbuf = 0;
    for (int b = blockIdx.x; b < nblocks; b += gridDim.x, pbuf ^= 1) {
        const int r0 = b * rows_per_block, r1 = min(a.N, r0 + rows_per_block);
        const int nrows = r1 - r0;
        float* pb = part + pbuf * (NMAT * kGtMaxBlockRows * kGtConsumerWarps);
        for (int r = r0; r < r1; r += kGtTileRows) {
            float acc[NMAT][4];
#pragma unroll
            for (int m = 0; m < NMAT; ++m) { acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f; }
            for (int c = 0; c < nchunk; ++c) {
                // B fragments of this (chunk, warp) slice: 8 bytes per k-step from the fragment-ordered vector
                const uint32_t xb = has_b ? xs_u32 + static_cast<uint32_t>((c * KC + warp * KW) * 2 + lane * 8) : zero_u32;
                const uint32_t xstep = has_b ? 32u : 0u;
#pragma unroll
                for (int m = 0; m < NMAT; ++m) {
                    const int st = sc % kGtStages;
                    mbar_wait(&full_bar[st], (sc / kGtStages) & 1u);
                    const uint32_t abase = ring_u32 + st * kGtStageBytes + a_lane_off;
#pragma unroll 4
                    for (int ks = 0; ks < ksteps; ++ks) {
                        uint32_t af[4], b0, b1;
                        ldmatrix_x4(af, abase + ks * 32);
                        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(xb + ks * xstep));
                        mma16816<T>(acc[m], af, b0, b1);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[st]);
                    ++sc;
                }
            }
            // column 0 of the accumulator tile lives in the lanes with lane % 4 == 0: rows lane/4 and lane/4 + 8
            if ((lane & 3) == 0) {
                const int ra = r - r0 + (lane >> 2), rb2 = ra + 8;
#pragma unroll
                for (int m = 0; m < NMAT; ++m) {
                    if (ra < nrows) pb[(m * kGtMaxBlockRows + ra) * kGtConsumerWarps + warp] = acc[m][0];
                    if (rb2 < nrows) pb[(m * kGtMaxBlockRows + rb2) * kGtConsumerWarps + warp] = acc[m][2];
                }
            }
        }
        named_bar_sync(kGtBarAll, kGtConsumers);   // all partials of this block are in smem

        // ---- epilogue: one thread per row, fixed-order sum of the 8 column-slice partials
        for (int r = ct; r < nrows; r += kGtConsumers) {
            const int n = r0 + r;
            float acc = 0.f, accb = 0.f;
#pragma unroll
            for (int w = 0; w < kGtConsumerWarps; ++w) {
                acc += pb[r * kGtConsumerWarps + w];
                if (NMAT == 2) accb += pb[(kGtMaxBlockRows + r) * kGtConsumerWarps + w];
            }
            const T* bias = reinterpret_cast<const T*>(a.bias);
            if (bias != nullptr) acc += Cvt<T>::to_f(bias[n]);
            T* y = reinterpret_cast<T*>(a.y);
            T* resid = reinterpret_cast<T*>(a.resid);
            switch (a.epi) {
                case GEPI_STORE: y[n] = Cvt<T>::from_f(acc); break;
                case GEPI_LEAKY: y[n] = Cvt<T>::from_f(leaky_relu_f(rnd<T>(acc))); break;
                case GEPI_RESID: resid[n] = Cvt<T>::from_f(Cvt<T>::to_f(resid[n]) + rnd<T>(acc)); break;
                case GEPI_ADD_TO: y[n] = Cvt<T>::from_f(Cvt<T>::to_f(resid[n]) + rnd<T>(acc)); break;
                case GEPI_F32: reinterpret_cast<float*>(a.y)[n] = rnd<T>(acc); break;
                case GEPI_SWIGLU: {
                    const float g = rnd<T>(silu_f(rnd<T>(acc)));
                    y[n] = Cvt<T>::from_f(g * rnd<T>(accb));
                    break;
                }
                case GEPI_MAMBA_CONV: {
                    if (n < a.d_inner) {
                        T* stt = reinterpret_cast<T*>(a.conv_state) + static_cast<size_t>(n) * a.d_conv;
                        const T* cw = reinterpret_cast<const T*>(a.conv_w) + static_cast<size_t>(n) * a.d_conv;
                        const float xn = rnd<T>(acc);
                        float c = 0.f;
                        for (int w = 0; w < a.d_conv - 1; ++w) {
                            const T sv = stt[w + 1];
                            stt[w] = sv;
                            c = fmaf(Cvt<T>::to_f(sv), Cvt<T>::to_f(cw[w]), c);
                        }
                        stt[a.d_conv - 1] = Cvt<T>::from_f(xn);
                        c = fmaf(xn, Cvt<T>::to_f(cw[a.d_conv - 1]), c);
                        c = rnd<T>(c + Cvt<T>::to_f(reinterpret_cast<const T*>(a.conv_b)[n]));
                        y[n] = Cvt<T>::from_f(silu_f(c));
                    } else {
                        reinterpret_cast<T*>(a.z_out)[n - a.d_inner] = Cvt<T>::from_f(acc);
                    }
                    break;
                }
            }
        }
    }
}

}  // namespace smb
