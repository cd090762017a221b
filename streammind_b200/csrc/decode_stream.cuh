// Persistent weight-streaming kernel for the batch-1 Mistral chains on the path: ONE launch = one greedy decode
// step of the LLM for NV streams (hf MistralForCausalLM.forward at L = 1 per stream + argmax; reference call site:
// streammind/model/language_model/videollama2_mistral.py:426-431).  HBM-bound: 14.22 GB of weights per step.
//
// Why one kernel: as separate launches (7 per layer) every projection pays a prologue (stage + normalise the input
// vector), a pipeline fill and a tail, during which HBM idles: 0.64 of the HBM roofline at ctx 4k (round 1).  Here
// one CTA per SM walks a device-resident op list (GEMV, attention, final argmax) with a grid barrier after each op,
// and the WEIGHT stream never stops: producer warps run ahead of the consumers through a shared-memory ring
// (kDsSlots x 32 KB, filled by cp.async.bulk, one mbarrier pair per slot) across op boundaries -- weights do not
// depend on activations -- so epilogue + grid barrier + next prologue (~2-3 us) hide under up to 192 KB / SM of
// weights already in flight (tools/hbm_read_bench.cu, profiles/r02_hbm_read_bench.txt: bulk copies of >= 32 KB from
// two issuing threads per SM reach 7.1 TB/s, 16-byte LDG streams 7.3; small copies are issue-bound).
//
// Work split of a GEMV op  y[n] = epi(sum_k W[n, k] * pro(x)[k]):  CTA c owns rows [c * rpc, (c+1) * rpc) of every
// matrix of the op; a ring slot holds R = 8 / P consecutive rows (two matrices: R/2 rows of each, the gate and up rows of
// the same outputs); consumer warp w reduces part w % P (K / P columns) of slot row w / P against the NV staged
// vectors: fp32 FMA chains, one butterfly per (row, part, vector), partials summed in FIXED order by the epilogue thread
// of the row (deterministic: greedy decode is reproducible run to run).
//
// Attention op (one query per stream): CTA (stream, kv head, KV slice) applies RoPE to the group's q heads and to the
// new k, appends k / v to the cache, runs online softmax over its slice (all query heads of the GQA group share each
// K / V read), writes an un-normalised partial; the LAST CTA to finish a (stream, kv head) merges the slices in
// fixed order (split-KV combine folded in: no extra launch, no extra grid barrier).
// Rounding points are those of the reference (oracle/restate.py mistral_forward): RMSNorm output, every projection
// output, RoPE products, softmax probabilities before P.V, silu and its product, residual sums, logits.
#pragma once
#include "ptx.cuh"

namespace smb {

constexpr int kDsGroupWarps = 8;            // warps that share one ring slot: one (row, part) each
#ifndef SMB_DS_GROUPS
#define SMB_DS_GROUPS 1
#endif
constexpr int kDsGroups = SMB_DS_GROUPS;    // consumer groups; group g takes the chunks with seq % kDsGroups == g
constexpr int kDsConsumerWarps = kDsGroupWarps * kDsGroups;
constexpr int kDsProducerWarps = 2;
constexpr int kDsConsumerThreads = kDsConsumerWarps * 32;
constexpr int kDsThreads = (kDsConsumerWarps + kDsProducerWarps) * 32;
constexpr int kDsSlotBytes = 32 * 1024;
constexpr int kDsMaxSlots = 6;
constexpr int kDsMaxStreams = 4;

enum DsOpType : int { DS_GEMV = 0, DS_ATTN = 1, DS_FINAL = 2 };
enum DsPro : int {
    DSP_PLAIN = 0,          // x = x0
    DSP_RMSNORM = 1,        // x = nw * T(x0 * rsqrt(mean(x0^2) + eps))         (hf MistralRMSNorm)
    DSP_EMBED_RMSNORM = 2,  // same with x0 = embed[token of the stream]        (first layer: embed_tokens fused in)
};
enum DsEpi : int {
    DSE_STORE = 0,        // y = T(acc)
    DSE_RESID = 1,        // resid[n] = T(resid[n] + T(acc))                     (in-place residual stream)
    DSE_RESID_EMBED = 2,  // resid[n] = T(embed[token][n] + T(acc))              (first layer: the residual stream starts here)
    DSE_SWIGLU = 3,       // two matrices: y = T(T(silu(T(acc0))) * T(acc1))     (hf MistralMLP)
    DSE_LOGITS = 4,       // y(float) = float(T(acc)); per-CTA argmax candidate  (lm_head, logits leave it in T)
};

struct DsOp {
    int type;
    // ---- DS_GEMV
    const void* W0;
    const void* W1;
    int nmat, N, K;        // N rows per matrix
    int R, P;              // slot rows (both matrices together), parts per row; R * P == kDsGroupWarps
    int pro, epi;
    const void* x;         // [NV][x_stride] model dtype (ignored for DSP_EMBED_RMSNORM)
    long long x_stride;
    const void* nw;        // RMSNorm weight
    float eps;
    void* y;               // [NV][y_stride] (T, or float for DSE_LOGITS)
    long long y_stride;
    void* resid;           // [NV][resid_stride]
    long long resid_stride;
    // ---- DS_ATTN
    const void* qkv;       // [NV][qkv_stride]: q (Hq*D) | k (Hk*D) | v (Hk*D) of the new token, before RoPE
    long long qkv_stride;
    void* kc;              // K cache of this layer: [n_streams][Hk][max_ctx][D]
    void* vc;
    long long kv_stream_stride;   // elements between the caches of consecutive streams
    void* att;             // [NV][Hq*D]
    int Hq, Hk, max_ctx;
    float rope_theta, scale_log2e;
};

// per-stream decode state in device memory (so one captured graph serves every step)
struct DsStreamState {
    int pos;        // position of the token being fed = KV length before this step
    int tok;        // token to feed
    int n_out;      // tokens produced so far (this sm_llm_decode call)
    int done;       // stop id produced or max_new reached
    int max_new;
    int kv_slot;    // which cache (stream index of the handle) this lane works on
    int pad[2];
};

struct DsParams {
    const DsOp* ops;
    int n_ops;
    int n_slots;             // ring depth (<= kDsMaxSlots)
    int xcap;                // elements per staged vector (max K over the ops, multiple of 8)
    int x_bytes;             // bytes of the staging region: max(NV * xcap * 2, attention scratch)
    int part_cap;            // floats in the partial-sum area
    unsigned* sync;          // [0] grid-barrier counter, [1] epoch (steps run since reset), [2] all-done flag, [8..] per (pass, lane, kv head) arrival counters
    int n_barriers;          // grid barriers per step
    DsStreamState* st;       // [NV]
    int* out_ids;            // [NV][out_stride]
    int out_stride;
    const int* stop;         // [0] = count, then ids
    const void* embed;       // [vocab][H]
    int H;
    float* att_part;         // [NV][Hq][S][D + 2] split-KV partials
    float* cand_val;         // [NV][gridDim.x] per-CTA argmax candidates
    int* cand_idx;
    int l2_ahead;            // chunks the L2 prefetch cursor runs ahead of the ring (even; 0 = off)
    int dbg_flags;           // measurement only: 1 skip the consumer math, 2 skip grid barriers, 4 skip the attention body, 8 no L2 prefetch
    long long* dbg;          // optional: CTA 0 accumulates ns per phase (0 prologue, 1 ring compute, 2 epilogue, 3 barrier, 4 attention, 5 final, 6 wait for an op's first chunk, 7 wait for later chunks)
};

__device__ __forceinline__ void ds_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ds_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned ds_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ds_red_release(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ds_atom_add_acq_rel(unsigned* p, unsigned v) {
    unsigned old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ long long ds_gtimer() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// mbarrier wait for the hot loops: try_wait with a suspend-time hint, so a waiting warp sleeps in hardware instead of
// executing a poll loop (bounded: a protocol error traps instead of hanging the box)
__device__ __forceinline__ void mbar_wait_hint(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(100000u) : "memory");
        if (!ok && ++spins > (1u << 20)) { printf("smb: decode ring mbarrier timeout (cta %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    } while (!ok);
}
__device__ __forceinline__ void ds_consumer_sync() { named_bar_sync(1, kDsConsumerThreads); }

// Grid barrier of the consumer side (thread 0 of every CTA arrives and polls; bounded spin so that a logic error traps
// instead of hanging the box).  All CTAs are co-resident: the launch uses one CTA per SM.
__device__ __forceinline__ void ds_grid_barrier(unsigned* ctr, unsigned target, bool skip = false) {
    if (skip) { ds_consumer_sync(); return; }
    // The CTA barrier orders every consumer thread's global writes before thread 0's release (release is cumulative);
    // the acquire on the other side is ordered before the other threads' reads by the second CTA barrier.  Data written by
    // other CTAs is read with ld.global.cg (L2), never through a possibly stale L1 line.
    ds_consumer_sync();
    if (threadIdx.x == 0) {
        ds_red_release(ctr, 1u);
        unsigned spins = 0;
        while (static_cast<int>(ds_ld_acquire(ctr) - target) < 0) {
            if (++spins > (1u << 25)) { printf("smb: decode grid barrier timeout (cta %d, target %u)\n", blockIdx.x, target); __trap(); }
        }
    }
    ds_consumer_sync();
}

template <typename T>
__device__ __forceinline__ float ds_dot8(const uint4& w, const uint4& xv, float s) {
    const float2 w0 = Cvt<T>::unpack2(w.x), w1 = Cvt<T>::unpack2(w.y), w2 = Cvt<T>::unpack2(w.z), w3 = Cvt<T>::unpack2(w.w);
    const float2 x0 = Cvt<T>::unpack2(xv.x), x1 = Cvt<T>::unpack2(xv.y), x2 = Cvt<T>::unpack2(xv.z), x3 = Cvt<T>::unpack2(xv.w);
    s = fmaf(w0.x, x0.x, s); s = fmaf(w0.y, x0.y, s);
    s = fmaf(w1.x, x1.x, s); s = fmaf(w1.y, x1.y, s);
    s = fmaf(w2.x, x2.x, s); s = fmaf(w2.y, x2.y, s);
    s = fmaf(w3.x, x3.x, s); s = fmaf(w3.y, x3.y, s);
    return s;
}

// D (fp32, 16 x 8) += A (16 x 16, row) * B (16 x 8, col) on the tensor cores; inputs in the model dtype
template <typename T>
__device__ __forceinline__ void ds_mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    if constexpr (Cvt<T>::kBf16) {
        asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    } else {
        asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
}
__device__ __forceinline__ float ds_silu(float v) { return v / (1.0f + __expf(-v)); }
// 16-bit load that bypasses L1 (activations written by other CTAs earlier in the same launch)
template <typename T>
__device__ __forceinline__ T ds_ldcg_t(const T* p) {
    const unsigned short u = __ldcg(reinterpret_cast<const unsigned short*>(p));
    T t;
    memcpy(&t, &u, 2);
    return t;
}

template <typename T, int NV>
__global__ void __launch_bounds__(kDsThreads, 1) decode_stream_kernel(const DsParams p) {
    extern __shared__ __align__(128) uint8_t ds_smem[];
    uint8_t* ring = ds_smem;                                                       // [n_slots][32 KB]
    T* xs = reinterpret_cast<T*>(ds_smem + static_cast<size_t>(p.n_slots) * kDsSlotBytes);   // [NV][xcap]; attention scratch aliases it
    float* part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(xs) + p.x_bytes);
    __shared__ uint64_t full_bar[kDsMaxSlots], empty_bar[kDsMaxSlots];
    __shared__ float red[kDsConsumerWarps * NV];
    __shared__ float cand_v[kDsConsumerWarps];
    __shared__ int cand_i[kDsConsumerWarps];
    __shared__ int flag_last;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x, cta = blockIdx.x;
    if (ds_ld_acquire(p.sync + 2) != 0u) return;        // every stream already finished: this launch is a no-op
    const unsigned epoch = ds_ld_acquire(p.sync + 1);
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.n_slots; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kDsGroupWarps); }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp >= kDsConsumerWarps) {
        // ================================================================== producers: the weight stream
        // Each producer walks the chunk sequence of this CTA twice: a far cursor that asks L2 to fetch chunk seq + l2_ahead
        // from HBM (cp.async.bulk.prefetch.L2: HBM keeps streaming while the consumers sit in an epilogue / grid barrier /
        // prologue and the ring is full), and the ring cursor that copies chunk seq into its slot.
        if (lane != 0) return;
        const int pi = warp - kDsConsumerWarps;
        struct Cursor { int oi, j, j1, RJ, seq; };
        auto open_op = [&](Cursor& c) {            // position the cursor on the first chunk of the next GEMV op (oi = n_ops: end)
            for (; c.oi < p.n_ops; ++c.oi) {
                const DsOp& op = p.ops[c.oi];
                if (op.type != DS_GEMV) continue;
                const int rpc = (op.N + G - 1) / G;
                c.j = min(op.N, cta * rpc); c.j1 = min(op.N, c.j + rpc); c.RJ = op.R / op.nmat;
                if (c.j < c.j1) return;
            }
        };
        auto advance = [&](Cursor& c) {
            c.j += c.RJ; ++c.seq;
            if (c.j >= c.j1) { ++c.oi; open_op(c); }
        };
        auto prefetch = [&](const Cursor& c) {
            const DsOp& op = p.ops[c.oi];
            const size_t row_bytes = static_cast<size_t>(op.K) * sizeof(T);
            const uint32_t bytes = static_cast<uint32_t>(min(c.RJ, c.j1 - c.j) * row_bytes);
            ds_prefetch_l2(static_cast<const uint8_t*>(op.W0) + c.j * row_bytes, bytes);
            if (op.nmat == 2) ds_prefetch_l2(static_cast<const uint8_t*>(op.W1) + c.j * row_bytes, bytes);
        };
        Cursor cur{0, 0, 0, 1, 0}, far{0, 0, 0, 1, 0};
        open_op(cur);
        open_op(far);
        const int ahead = (p.dbg_flags & 8) ? 0 : p.l2_ahead;
        for (int i = 0; i < ahead && far.oi < p.n_ops; ++i) {
            if (far.seq % kDsProducerWarps == pi) prefetch(far);
            advance(far);
        }
        while (cur.oi < p.n_ops) {
            if (cur.seq % kDsProducerWarps == pi) {
                if (ahead > 0 && far.oi < p.n_ops) prefetch(far);      // far.seq == cur.seq + ahead: same producer parity when ahead is even
                const DsOp& op = p.ops[cur.oi];
                const size_t row_bytes = static_cast<size_t>(op.K) * sizeof(T);
                const int slot = cur.seq % p.n_slots, use = cur.seq / p.n_slots;
                if (use > 0) mbar_wait_hint(&empty_bar[slot], (use - 1) & 1);
                const int nr = min(cur.RJ, cur.j1 - cur.j);
                const uint32_t bytes = static_cast<uint32_t>(nr * row_bytes);
                mbar_arrive_expect_tx(&full_bar[slot], bytes * op.nmat);
                uint8_t* dst = ring + static_cast<size_t>(slot) * kDsSlotBytes;
                ds_bulk_g2s(dst, static_cast<const uint8_t*>(op.W0) + cur.j * row_bytes, bytes, &full_bar[slot]);
                if (op.nmat == 2)
                    ds_bulk_g2s(dst + cur.RJ * row_bytes, static_cast<const uint8_t*>(op.W1) + cur.j * row_bytes, bytes, &full_bar[slot]);
            }
            advance(cur);
            if (far.oi < p.n_ops) advance(far);
        }
        return;
    }

    // ====================================================================== consumers
    const int tid = threadIdx.x;                      // 0 .. kDsConsumerThreads-1
    const bool timed = p.dbg != nullptr && cta == 0 && tid == 0;
    long long t_last = timed ? ds_gtimer() : 0;
    long long t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};      // accumulated in registers (a global read-modify-write per stamp would dominate)
    auto stamp = [&](int cat) {
        if (timed) {
            const long long t = ds_gtimer();
#pragma unroll
            for (int c = 0; c < 8; ++c) t_acc[c] += c == cat ? t - t_last : 0;
            t_last = t;
        }
    };
    unsigned bar_target = epoch * static_cast<unsigned>(p.n_barriers) * G;
    int seq = 0, ring_slot = 0, ring_par = 0;            // chunk sequence number, its ring slot and the parity of that slot's use
    for (int oi = 0; oi < p.n_ops; ++oi) {
        const DsOp& op = p.ops[oi];
        if (op.type == DS_GEMV) {
            const int K = op.K;
            // ---- prologue: stage pro(x_v) for every stream
#pragma unroll 1
            for (int v = 0; v < NV; ++v) {
                const T* x0 = op.pro == DSP_EMBED_RMSNORM
                                  ? reinterpret_cast<const T*>(p.embed) + static_cast<size_t>(p.st[v].tok) * p.H
                                  : reinterpret_cast<const T*>(op.x) + v * op.x_stride;
                T* xv = xs + static_cast<size_t>(v) * p.xcap;
                if (op.pro == DSP_PLAIN) {
                    for (int k = tid * 8; k < K; k += kDsConsumerThreads * 8)
                        *reinterpret_cast<uint4*>(xv + k) = __ldcg(reinterpret_cast<const uint4*>(x0 + k));   // written by other CTAs: L2 only
                } else {
                    float s2 = 0.f;
                    for (int k = tid * 8; k < K; k += kDsConsumerThreads * 8) {
                        const uint4 u = __ldcg(reinterpret_cast<const uint4*>(x0 + k));
                        *reinterpret_cast<uint4*>(xv + k) = u;
                        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) { const float2 f = Cvt<T>::unpack2(w[i]); s2 = fmaf(f.x, f.x, s2); s2 = fmaf(f.y, f.y, s2); }
                    }
                    s2 = warp_sum(s2);
                    if (lane == 0) red[warp * NV + v] = s2;
                }
            }
            ds_consumer_sync();
            if (op.pro != DSP_PLAIN) {
#pragma unroll 1
                for (int v = 0; v < NV; ++v) {
                    float t2 = 0.f;
#pragma unroll
                    for (int w = 0; w < kDsConsumerWarps; ++w) t2 += red[w * NV + v];
                    const float r = rsqrtf(t2 / static_cast<float>(K) + op.eps);
                    T* xv = xs + static_cast<size_t>(v) * p.xcap;
                    const T* nw = reinterpret_cast<const T*>(op.nw);
                    for (int k = tid * 8; k < K; k += kDsConsumerThreads * 8) {
                        const uint4 u = *reinterpret_cast<const uint4*>(xv + k);
                        const uint4 g = *reinterpret_cast<const uint4*>(nw + k);
                        const uint32_t w[4] = {u.x, u.y, u.z, u.w}, gw[4] = {g.x, g.y, g.z, g.w};
                        uint32_t o[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 f = Cvt<T>::unpack2(w[i]), gg = Cvt<T>::unpack2(gw[i]);
                            o[i] = Cvt<T>::pack2(gg.x * rnd<T>(f.x * r), gg.y * rnd<T>(f.y * r));
                        }
                        *reinterpret_cast<uint4*>(xv + k) = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                }
                ds_consumer_sync();
            }
            stamp(0);
            // ---- stream this CTA's rows through the ring
            const int rpc = (op.N + G - 1) / G;
            const int j0 = min(op.N, cta * rpc), j1 = min(op.N, j0 + rpc);
            const int nloc = j1 - j0;
            const int RJ = op.R / op.nmat, P = op.P;
            const int grp = warp / kDsGroupWarps, wg = warp - grp * kDsGroupWarps;   // consumer group and warp within it
            const int sr = wg / P, pt = wg - sr * P;                   // slot row and part of this warp
            const int m = sr / RJ, jr = sr - m * RJ;                   // matrix and row within the chunk
            const int cols = K / P, c0 = pt * cols;
            // The reduction runs on the tensor cores (mma.sync m16n8k16, fp32 accumulate) with a "diagonal" arrangement: the B
            // operand's 8 columns are 8 consecutive 16-weight segments of ONE weight row (256 contiguous bytes: lane l reads
            // bytes [8 l, 8 l + 8), conflict-free), the A operand's rows 0..7 are the matching 16-element segments of vector 0
            // (rows 8..15: vector 1), so D[m][m] accumulates sum_k w[16 m + k] x[16 m + k]; off-diagonal products are discarded.
            // A slot costs 128 HMMA + 256 LDS.64 instead of ~1100 FMA / unpack / LDS, and a second stream rides along for free.
            // Measured (tools/decode_probe.py, profiles/r02_decode_kernel.md): this drains a slot in ~800 clk -- about the rate
            // HBM delivers -- because legacy HMMA on sm_100 issues one m16n8k16 per ~26 clk per sub-partition; an fp32-FMA
            // consumer with the vector in registers is equally issue-bound, and running both pipes side by side did not beat
            // either.  Per (row, part): two accumulators (even / odd blocks), one butterfly of the 8 diagonal elements;
            // partials over the parts are summed in FIXED order by the epilogue.
            const int nblk = cols / 128;                                   // 128-weight blocks of this warp's part (even)
            constexpr int NPAIR = (NV + 1) / 2;                            // A operands: vectors (2 q, 2 q + 1) on rows (0..7, 8..15)
            const bool diag0 = (lane >> 2) == 2 * (lane & 3), diag1 = (lane >> 2) == 2 * (lane & 3) + 1;   // this lane holds D[g][g] in c0 / c1
            const uint2* xp = reinterpret_cast<const uint2*>(xs + c0) + lane;   // block b of vector v: xp[v * xq + 32 b]
            const int xq = p.xcap / 4;
            // The hot loop is kept branch-free and division-free (ncu, profiles/r02_decode_kernel.md: an earlier form with a
            // guard per block executed ~3700 warp instructions per 32 KB slot, 22 % of them mbarrier polls): blocks go in
            // unguarded batches of 8 and 2 (nblk is even), the ring position advances incrementally.
            const uint2* ring_u2 = reinterpret_cast<const uint2*>(ring) + (static_cast<size_t>(sr) * K + c0) / 4 + lane;
            float* part_w = part + (static_cast<size_t>(m) * nloc * P + pt) * NV;           // + local row * P * NV
            const int skip_math = p.dbg_flags & 1;
            for (int j = j0; j < j1; j += RJ, ++seq) {
                const int slot = ring_slot, par = ring_par;
                if (++ring_slot == p.n_slots) { ring_slot = 0; ring_par ^= 1; }
                if (kDsGroups > 1 && seq % kDsGroups != grp) continue;
                mbar_wait_hint(&full_bar[slot], par);
                if (j + jr < j1 && !skip_math) {
                    const uint2* wp = ring_u2 + slot * (kDsSlotBytes / 8);
                    float acc[NPAIR][2][4];
#pragma unroll
                    for (int q = 0; q < NPAIR; ++q)
#pragma unroll
                        for (int e = 0; e < 2; ++e) { acc[q][e][0] = 0.f; acc[q][e][1] = 0.f; acc[q][e][2] = 0.f; acc[q][e][3] = 0.f; }
                    int b = 0;
                    for (; b + 8 <= nblk; b += 8) {                  // 8 blocks per batch: all their loads in flight together
                        uint2 w[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) w[u] = wp[(b + u) * 32];
#pragma unroll
                        for (int q = 0; q < NPAIR; ++q) {
                            uint2 xa[8], xb[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                xa[u] = xp[(2 * q) * xq + (b + u) * 32];
                                xb[u] = 2 * q + 1 < NV ? xp[(2 * q + 1) * xq + (b + u) * 32] : make_uint2(0u, 0u);
                            }
#pragma unroll
                            for (int u = 0; u < 8; ++u) ds_mma_16816<T>(acc[q][u & 1], xa[u].x, xb[u].x, xa[u].y, xb[u].y, w[u].x, w[u].y);
                        }
                    }
                    for (; b < nblk; b += 2) {
                        const uint2 w0 = wp[b * 32], w1 = wp[b * 32 + 32];
#pragma unroll
                        for (int q = 0; q < NPAIR; ++q) {
                            const uint2 xa0 = xp[(2 * q) * xq + b * 32], xa1 = xp[(2 * q) * xq + b * 32 + 32];
                            uint2 xb0 = make_uint2(0u, 0u), xb1 = make_uint2(0u, 0u);
                            if (2 * q + 1 < NV) { xb0 = xp[(2 * q + 1) * xq + b * 32]; xb1 = xp[(2 * q + 1) * xq + b * 32 + 32]; }
                            ds_mma_16816<T>(acc[q][0], xa0.x, xb0.x, xa0.y, xb0.y, w0.x, w0.y);
                            ds_mma_16816<T>(acc[q][1], xa1.x, xb1.x, xa1.y, xb1.y, w1.x, w1.y);
                        }
                    }
                    float tot[NV];
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        const int q = v >> 1, o = (v & 1) * 2;
                        const float d0 = acc[q][0][o] + acc[q][1][o], d1 = acc[q][0][o + 1] + acc[q][1][o + 1];
                        tot[v] = warp_sum(diag0 ? d0 : diag1 ? d1 : 0.f);
                    }
                    if (lane == 0) {
                        float* dst = part_w + static_cast<size_t>(j - j0 + jr) * P * NV;
#pragma unroll
                        for (int v = 0; v < NV; ++v) dst[v] = tot[v];
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[slot]);
            }
            ds_consumer_sync();
            stamp(1);
            // ---- epilogue: one thread per output row, fixed-order sum of the parts
            float best = -INFINITY;
            int best_i = 0x7fffffff;
#pragma unroll 1
            for (int v = 0; v < NV; ++v) {
                if (op.epi == DSE_LOGITS) { best = -INFINITY; best_i = 0x7fffffff; }
                for (int r = tid; r < nloc; r += kDsConsumerThreads) {
                    const int n = j0 + r;
                    float a0 = 0.f, a1 = 0.f;
                    for (int q = 0; q < P; ++q) a0 += part[(static_cast<size_t>(r) * P + q) * NV + v];
                    if (op.nmat == 2)
                        for (int q = 0; q < P; ++q) a1 += part[((static_cast<size_t>(nloc) + r) * P + q) * NV + v];
                    switch (op.epi) {
                        case DSE_STORE: (reinterpret_cast<T*>(op.y) + v * op.y_stride)[n] = Cvt<T>::from_f(a0); break;
                        case DSE_RESID: {
                            T* rs = reinterpret_cast<T*>(op.resid) + v * op.resid_stride;
                            rs[n] = Cvt<T>::from_f(Cvt<T>::to_f(ds_ldcg_t(rs + n)) + rnd<T>(a0));
                            break;
                        }
                        case DSE_RESID_EMBED: {
                            const T* e = reinterpret_cast<const T*>(p.embed) + static_cast<size_t>(p.st[v].tok) * p.H;
                            (reinterpret_cast<T*>(op.resid) + v * op.resid_stride)[n] = Cvt<T>::from_f(Cvt<T>::to_f(e[n]) + rnd<T>(a0));
                            break;
                        }
                        case DSE_SWIGLU: {
                            const float g = rnd<T>(ds_silu(rnd<T>(a0)));
                            (reinterpret_cast<T*>(op.y) + v * op.y_stride)[n] = Cvt<T>::from_f(g * rnd<T>(a1));
                            break;
                        }
                        case DSE_LOGITS: {
                            const float lg = rnd<T>(a0);
                            (reinterpret_cast<float*>(op.y) + v * op.y_stride)[n] = lg;
                            if (lg > best || (lg == best && n < best_i)) { best = lg; best_i = n; }
                            break;
                        }
                    }
                }
                if (op.epi == DSE_LOGITS) {
                    // CTA-level argmax candidate (first index wins ties, like torch.argmax)
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                        const int oi2 = __shfl_xor_sync(0xffffffffu, best_i, o);
                        if (ov > best || (ov == best && oi2 < best_i)) { best = ov; best_i = oi2; }
                    }
                    if (lane == 0) { cand_v[warp] = best; cand_i[warp] = best_i; }
                    ds_consumer_sync();
                    if (tid == 0) {
                        for (int w = 1; w < kDsConsumerWarps; ++w)
                            if (cand_v[w] > best || (cand_v[w] == best && cand_i[w] < best_i)) { best = cand_v[w]; best_i = cand_i[w]; }
                        p.cand_val[v * G + cta] = best;
                        p.cand_idx[v * G + cta] = best_i;
                    }
                    ds_consumer_sync();
                }
            }
            stamp(2);
            bar_target += G;
            ds_grid_barrier(p.sync, bar_target, (p.dbg_flags & 2) != 0);
            stamp(3);
        } else if (op.type == DS_ATTN) {
            // ------------------------------------------------------------------ decode attention, split over the KV length
            // Every dependent global access costs ~2 us while the weight stream saturates HBM, so the op is arranged as few
            // dependent rounds as possible: (1) q / new k / new v AND this warp's K rows AND its V rows are requested
            // together; (2) RoPE, scores, softmax, P.V from registers / shared memory; (3) partial + arrival counter;
            // (4) the last CTA of a (lane, kv head) merges all slices with one round of loads.
            constexpr int D = 128, VPL = 4, KB = 16, GM = 4;   // GM query heads per pass (GQA groups of 8: two passes)              // KB keys per warp block: lane pair = key, lane = 4 output dims
            const int Hq = op.Hq, Hk = op.Hk, group = Hq / Hk;
            const int S = max(1, G / Hk);                        // KV slices per (lane, kv head): independent of NV, so a stream's
                                                                 // arithmetic (and ids) do not depend on what it is batched with
            float* sm_q = reinterpret_cast<float*>(xs);           // [group][D] rotated q heads of the group (fp32 values of T)
            float* sm_kn = sm_q + group * D;                      // [D] rotated new k
            float* sm_vn = sm_kn + D;                             // [D] new v
            float* sm_cs = sm_vn + D;                             // [D/2] cos, [D/2] sin
            float* sm_m = sm_cs + D;                              // [8][8]
            float* sm_l = sm_m + kDsGroupWarps * 8;
            float* sm_o = sm_l + kDsGroupWarps * 8;               // [8][group][D]
            float* sm_p = sm_o + kDsGroupWarps * group * D + warp * group * KB;   // this warp's [group][KB] probabilities of a key block
            for (int item = cta; item < NV * Hk * S && !(p.dbg_flags & 4); item += G) {
                const int v = item / (Hk * S), hk = (item / S) % Hk, s = item % S;
                const int pos = p.st[v].pos, kv_len = pos + 1;
                const int slot_kv = p.st[v].kv_slot;
                const T* qkv = reinterpret_cast<const T*>(op.qkv) + v * op.qkv_stride;
                T* kcache = reinterpret_cast<T*>(op.kc) + slot_kv * op.kv_stream_stride + static_cast<long long>(hk) * op.max_ctx * D;
                T* vcache = reinterpret_cast<T*>(op.vc) + slot_kv * op.kv_stream_stride + static_cast<long long>(hk) * op.max_ctx * D;
                const int per = (kv_len + S - 1) / S;
                const int kbeg = s * per, kend = min(kv_len, kbeg + per);
                const int half = lane & 1, kslot = lane >> 1;        // this lane scores dims [64 half, 64 half + 64) of key kb0 + kslot
                // ---- round 1: everything that does not depend on anything else is requested now
                float x1[2], x2[2];                                 // this thread's (head, d) pairs of q / new k before RoPE
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int i = tid + r * kDsConsumerThreads;
                    x1[r] = 0.f; x2[r] = 0.f;
                    if (i < (group + 1) * (D / 2)) {
                        const int hh = i / (D / 2), d = i % (D / 2);
                        const T* src = hh < group ? qkv + (hk * group + hh) * D : qkv + (Hq + hk) * D;
                        x1[r] = Cvt<T>::to_f(ds_ldcg_t(src + d));
                        x2[r] = Cvt<T>::to_f(ds_ldcg_t(src + d + D / 2));
                    }
                }
                float vnew = 0.f;
                if (tid < D) vnew = Cvt<T>::to_f(ds_ldcg_t(qkv + (Hq + Hk + hk) * D + tid));
                uint4 kreg[D / 16];                                 // half a K row (64 dims) of this lane's key
                uint2 vreg[KB];                                     // V (this lane's 4 dims) of the block's keys
                const int kb_first = kbeg + warp * KB;
                auto load_block = [&](int kb0) {
                    const int key = kb0 + kslot;
                    if (key < kend && key != pos) {
                        const uint4* krow = reinterpret_cast<const uint4*>(kcache + static_cast<long long>(key) * D + half * (D / 2));
#pragma unroll
                        for (int c = 0; c < D / 16; ++c) kreg[c] = krow[c];
                    }
#pragma unroll
                    for (int b = 0; b < KB; ++b) {
                        const int kk = kb0 + b;
                        if (kk < kend && kk != pos) vreg[b] = *reinterpret_cast<const uint2*>(vcache + static_cast<long long>(kk) * D + lane * VPL);
                    }
                };
                load_block(kb_first);
                // RoPE tables for this position (hf MistralRotaryEmbedding: fp32 angle, cos / sin cast to T)
                if (tid < D / 2) {
                    const float inv = powf(op.rope_theta, -2.0f * tid / D);
                    float sn, cs;
                    sincosf(pos * inv, &sn, &cs);
                    sm_cs[tid] = rnd<T>(cs);
                    sm_cs[D / 2 + tid] = rnd<T>(sn);
                }
                ds_consumer_sync();
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int i = tid + r * kDsConsumerThreads;
                    if (i < (group + 1) * (D / 2)) {
                        const int hh = i / (D / 2), d = i % (D / 2);
                        const float cs = sm_cs[d], sn = sm_cs[D / 2 + d];
                        float* dst = hh < group ? sm_q + hh * D : sm_kn;
                        dst[d] = rnd<T>(rnd<T>(x1[r] * cs) + rnd<T>(-x2[r] * sn));
                        dst[d + D / 2] = rnd<T>(rnd<T>(x2[r] * cs) + rnd<T>(x1[r] * sn));
                    }
                }
                if (tid < D) sm_vn[tid] = vnew;
                ds_consumer_sync();
                if (kbeg <= pos && pos < kend && tid < D) {          // the slice holding the new position appends it
                    kcache[static_cast<long long>(pos) * D + tid] = Cvt<T>::from_f(sm_kn[tid]);
                    vcache[static_cast<long long>(pos) * D + tid] = Cvt<T>::from_f(sm_vn[tid]);
                }
                for (int gb = 0; gb < group; gb += GM) {
                if (gb > 0) { ds_consumer_sync(); load_block(kb_first); }     // second pass of a wide GQA group re-reads K / V
                // ---- round 2: per warp, blocks of KB keys.  Scores: a lane pair owns a key (64 dims each against the group's
                // q heads broadcast from shared memory, one shuffle per head); block-wise online softmax (one warp_max per
                // head per block); P.V: lane = 4 output dims, probabilities broadcast from shared memory.
                float mx[GM], l[GM], o[GM][VPL];
#pragma unroll
                for (int g = 0; g < GM; ++g) {
                    mx[g] = -INFINITY; l[g] = 0.f;
#pragma unroll
                    for (int i = 0; i < VPL; ++i) o[g][i] = 0.f;
                }
                for (int kb0 = kb_first; kb0 < kend; kb0 += kDsConsumerWarps * KB) {
                    if (kb0 != kb_first) load_block(kb0);     // slices longer than 16 warps x 16 keys (ctx > 4.6k): next round
                    const int key = kb0 + kslot;
                    const bool valid = key < kend;
                    float sc[GM];
#pragma unroll
                    for (int g = 0; g < GM; ++g) sc[g] = 0.f;
                    if (valid && key != pos) {
#pragma unroll
                        for (int c = 0; c < D / 16; ++c) {
                            const uint4 u = kreg[c];
                            const float2 k0 = Cvt<T>::unpack2(u.x), k1 = Cvt<T>::unpack2(u.y), k2 = Cvt<T>::unpack2(u.z), k3 = Cvt<T>::unpack2(u.w);
#pragma unroll
                            for (int g = 0; g < GM; ++g) {
                                if (gb + g < group) {
                                    const float4 qa = *reinterpret_cast<const float4*>(sm_q + (gb + g) * D + half * (D / 2) + c * 8);
                                    const float4 qb = *reinterpret_cast<const float4*>(sm_q + (gb + g) * D + half * (D / 2) + c * 8 + 4);
                                    float t = sc[g];
                                    t = fmaf(qa.x, k0.x, t); t = fmaf(qa.y, k0.y, t); t = fmaf(qa.z, k1.x, t); t = fmaf(qa.w, k1.y, t);
                                    t = fmaf(qb.x, k2.x, t); t = fmaf(qb.y, k2.y, t); t = fmaf(qb.z, k3.x, t); t = fmaf(qb.w, k3.y, t);
                                    sc[g] = t;
                                }
                            }
                        }
                    } else if (valid) {                      // the new token: its k is still in shared memory
                        for (int d = half * (D / 2); d < (half + 1) * (D / 2); ++d) {
                            const float kd = sm_kn[d];
#pragma unroll
                            for (int g = 0; g < GM; ++g)
                                if (gb + g < group) sc[g] = fmaf(sm_q[(gb + g) * D + d], kd, sc[g]);
                        }
                    }
                    float cfac[GM];
#pragma unroll
                    for (int g = 0; g < GM; ++g) {
                        if (gb + g < group) {
                            const float full = sc[g] + __shfl_xor_sync(0xffffffffu, sc[g], 1);   // both halves of the key
                            const float sv = valid ? full * op.scale_log2e : -INFINITY;
                            const float mn = fmaxf(mx[g], warp_max(sv));
                            cfac[g] = exp2f(mx[g] - mn);                              // 0 for the first block (mx = -inf)
                            const float pr = valid ? rnd<T>(exp2f(sv - mn)) : 0.f;    // P is rounded to T before it multiplies V
                            l[g] = l[g] * cfac[g] + (half == 0 ? pr : 0.f);           // lane-local partial of the row sum
                            if (half == 0) sm_p[g * KB + kslot] = pr;
                            mx[g] = mn;
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int g = 0; g < GM; ++g)
                        if (gb + g < group) {
#pragma unroll
                            for (int i = 0; i < VPL; ++i) o[g][i] *= cfac[g];
                        }
#pragma unroll
                    for (int b = 0; b < KB; ++b) {
                        const int kk = kb0 + b;
                        if (kk < kend) {
                            float vv[VPL];
                            if (kk == pos) {
#pragma unroll
                                for (int i = 0; i < VPL; ++i) vv[i] = sm_vn[lane * VPL + i];
                            } else {
                                const float2 a = Cvt<T>::unpack2(vreg[b].x), bb = Cvt<T>::unpack2(vreg[b].y);
                                vv[0] = a.x; vv[1] = a.y; vv[2] = bb.x; vv[3] = bb.y;
                            }
#pragma unroll
                            for (int g = 0; g < GM; ++g)
                                if (gb + g < group) {
                                    const float pr = sm_p[g * KB + b];
#pragma unroll
                                    for (int i = 0; i < VPL; ++i) o[g][i] = fmaf(pr, vv[i], o[g][i]);
                                }
                        }
                    }
                    __syncwarp();
                }
#pragma unroll
                for (int g = 0; g < GM; ++g)
                    if (gb + g < group) l[g] = warp_sum(l[g]);
                // ---- round 3: merge the warps in fixed order (upper half into lower half, then across the 8), write this
                // slice's partial, count the arrival
                const int w8 = warp & (kDsGroupWarps - 1);
                if (kDsGroups > 1 && warp >= kDsGroupWarps) {
#pragma unroll
                    for (int g = 0; g < GM; ++g) {
                        if (gb + g < group) {
                            if (lane == 0) { sm_m[w8 * 8 + g] = mx[g]; sm_l[w8 * 8 + g] = l[g]; }
#pragma unroll
                            for (int i = 0; i < VPL; ++i) sm_o[(w8 * group + g) * D + lane * VPL + i] = o[g][i];
                        }
                    }
                }
                if (kDsGroups > 1) ds_consumer_sync();
                if (warp < kDsGroupWarps) {
#pragma unroll
                    for (int g = 0; g < GM; ++g) {
                        if (kDsGroups > 1 && gb + g < group) {
                            const float pm = sm_m[w8 * 8 + g], pl = sm_l[w8 * 8 + g];
                            const float mn = fmaxf(mx[g], pm);
                            const float c0 = mx[g] == -INFINITY ? 0.f : exp2f(mx[g] - mn), c1 = pm == -INFINITY ? 0.f : exp2f(pm - mn);
                            l[g] = l[g] * c0 + pl * c1;
#pragma unroll
                            for (int i = 0; i < VPL; ++i) o[g][i] = o[g][i] * c0 + sm_o[(w8 * group + g) * D + lane * VPL + i] * c1;
                            mx[g] = mn;
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int g = 0; g < GM; ++g) {
                        if (gb + g < group) {
                            if (lane == 0) { sm_m[w8 * 8 + g] = mx[g]; sm_l[w8 * 8 + g] = l[g]; }
#pragma unroll
                            for (int i = 0; i < VPL; ++i) sm_o[(w8 * group + g) * D + lane * VPL + i] = o[g][i];
                        }
                    }
                }
                ds_consumer_sync();
                float* pbase = p.att_part + (static_cast<long long>(v) * Hq + hk * group) * S * (D + 2);
                const int npass = min(GM, group - gb);
                for (int idx = tid; idx < npass * D; idx += kDsConsumerThreads) {
                    const int g = idx / D, d = idx % D;
                    float mm = -INFINITY;
                    for (int w = 0; w < kDsGroupWarps; ++w) mm = fmaxf(mm, sm_m[w * 8 + g]);
                    float ll = 0.f, oo = 0.f;
                    for (int w = 0; w < kDsGroupWarps; ++w) {
                        const float c = sm_m[w * 8 + g] == -INFINITY ? 0.f : exp2f(sm_m[w * 8 + g] - mm);
                        ll += sm_l[w * 8 + g] * c;
                        oo += sm_o[(w * group + g) * D + d] * c;
                    }
                    float* dst = pbase + (static_cast<long long>(gb + g) * S + s) * (D + 2);
                    dst[2 + d] = oo;
                    if (d == 0) { dst[0] = mm; dst[1] = ll; }
                }
                ds_consumer_sync();
                if (tid == 0) {
                    unsigned* ctr = p.sync + 8 + (gb / GM) * (kDsMaxStreams * Hk) + v * Hk + hk;      // one counter per pass
                    const unsigned old = ds_atom_add_acq_rel(ctr, 1u);     // release this CTA's partial, acquire the others'
                    flag_last = (old == static_cast<unsigned>(S - 1));
                    if (flag_last) *ctr = 0u;
                }
                ds_consumer_sync();
                // ---- round 4: the last CTA of this (lane, kv head) merges the slices in fixed order, all loads in one round
                if (flag_last) {
                    T* att = reinterpret_cast<T*>(op.att) + static_cast<long long>(v) * Hq * D;
                    for (int idx = tid; idx < npass * D; idx += kDsConsumerThreads) {
                        const int g = gb + idx / D, d = idx % D;
                        const float* pp = pbase + static_cast<long long>(g) * S * (D + 2);
                        float mm = -INFINITY, ll = 0.f, oo = 0.f;
                        for (int z0 = 0; z0 < S; z0 += 6) {
                            float pm[6], pl[6], po[6];
#pragma unroll
                            for (int u = 0; u < 6; ++u) {
                                const int z = min(z0 + u, S - 1);
                                pm[u] = __ldcg(pp + z * (D + 2));
                                pl[u] = __ldcg(pp + z * (D + 2) + 1);
                                po[u] = __ldcg(pp + z * (D + 2) + 2 + d);
                            }
#pragma unroll
                            for (int u = 0; u < 6; ++u) {
                                if (z0 + u < S && pm[u] != -INFINITY) {
                                    const float mn = fmaxf(mm, pm[u]);
                                    const float c0 = exp2f(mm - mn), c1 = exp2f(pm[u] - mn);     // c0 = 0 while mm = -inf
                                    ll = ll * c0 + pl[u] * c1;
                                    oo = oo * c0 + po[u] * c1;
                                    mm = mn;
                                }
                            }
                        }
                        att[(hk * group + g) * D + d] = Cvt<T>::from_f(oo / ll);
                    }
                }
                }   // gb
                if (item + G < NV * Hk * S) ds_consumer_sync();      // smem scratch is reused by the next item
            }
            stamp(4);
            bar_target += G;
            ds_grid_barrier(p.sync, bar_target, (p.dbg_flags & 2) != 0);
            stamp(3);
        } else {
            // ------------------------------------------------------------------ DS_FINAL: token selection (CTA 0)
            if (timed) {
#pragma unroll
                for (int c = 0; c < 8; ++c) p.dbg[c] += t_acc[c];
            }
            if (cta == 0 && warp == 0) {
                int all_done = 1;
                for (int v = 0; v < NV; ++v) {
                    float best = -INFINITY;
                    int best_i = 0x7fffffff;
                    for (int c = lane; c < G; c += 32) {
                        const float cv = __ldcg(p.cand_val + v * G + c);
                        const int ci = __ldcg(p.cand_idx + v * G + c);
                        if (cv > best || (cv == best && ci < best_i)) { best = cv; best_i = ci; }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                        const int oi2 = __shfl_xor_sync(0xffffffffu, best_i, o);
                        if (ov > best || (ov == best && oi2 < best_i)) { best = ov; best_i = oi2; }
                    }
                    if (lane == 0) {
                        DsStreamState& st = p.st[v];
                        if (!st.done) {
                            st.tok = best_i;
                            p.out_ids[v * p.out_stride + st.n_out] = best_i;
                            st.n_out += 1;
                            st.pos += 1;
                            const int n_stop = p.stop ? p.stop[0] : 0;
                            for (int z = 0; z < n_stop; ++z)
                                if (p.stop[1 + z] == best_i) st.done = 1;
                            if (st.n_out >= st.max_new) st.done = 1;
                        }
                        all_done &= st.done;
                    }
                }
                if (lane == 0) {
                    __threadfence();
                    p.sync[1] = epoch + 1;
                    if (all_done) p.sync[2] = 1u;
                }
            }
        }
    }
}

// First token of a decode call: greedy argmax (first index wins ties, like torch.argmax) over the fp32 logits the
// prefill of each lane's stream left behind; one CTA walks the lanes.  The token is NOT fed back yet (position
// unchanged), exactly as hf generate() produces token 0 from the prefill logits.
__global__ void __launch_bounds__(1024) ds_first_token_kernel(const float* __restrict__ logits, long long logits_stride, int n,
                                                              int nv, DsStreamState* st, int* out_ids, int out_stride,
                                                              const int* __restrict__ stop, unsigned* sync) {
    __shared__ float sv[32];
    __shared__ int si[32];
    int all_done = 1;
    for (int v = 0; v < nv; ++v) {
        const float* lg = logits + st[v].kv_slot * logits_stride;
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float x = lg[i];
            if (x > best || (x == best && i < bi)) { best = x; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < (blockDim.x >> 5); ++w)
                if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
            st[v].tok = bi;
            out_ids[v * out_stride] = bi;
            st[v].n_out = 1;
            const int n_stop = stop ? stop[0] : 0;
            for (int z = 0; z < n_stop; ++z)
                if (stop[1 + z] == bi) st[v].done = 1;
            if (st[v].max_new <= 1) st[v].done = 1;
            all_done &= st[v].done;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && all_done) sync[2] = 1u;
}

// bytes of the attention scratch that aliases the vector staging region (GQA group g)
inline size_t decode_stream_attn_scratch_bytes(int group) {
    return (static_cast<size_t>(group) * 128 + 3 * 128 + 2 * kDsGroupWarps * 8 + static_cast<size_t>(kDsGroupWarps) * group * 128 +
            static_cast<size_t>(kDsConsumerWarps) * group * 16) * sizeof(float);
}
inline size_t decode_stream_smem_bytes(int n_slots, int x_bytes, int part_cap) {
    return static_cast<size_t>(n_slots) * kDsSlotBytes + static_cast<size_t>(x_bytes) + static_cast<size_t>(part_cap) * sizeof(float);
}

}  // namespace smb
