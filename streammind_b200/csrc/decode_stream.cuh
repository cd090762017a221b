// Persistent weight-streaming kernel for the batch-1 Mistral chains on the path: ONE launch = one greedy decode
// step of the LLM for NV streams (hf MistralForCausalLM.forward at L = 1 per stream + argmax; reference call site:
// streammind/model/language_model/videollama2_mistral.py:426-431).  HBM-bound: 14.22 GB of weights per step.
//
// Why one kernel: as separate launches (7 per layer) every projection pays a prologue (stage + normalise the input
// vector), a pipeline fill and a tail, during which HBM idles: 0.64 of the HBM roofline at ctx 4k (round 1).  Here
// one CTA per SM walks a device-resident op list (GEMV, attention, final argmax); activations travel between the CTAs as
// tagged 8-byte words that the consumer polls (no grid barrier anywhere in the step, see "activation exchange" below),
// and the WEIGHT stream runs ahead: producer warps copy chunks into a shared-memory ring (kDsSlots x 32 KB, cp.async.bulk,
// one mbarrier pair per slot) across op boundaries -- weights do not depend on activations -- so an op starts with up to
// 192 KB / SM already on chip (tools/hbm_read_bench.cu, profiles/r02_hbm_read_bench*.txt: bulk copies of >= 32 KB from
// two issuing threads per SM reach 7.1 TB/s, 16-byte LDG streams 7.3; a thread's bulk copies execute one at a time).
// What bounds a step is measured in profiles/r02_ncu_decode_kernel.md: the weights at the HBM read rate (2.0 ms) plus the
// serial exchange chain (0.7 ms), which the ring does not hide because the consumers are no faster than HBM.
//
// Work split of a GEMV op  y[n] = epi(sum_k W[n, k] * pro(x)[k]):  CTA c owns rows [c * rpc, (c+1) * rpc) of every
// matrix of the op; a ring slot holds R = 8 / P consecutive rows (two matrices: R/2 rows of each, the gate and up rows of
// the same outputs); consumer warp w reduces part w % P (K / P columns) of slot row w / P against the NV staged
// vectors on mma.sync in a diagonal arrangement (see ds_ring_phase), one butterfly per (row, part, vector), partials
// summed in FIXED order by the epilogue thread of the row (deterministic: greedy decode is reproducible run to run).
//
// Attention op (one query per stream): CTA (stream, kv head, KV slice).  The slice's K / V rows arrive through the same ring
// as the weights (128-key chunks copied by the producers); the CTA applies RoPE (cos / sin from a per-launch table) to the
// group's q heads and to the new k, appends k / v to the cache and patches them into the chunk copies, runs online softmax
// over its slice (all query heads of the GQA group share each K / V read), publishes an un-normalised partial; the CTAs
// of a (stream, kv head) share the merge of the slices, in fixed order (no extra launch, no counter, no barrier).
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
// Producer warps of a launch with nv streams (one issuing thread each; a thread's bulk copies execute one at a time, so the copies
// in flight per SM = the issuing threads).  Measured at ctx 2048 (tools/decode_probe.py): one stream 3.05 ms per step with three
// producer warps vs 3.14 with two; two / four streams 4.32 / 6.87 vs 4.05 / 6.27 -- the third warp is worth it for one stream only
// (how many of the warps issue matters less than the block shape: 2 .. 6 issuing threads in the 11-warp block all give 3.07-3.15).
__host__ __device__ constexpr int ds_producer_warps(int nv) { return (nv == 1 && kDsGroups == 1) ? 3 : 2; }
__host__ __device__ constexpr int ds_threads(int nv) { return (kDsConsumerWarps + ds_producer_warps(nv)) * 32; }
constexpr int kDsConsumerThreads = kDsConsumerWarps * 32;
constexpr int kDsSlotBytes = 32 * 1024;
constexpr int kDsMaxSlots = 6;
constexpr int kDsMaxStreams = 4;
#ifndef SMB_DS_WIDE_UB
#define SMB_DS_WIDE_UB 7
#endif
constexpr int kDsKvBlock = 128;             // keys per K / V chunk of the attention op (128 rows of 256 bytes = one ring slot)
constexpr int kDsResidRows = 64;            // rows of the residual stream one CTA may own (hidden <= 64 x SMs)

// -DSMB_DS_WAITPROBE (measurement builds only, tools/gpu/ds_probe_build.sh): producer 0 of every CTA accumulates the time it is blocked on
// a full ring (category 3) and its whole life (5); consumer thread 0 the time it waits for a chunk to land (6 first chunk of an op, 7 later)
#ifdef SMB_DS_WAITPROBE
#define DS_PROBE(...) __VA_ARGS__
#else
#define DS_PROBE(...)
#endif

enum DsOpType : int { DS_GEMV = 0, DS_ATTN = 1, DS_FINAL = 2 };
enum DsPro : int {
    DSP_PLAIN = 0,          // x = x0
    DSP_RMSNORM = 1,        // x = nw * T(x0 * rsqrt(mean(x0^2) + eps))         (hf MistralRMSNorm)
    DSP_EMBED_RMSNORM = 2,  // same with x0 = embed[token of the stream]        (first layer: embed_tokens fused in)
};
enum DsEpi : int {
    DSE_STORE = 0,        // y = T(acc)
    DSE_RESID = 1,        // y = resid[n] = T(resid[n] + T(acc))                 (the residual stream: row n lives in the shared memory of its
                          //                                                      owner CTA for the whole step, starting as embed[token][n])
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
    const unsigned long long* xll;   // [NV][xll_stride] tagged words, two elements each: the input vector as the previous op
    long long xll_stride;            // published it (ignored for DSP_EMBED_RMSNORM)
    const void* nw;        // RMSNorm weight
    float eps;
    unsigned long long* yll;         // [NV][yll_stride] tagged words: what this op publishes (not DSE_LOGITS)
    long long yll_stride;
    float* logits;         // DSE_LOGITS: [NV][logits_stride] fp32
    long long logits_stride;
    // ---- DS_ATTN
    const unsigned long long* qkv_ll;   // [NV][qkv_ll_stride] tagged words: q (Hq*D) | k (Hk*D) | v (Hk*D) of the new token, before RoPE
    long long qkv_ll_stride;
    unsigned long long* att_ll;         // [NV][Hq*D/2] tagged words: the attention output
    void* kc;              // K cache of this layer: [n_streams][Hk][max_ctx][D]
    void* vc;
    long long kv_stream_stride;   // elements between the caches of consecutive streams
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
    unsigned* sync;          // [1] epoch (steps run since the handle was created: never reset, the exchange tags derive from it), [2] all-done flag
    DsStreamState* st;       // [NV]
    int* out_ids;            // [NV][out_stride]
    int out_stride;
    const int* stop;         // [0] = count, then ids
    const void* embed;       // [vocab][H]
    int H;
    unsigned long long* att_part;   // [NV][Hq][S][D + 2] tagged words: split-KV partials (m, l, o[D]) as fp32 bits
    unsigned long long* cand;       // [NV][gridDim.x][2] tagged words: per-CTA argmax candidates (value bits, index)
    int dbg_flags;           // measurement only: 1 skip the consumer math, 2 no weight stream (exchange chain alone), 16 weight chunks from L2 (every chunk re-reads the op's first rows), 32 no copies after the first lap of the ring (the consumers alone)
    long long* dbg;          // optional: CTA 0 accumulates ns per phase (0 prologue incl. waiting for the input, 1 ring compute, 2 epilogue, 4 attention)
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

// ---- activation exchange between CTAs: tagged 8-byte words ("LL" protocol).
// Every value a later op needs from other CTAs travels as ONE naturally aligned 64-bit store: 32 bits of payload (two
// model-dtype elements, or one fp32 / int) + a 32-bit tag that names (decode step, producing op).  A 64-bit scalar access
// is single-copy atomic, so a reader that sees the tag it expects has the payload of that very store: no fence, no
// counter, no grid barrier -- the consumer polls the words it needs and the gap between two ops is one store transit
// plus one L2 read instead of [fence, atomic, poll round trips, then the read].  Why a buffer can be re-used by the same
// op of the next layer without further handshakes: a CTA reads op k's output only in an op in which it also produces
// something that every CTA's next op consumes, so nobody reaches op k + 2 (let alone k + 5) before all readers are done.
__device__ __forceinline__ void ds_ll_store(unsigned long long* p, uint32_t payload, uint32_t tag) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"((static_cast<unsigned long long>(tag) << 32) | payload) : "memory");
}
__device__ __forceinline__ unsigned long long ds_ll_load(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// two adjacent words with one 16-byte request (each 64-bit half is its own atomic access: both tags are checked)
__device__ __forceinline__ void ds_ll_load2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
// poll until the word carries `tag` (r = a first load already issued); bounded, so that a protocol error traps instead of hanging the box
__device__ __forceinline__ uint32_t ds_ll_wait(const unsigned long long* p, unsigned long long r, uint32_t tag) {
    unsigned spins = 0;
    while (static_cast<uint32_t>(r >> 32) != tag) {
        if (++spins > (1u << 21)) { printf("smb: decode exchange timeout (cta %d thread %d, tag %u, found %u)\n", blockIdx.x, threadIdx.x, tag, static_cast<uint32_t>(r >> 32)); __trap(); }
        r = ds_ll_load(p);
    }
    return static_cast<uint32_t>(r);
}
// one round of a batched poll failed: back off briefly (pollers share the L2 ports with the weight stream); bounded
__device__ unsigned ds_poll_ns = 20;
__device__ __forceinline__ void ds_ll_retry(unsigned& spins, uint32_t tag) {
    __nanosleep(ds_poll_ns);
    if (++spins > (1u << 21)) { printf("smb: decode exchange timeout (cta %d thread %d, tag %u)\n", blockIdx.x, threadIdx.x, tag); __trap(); }
}
__device__ __forceinline__ bool ds_ll_ok(unsigned long long r, uint32_t tag) { return static_cast<uint32_t>(r >> 32) == tag; }
__device__ __forceinline__ uint32_t ds_ll_get(const unsigned long long* p, uint32_t tag) { return ds_ll_wait(p, ds_ll_load(p), tag); }
// rows per CTA of a GEMV op: even, so that a CTA's outputs are whole two-element words
__host__ __device__ __forceinline__ int ds_rows_per_cta(int N, int G) { return (((N + G - 1) / G) + 1) & ~1; }

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

struct DsRingState { int seq, slot, par; };   // chunk sequence number of this CTA, its ring slot and the parity of that slot's use

// The attention op of the decode kernel.  Register discipline matters here: if the kernel spills anywhere, ptxas schedules the
// WHOLE kernel for low register use and pairs every shared-memory load of the weight-streaming loop with its MMA (the loop
// then runs at load latency: +30 % per step, measured).  So the projection's words are polled at most two at a time while
// the K / V rows of the first block (64 registers) are in flight.
template <typename T, int NV>
__device__ __forceinline__ void ds_attention_op(const DsOp& op, const DsStreamState* st, unsigned long long* att_part, T* xs,
                                            const float (*rope_cs)[128], uint32_t tag_in, uint32_t tag_out, const uint8_t* ring,
                                            uint64_t* full_bar, uint64_t* empty_bar, int n_slots, DsRingState* state, long long* aprobe) {
    // probe builds (-DSMB_DS_WAITPROBE): thread 0 of CTA 1 accumulates the time between these marks (sm_debug_decode_phases, row 149)
    DS_PROBE(long long ap_t = (aprobe != nullptr) ? ds_gtimer() : 0;)
#define DS_ASTAMP(i) DS_PROBE(if (aprobe != nullptr) { const long long t_ = ds_gtimer(); aprobe[i] += t_ - ap_t; ap_t = t_; })
    const int tid = threadIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x, cta = blockIdx.x;
    // ------------------------------------------------------------------ decode attention, split over the KV length
    // Every dependent global access costs ~2 us while the weight stream saturates HBM (a cache row that misses L2 waits behind
    // ~28 MB of queued weight copies: measured ~6 us per layer when the K / V rows were loaded by the consumers), so the op is
    // arranged as few dependent rounds as possible: (1) the K / V rows of the slice travel through the WEIGHT RING -- the
    // producers copy them, 128 keys per chunk, K then V, between the chunks of the projection before and after this op, i.e.
    // several chunks ahead of the consumers -- and q / new k / new v are polled from the projection's exchange words;
    // (2) RoPE, scores, softmax, P.V from shared memory; (3) the slice's partial is published as tagged words; (4) the CTAs
    // of a (lane, kv head) poll all slices and merge them in fixed order: no counter, no fence, no grid barrier.
    constexpr int D = 128, VPL = 4, KB = 16, GM = 4;   // GM query heads per pass (GQA groups of 8: two passes)              // KB keys per warp block: lane pair = key, lane = 4 output dims
    const int Hq = op.Hq, Hk = op.Hk, group = Hq / Hk;
    const int S = max(1, G / Hk);                        // KV slices per (lane, kv head): independent of NV, so a stream's
                                                         // arithmetic (and ids) do not depend on what it is batched with
    constexpr int QH = D / 2 + 4, QS = 2 * QH;            // q head = two halves of 64 + 4 pad floats: the two lanes of a key read different banks
    float* sm_q = reinterpret_cast<float*>(xs);           // [group][QS] rotated q heads of the group (fp32 values of T)
    float* sm_kn = sm_q + group * QS;                     // [D] rotated new k
    float* sm_vn = sm_kn + D;                             // [D] new v
    float* sm_cs = sm_vn + D;                             // [D/2] cos, [D/2] sin
    float* sm_m = sm_cs + D;                              // [8][8]
    float* sm_l = sm_m + kDsGroupWarps * 8;
    float* sm_o = sm_l + kDsGroupWarps * 8;               // [8][group][D]
    float* sm_p = sm_o + kDsGroupWarps * group * D + warp * max(group, GM) * KB;   // this warp's [KB][GM] probabilities of a key block (one pass)
    DsRingState rs = *state;
    for (int item = cta; item < NV * Hk * S; item += G) {
        const int v = item / (Hk * S), hk = (item / S) % Hk, s = item % S;
        const int pos = st[v].pos, kv_len = pos + 1;
        const int slot_kv = st[v].kv_slot;
        const unsigned long long* ql = op.qkv_ll + v * op.qkv_ll_stride;
        T* kcache = reinterpret_cast<T*>(op.kc) + slot_kv * op.kv_stream_stride + static_cast<long long>(hk) * op.max_ctx * D;
        T* vcache = reinterpret_cast<T*>(op.vc) + slot_kv * op.kv_stream_stride + static_cast<long long>(hk) * op.max_ctx * D;
        const int per = (kv_len + S - 1) / S;
        const int kbeg = s * per, kend = min(kv_len, kbeg + per);
        const int half = lane & 1, kslot = lane >> 1;        // this lane scores dims [64 half, 64 half + 64) of key kb0 + kslot
        const int nblk = (max(0, kend - kbeg) + kDsKvBlock - 1) / kDsKvBlock;   // 128-key blocks of the slice = (K, V) chunk pairs in the ring, per pass
        // Poll the projection's words of this kv head and rotate them in the same thread: a job is two adjacent dims d, d + 1
        // (d < D/2) of a q head / the new k -- the word holding them and the word holding d + D/2, d + D/2 + 1 -- or one word
        // of the new v.  cos / sin of the position come from the step's table (rope_cs, filled once per launch).
        for (int jb = tid; jb < (group + 1) * (D / 4) + D / 2; jb += kDsConsumerThreads) {
            if (jb < (group + 1) * (D / 4)) {
                const int hh = jb / (D / 4), d = 2 * (jb % (D / 4));
                const unsigned long long* wa = ql + (((hh < group ? (hk * group + hh) * D : (Hq + hk) * D) + d) >> 1);
                unsigned long long ra = ds_ll_load(wa), rb = ds_ll_load(wa + D / 4);
                const float cs0 = rope_cs[v][d], sn0 = rope_cs[v][D / 2 + d], cs1 = rope_cs[v][d + 1], sn1 = rope_cs[v][D / 2 + d + 1];
                unsigned spins = 0;
                while (!(ds_ll_ok(ra, tag_in) && ds_ll_ok(rb, tag_in))) { ds_ll_retry(spins, tag_in); ra = ds_ll_load(wa); rb = ds_ll_load(wa + D / 4); }
                const float2 x1 = Cvt<T>::unpack2(static_cast<uint32_t>(ra)), x2 = Cvt<T>::unpack2(static_cast<uint32_t>(rb));
                float* dst = hh < group ? sm_q + hh * QS : sm_kn;
                const int h2 = hh < group ? QH : D / 2;          // offset of the second half of the row
                dst[d] = rnd<T>(rnd<T>(x1.x * cs0) + rnd<T>(-x2.x * sn0));
                dst[d + 1] = rnd<T>(rnd<T>(x1.y * cs1) + rnd<T>(-x2.y * sn1));
                dst[h2 + d] = rnd<T>(rnd<T>(x2.x * cs0) + rnd<T>(x1.x * sn0));
                dst[h2 + d + 1] = rnd<T>(rnd<T>(x2.y * cs1) + rnd<T>(x1.y * sn1));
            } else {
                const int w = jb - (group + 1) * (D / 4);
                const float2 f = Cvt<T>::unpack2(ds_ll_get(ql + ((Hq + Hk + hk) * D) / 2 + w, tag_in));
                sm_vn[2 * w] = f.x;
                sm_vn[2 * w + 1] = f.y;
            }
        }
        ds_consumer_sync();
        DS_ASTAMP(0)      // q / k / v of the projection arrived and rotated
        if (kbeg <= pos && pos < kend && tid < D) {          // the slice holding the new position appends it
            kcache[static_cast<long long>(pos) * D + tid] = Cvt<T>::from_f(sm_kn[tid]);
            vcache[static_cast<long long>(pos) * D + tid] = Cvt<T>::from_f(sm_vn[tid]);
        }
        for (int gb = 0; gb < group; gb += GM) {
        if (gb > 0) ds_consumer_sync();      // second pass of a wide GQA group: the producers stream K / V once more
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
        for (int blk = 0; blk < nblk; ++blk) {
            // chunk pair of this block: K rows then V rows of keys [kbeg + 128 blk, + 128), one 256-byte row per key
            const int slot_k = rs.slot, par_k = rs.par;
            if (++rs.slot == n_slots) { rs.slot = 0; rs.par ^= 1; }
            const int slot_v = rs.slot, par_v = rs.par;
            if (++rs.slot == n_slots) { rs.slot = 0; rs.par ^= 1; }
            rs.seq += 2;
            const uint8_t* sk = ring + static_cast<size_t>(slot_k) * kDsSlotBytes + (warp * KB) * (D * sizeof(T));
            const uint8_t* sv = ring + static_cast<size_t>(slot_v) * kDsSlotBytes + (warp * KB) * (D * sizeof(T));
            const int kb0 = kbeg + blk * kDsKvBlock + warp * KB;          // this warp's 16 keys of the block
            if (kDsGroups == 1 || warp < kDsGroupWarps) mbar_wait_hint(&full_bar[slot_k], par_k);
            {   // The block that holds the new position (CTA-uniform): its cache row is not written yet -- the rotated k and the v
                // of this step are patched into the shared-memory copies of the chunks, so the loops below treat every key alike.
                const int blk0 = kbeg + blk * kDsKvBlock;
                if (pos >= blk0 && pos < min(kend, blk0 + kDsKvBlock)) {
                    if (kDsGroups > 1 && warp >= kDsGroupWarps) mbar_wait_hint(&full_bar[slot_k], par_k);
                    mbar_wait_hint(&full_bar[slot_v], par_v);
                    if (tid < D) {
                        const size_t row = static_cast<size_t>(pos - blk0) * D;
                        reinterpret_cast<T*>(const_cast<uint8_t*>(ring) + static_cast<size_t>(slot_k) * kDsSlotBytes)[row + tid] = Cvt<T>::from_f(sm_kn[tid]);
                        reinterpret_cast<T*>(const_cast<uint8_t*>(ring) + static_cast<size_t>(slot_v) * kDsSlotBytes)[row + tid] = Cvt<T>::from_f(sm_vn[tid]);
                        fence_proxy_async_smem();      // generic-proxy writes to a slot that the next bulk copy (async proxy) overwrites
                    }
                    ds_consumer_sync();
                }
            }
            if (kDsGroups > 1 && warp >= kDsGroupWarps) continue;      // a second consumer group (experiment builds) has no part in the key blocks
            if (kb0 < kend) {
            const int key = kb0 + kslot;
            const bool valid = key < kend;
            float sc[GM];
#pragma unroll
            for (int g = 0; g < GM; ++g) sc[g] = 0.f;
            if (valid) {
#pragma unroll
                for (int c = 0; c < D / 16; ++c) {
                    // 16-byte pieces of the half row in an order rotated by the key (4 rotations): rows are 256 bytes apart, so the
                    // lanes would otherwise all read the same banks; with 4 rotations and the padded q layout a K read is 8 wavefronts
                    // and a q read 1 (8 rotations: 4 and 4 per q read, of which there are eight per K read)
                    const int cc = (c + (kslot & 3)) & (D / 16 - 1);
                    const uint4 u = *reinterpret_cast<const uint4*>(sk + kslot * (D * sizeof(T)) + half * (D / 2) * sizeof(T) + cc * 16);
                    const float2 k0 = Cvt<T>::unpack2(u.x), k1 = Cvt<T>::unpack2(u.y), k2 = Cvt<T>::unpack2(u.z), k3 = Cvt<T>::unpack2(u.w);
#pragma unroll
                    for (int g = 0; g < GM; ++g) {
                        if (gb + g < group) {
                            const float4 qa = *reinterpret_cast<const float4*>(sm_q + (gb + g) * QS + half * QH + cc * 8);
                            const float4 qb = *reinterpret_cast<const float4*>(sm_q + (gb + g) * QS + half * QH + cc * 8 + 4);
                            float t = sc[g];
                            t = fmaf(qa.x, k0.x, t); t = fmaf(qa.y, k0.y, t); t = fmaf(qa.z, k1.x, t); t = fmaf(qa.w, k1.y, t);
                            t = fmaf(qb.x, k2.x, t); t = fmaf(qb.y, k2.y, t); t = fmaf(qb.z, k3.x, t); t = fmaf(qb.w, k3.y, t);
                            sc[g] = t;
                        }
                    }
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
                    if (half == 0) sm_p[kslot * GM + g] = pr;
                    mx[g] = mn;
                }
            }
            __syncwarp();
            mbar_wait_hint(&full_bar[slot_v], par_v);
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
                    const uint2 vr = *reinterpret_cast<const uint2*>(sv + b * (D * sizeof(T)) + lane * (VPL * sizeof(T)));
                    const float2 a = Cvt<T>::unpack2(vr.x), bb = Cvt<T>::unpack2(vr.y);
                    vv[0] = a.x; vv[1] = a.y; vv[2] = bb.x; vv[3] = bb.y;
                    const float4 p4 = *reinterpret_cast<const float4*>(sm_p + b * GM);      // the key's probabilities for the pass's heads
                    const float prs[GM] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
                    for (int g = 0; g < GM; ++g)
                        if (gb + g < group) {
#pragma unroll
                            for (int i = 0; i < VPL; ++i) o[g][i] = fmaf(prs[g], vv[i], o[g][i]);
                        }
                }
            }
            } else {
                mbar_wait_hint(&full_bar[slot_v], par_v);      // nothing of this block for this warp: it still releases the chunks
            }
            __syncwarp();
            if (lane == 0) { mbar_arrive(&empty_bar[slot_k]); mbar_arrive(&empty_bar[slot_v]); }
        }
        DS_ASTAMP(4)      // key blocks: scores, softmax, P V
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
        DS_ASTAMP(5)      // row sums, per-warp partials in shared memory
        unsigned long long* pbase = att_part + (static_cast<long long>(v) * Hq + hk * group) * S * (D + 2);
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
            unsigned long long* dst = pbase + (static_cast<long long>(gb + g) * S + s) * (D + 2);
            ds_ll_store(dst + 2 + d, __float_as_uint(oo), tag_out);
            if (d == 0) { ds_ll_store(dst, __float_as_uint(mm), tag_out); ds_ll_store(dst + 1, __float_as_uint(ll), tag_out); }
        }
        ds_consumer_sync();      // the scratch is rewritten by the next pass / item / op
        DS_ASTAMP(1)      // warps merged, partial published
        }   // gb
    }
    *state = rs;
    // ---- round 4: the S CTAs of a (lane, kv head) share the merge: CTA s takes output pairs [s cnt, (s + 1) cnt) of the
    // group's heads, polls their partials of ALL slices in one round (a thread per (pair, slice)), then one thread
    // per pair folds the slices in fixed order and publishes the word
    {
        const int npair = group * (D / 2), cnt = (npair + S - 1) / S;
        float4* sm_mg = reinterpret_cast<float4*>(sm_o);                 // [cnt][S] (m, l, o0, o1)
        for (int item = cta; item < NV * Hk * S; item += G) {
            const int v = item / (Hk * S), hk = (item / S) % Hk, s = item % S;
            const int p0 = s * cnt, p1 = min(npair, p0 + cnt);
            const unsigned long long* pbase = att_part + (static_cast<long long>(v) * Hq + hk * group) * S * (D + 2);
            constexpr int UM = 1;      // (pairs x slices) of a CTA can exceed the 256 threads by a few: both polls of a thread in flight together
            for (int i0 = tid; i0 < (p1 - p0) * S; i0 += UM * kDsConsumerThreads) {
                unsigned long long r[UM][4];
                const unsigned long long* q[UM];
                int dp2[UM];
#pragma unroll
                for (int u = 0; u < UM; ++u) {
                    const int i = i0 + u * kDsConsumerThreads;
                    q[u] = pbase; dp2[u] = 0;
                    if (i < (p1 - p0) * S) {
                        const int pr = p0 + i / S, z = i % S, g = pr / (D / 2);
                        dp2[u] = 2 * (pr % (D / 2));
                        q[u] = pbase + (static_cast<long long>(g) * S + z) * (D + 2);
                    }
                }
                unsigned spins = 0;
                for (bool ok = false; !ok;) {
#pragma unroll
                    for (int u = 0; u < UM; ++u)
                        if (i0 + u * kDsConsumerThreads < (p1 - p0) * S) {
                            ds_ll_load2(q[u], r[u][0], r[u][1]);
                            ds_ll_load2(q[u] + 2 + dp2[u], r[u][2], r[u][3]);
                        }
                    ok = true;
#pragma unroll
                    for (int u = 0; u < UM; ++u)
                        if (i0 + u * kDsConsumerThreads < (p1 - p0) * S)
                            ok = ok && ds_ll_ok(r[u][0], tag_out) && ds_ll_ok(r[u][1], tag_out) && ds_ll_ok(r[u][2], tag_out) && ds_ll_ok(r[u][3], tag_out);
                    if (!ok) ds_ll_retry(spins, tag_out);
                }
#pragma unroll
                for (int u = 0; u < UM; ++u) {
                    const int i = i0 + u * kDsConsumerThreads;
                    if (i < (p1 - p0) * S)
                        sm_mg[i] = make_float4(__uint_as_float(static_cast<uint32_t>(r[u][0])), __uint_as_float(static_cast<uint32_t>(r[u][1])),
                                               __uint_as_float(static_cast<uint32_t>(r[u][2])), __uint_as_float(static_cast<uint32_t>(r[u][3])));
                }
            }
            ds_consumer_sync();
            DS_ASTAMP(2)      // partials of all slices polled
            if (tid < p1 - p0) {
                const int pr = p0 + tid, g = pr / (D / 2), dp = pr % (D / 2);
                // two passes (the maximum over the slices first, then independent weights 2^(m_z - M) summed in slice order): a
                // chain of S dependent rescales would put ~1.5 us per layer on the step's critical path
                float mm = -INFINITY, ll = 0.f, o0 = 0.f, o1 = 0.f;
                for (int z = 0; z < S; ++z) mm = fmaxf(mm, sm_mg[tid * S + z].x);
                for (int z = 0; z < S; ++z) {
                    const float4 t = sm_mg[tid * S + z];
                    const float c1 = t.x == -INFINITY ? 0.f : exp2f(t.x - mm);
                    ll = fmaf(t.y, c1, ll);
                    o0 = fmaf(t.z, c1, o0);
                    o1 = fmaf(t.w, c1, o1);
                }
                ds_ll_store(op.att_ll + static_cast<long long>(v) * (Hq * D / 2) + (hk * group + g) * (D / 2) + dp, Cvt<T>::pack2(o0 / ll, o1 / ll), tag_out);
            }
            ds_consumer_sync();      // the scratch is rewritten by the next item / op
            DS_ASTAMP(3)      // slices folded, output published
        }
    }
#undef DS_ASTAMP
}


// The weight-streaming phase of a GEMV op.
template <typename T, int NV>
__device__ __forceinline__ void ds_ring_phase(const DsOp& op, uint64_t* full_bar, uint64_t* empty_bar, int n_slots, int x_bytes, int xcap,
                                          int dbg_flags, int j0, int j1, DsRingState* state, long long* probe) {
    extern __shared__ __align__(128) uint8_t ds_smem[];            // same carve-up as the kernel: ring | staged vectors | partial sums
    const uint8_t* ring = ds_smem;
    const T* xs = reinterpret_cast<const T*>(ds_smem + static_cast<size_t>(n_slots) * kDsSlotBytes);
    float* part = reinterpret_cast<float*>(ds_smem + static_cast<size_t>(n_slots) * kDsSlotBytes + x_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = op.K, nloc = j1 - j0;
    DsRingState rs = *state;
    // ---- stream this CTA's rows through the ring
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
    const int xq = xcap / 4;
    // The hot loop is kept branch-free and division-free (ncu, profiles/r02_decode_kernel.md: an earlier form with a
    // guard per block executed ~3700 warp instructions per 32 KB slot, 22 % of them mbarrier polls): blocks go in
    // unguarded batches of 8 and 2 (nblk is even), the ring position advances incrementally.
    const uint2* ring_u2 = reinterpret_cast<const uint2*>(ring) + (static_cast<size_t>(sr) * K + c0) / 4 + lane;
    float* part_w = part + (static_cast<size_t>(m) * nloc * P + pt) * NV;           // + local row * P * NV
    const int skip_math = dbg_flags & 1;
    if (dbg_flags & 2) return;
    for (int j = j0; j < j1; j += RJ, ++rs.seq) {
        const int slot = rs.slot, par = rs.par;
        if (++rs.slot == n_slots) { rs.slot = 0; rs.par ^= 1; }
        if (kDsGroups > 1 && rs.seq % kDsGroups != grp) continue;
        DS_PROBE(const long long tw0 = (probe != nullptr) ? ds_gtimer() : 0;)
        mbar_wait_hint(&full_bar[slot], par);
        DS_PROBE(if (probe != nullptr) probe[j == j0 ? 0 : 1] += ds_gtimer() - tw0;)
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
                for (int u = 0; u < 8; ++u) w[u] = *(wp + (b + u) * 32);
#pragma unroll
                for (int q = 0; q < NPAIR; ++q) {
                    uint2 xa[8], xb[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        xa[u] = *(xp + (2 * q) * xq + (b + u) * 32);
                        xb[u] = 2 * q + 1 < NV ? *(xp + (2 * q + 1) * xq + (b + u) * 32) : make_uint2(0u, 0u);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) ds_mma_16816<T>(acc[q][u & 1], xa[u].x, xb[u].x, xa[u].y, xb[u].y, w[u].x, w[u].y);
                }
            }
            for (; b < nblk; b += 2) {
                const uint2 w0 = *(wp + b * 32), w1 = *(wp + b * 32 + 32);
#pragma unroll
                for (int q = 0; q < NPAIR; ++q) {
                    const uint2 xa0 = *(xp + (2 * q) * xq + b * 32), xa1 = *(xp + (2 * q) * xq + b * 32 + 32);
                    uint2 xb0 = make_uint2(0u, 0u), xb1 = make_uint2(0u, 0u);
                    if (2 * q + 1 < NV) { xb0 = *(xp + (2 * q + 1) * xq + b * 32); xb1 = *(xp + (2 * q + 1) * xq + b * 32 + 32); }
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
    *state = rs;
}

template <typename T, int NV>
__global__ void __launch_bounds__(ds_threads(NV), 1) decode_stream_kernel(const DsParams p) {
    extern __shared__ __align__(128) uint8_t ds_smem[];
    uint8_t* ring = ds_smem;                                                       // [n_slots][32 KB]
    T* xs = reinterpret_cast<T*>(ds_smem + static_cast<size_t>(p.n_slots) * kDsSlotBytes);   // [NV][xcap]; attention scratch aliases it
    float* part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(xs) + p.x_bytes);
    __shared__ uint64_t full_bar[kDsMaxSlots], empty_bar[kDsMaxSlots];
    __shared__ float red[kDsConsumerWarps * NV];
    __shared__ float cand_v[kDsConsumerWarps];
    __shared__ int cand_i[kDsConsumerWarps];
    __shared__ float resid_own[NV][kDsResidRows];      // this CTA's rows of the residual stream (values of T)
    __shared__ float rope_cs[NV][128];                 // cos [0, 64) and sin [64, 128) of each stream's position (values of T)

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
        // Each producer walks the chunk sequence of this CTA (chunks seq % (producer warps) == its index) and copies chunk seq into
        // ring slot seq % n_slots as soon as the consumers have released it.  The sequence follows the op list: the row chunks
        // of a GEMV op; for an attention op the K rows and the V rows of every 128-key block of this CTA's KV slice(s) -- the cache
        // rows reach shared memory through the same pipelined stream as the weights, requested while the projection before the
        // attention is still being reduced.  L2 prefetching of later chunks was measured and rejected twice: a cursor at a fixed
        // distance of 8 / 16 chunks (3.65 / 5.4 ms per step vs 3.32) and a cursor that only runs while the ring is full, i.e. while
        // the consumers sit in an exchange poll (3.17 / 3.25 / 3.87 / 4.51 ms at <= 4 / 8 / 12 / 16 chunks vs 3.12).
        if (lane != 0) return;
        constexpr int np = ds_producer_warps(NV);
        const int pi = warp - kDsConsumerWarps;
        int seq = 0;
        DS_PROBE(long long pb_blocked = 0; const long long pb_t0 = ds_gtimer();)
        auto put = [&](const void* s0, const void* s1, uint32_t bytes, uint32_t off1) {       // chunk seq: one or two copies into its slot
            if (seq % np == pi) {
                const int slot = seq % p.n_slots, use = seq / p.n_slots;
                if (use > 0) {
                    DS_PROBE(const long long tb = ds_gtimer();)
                    mbar_wait_hint(&empty_bar[slot], (use - 1) & 1);
                    DS_PROBE(pb_blocked += ds_gtimer() - tb;)
                }
                if ((p.dbg_flags & 32) && use > 0) {      // measurement only: no copy after the first lap, the slot is declared full as it is
                    mbar_arrive(&full_bar[slot]);
                } else {
                    mbar_arrive_expect_tx(&full_bar[slot], s1 ? 2 * bytes : bytes);
                    uint8_t* dst = ring + static_cast<size_t>(slot) * kDsSlotBytes;
                    ds_bulk_g2s(dst, s0, bytes, &full_bar[slot]);
                    if (s1) ds_bulk_g2s(dst + off1, s1, bytes, &full_bar[slot]);
                }
            }
            ++seq;
        };
        for (int oi = 0; oi < p.n_ops; ++oi) {
            const DsOp& op = p.ops[oi];
            if (op.type == DS_GEMV) {
                if (p.dbg_flags & 2) continue;          // measurement only: no weight stream (the exchange chain alone)
                const int rpc = ds_rows_per_cta(op.N, G);
                const int j0 = min(op.N, cta * rpc), j1 = min(op.N, j0 + rpc), RJ = op.R / op.nmat;
                const size_t row_bytes = static_cast<size_t>(op.K) * sizeof(T);
                for (int j = j0; j < j1; j += RJ) {
                    const uint32_t bytes = static_cast<uint32_t>(min(RJ, j1 - j) * row_bytes);
                    const int js = (p.dbg_flags & 16) ? j0 : j;      // measurement only: every chunk re-reads the op's first rows (L2 hits: the consumers' own speed)
                    put(static_cast<const uint8_t*>(op.W0) + js * row_bytes,
                        op.nmat == 2 ? static_cast<const uint8_t*>(op.W1) + js * row_bytes : nullptr, bytes, static_cast<uint32_t>(RJ * row_bytes));
                }
            } else if (op.type == DS_ATTN) {
                const int Hk = op.Hk, group = op.Hq / Hk, S = max(1, G / Hk), npass = (group + 3) / 4;
                for (int item = cta; item < NV * Hk * S; item += G) {
                    const int v = item / (Hk * S), hk = (item / S) % Hk, s = item % S;
                    const int kv_len = p.st[v].pos + 1, per = (kv_len + S - 1) / S;
                    const int kbeg = s * per, kend = min(kv_len, kbeg + per);
                    const long long base = p.st[v].kv_slot * op.kv_stream_stride + static_cast<long long>(hk) * op.max_ctx * 128;
                    for (int pass = 0; pass < npass; ++pass)
                        for (int kb = kbeg; kb < kend; kb += kDsKvBlock) {
                            const uint32_t bytes = static_cast<uint32_t>(min(kDsKvBlock, kend - kb)) * 128u * static_cast<uint32_t>(sizeof(T));
                            put(reinterpret_cast<const T*>(op.kc) + base + static_cast<long long>(kb) * 128, nullptr, bytes, 0u);
                            put(reinterpret_cast<const T*>(op.vc) + base + static_cast<long long>(kb) * 128, nullptr, bytes, 0u);
                        }
                }
            }
        }
        DS_PROBE(if (p.dbg != nullptr && pi == 0) { p.dbg[8 * (cta + 1) + 3] += pb_blocked; p.dbg[8 * (cta + 1) + 5] += ds_gtimer() - pb_t0; if (cta == 0) { p.dbg[3] += pb_blocked; p.dbg[5] += ds_gtimer() - pb_t0; } })
        return;
    }

    // ====================================================================== consumers
    const int tid = threadIdx.x;                      // 0 .. kDsConsumerThreads-1
    const bool timed = p.dbg != nullptr && tid == 0;        // every CTA keeps its own phase clock: rows 8 (cta + 1) .. of the buffer; CTA 0 also rows 0..7
    // phase clocks: 32-bit ns, accumulated in registers (a global read-modify-write per stamp would dominate); categories 0, 1, 2, 4
    unsigned t_last = timed ? static_cast<unsigned>(ds_gtimer()) : 0u;
    unsigned t_acc[4] = {0u, 0u, 0u, 0u};
    auto stamp = [&](int cat) {
        if (timed) {
            const unsigned t = static_cast<unsigned>(ds_gtimer());
            const int ci = cat == 4 ? 3 : cat;
#pragma unroll
            for (int c = 0; c < 4; ++c) t_acc[c] += c == ci ? t - t_last : 0u;
            t_last = t;
        }
    };
    {   // the residual stream starts as the embedding of the token fed (hf MistralModel: inputs_embeds = embed_tokens(ids))
        const int rpcH = ds_rows_per_cta(p.H, G), h0 = min(p.H, cta * rpcH), hn = min(p.H, h0 + rpcH) - h0;
        for (int i = tid; i < NV * hn; i += kDsConsumerThreads) {
            const int v = i / hn, r = i - v * hn;
            resid_own[v][r] = Cvt<T>::to_f((reinterpret_cast<const T*>(p.embed) + static_cast<size_t>(p.st[v].tok) * p.H)[h0 + r]);
        }
    }
    {   // cos / sin of every stream's position, once per step instead of once per layer (hf MistralRotaryEmbedding: fp32 angle
        // pos * theta^(-2 d / D), cos / sin cast to T); first read after the consumer barriers of the first GEMV op
        float theta = 0.f;
        for (int oi = 0; oi < p.n_ops; ++oi)
            if (p.ops[oi].type == DS_ATTN) { theta = p.ops[oi].rope_theta; break; }
        for (int i = tid; i < NV * 64; i += kDsConsumerThreads) {
            const int v = i >> 6, d = i & 63;
            float sn, cs;
            sincosf(p.st[v].pos * powf(theta, -2.0f * d / 128), &sn, &cs);
            rope_cs[v][d] = rnd<T>(cs);
            rope_cs[v][64 + d] = rnd<T>(sn);
        }
    }
    // tag of what op oi publishes in this step (never 0: a buffer that was never written cannot match)
    const uint32_t tag0 = epoch * static_cast<uint32_t>(p.n_ops + 1) + 1u;
    auto tag_of = [&](int oi) { return (tag0 + static_cast<uint32_t>(oi)) | 0x80000000u; };
    DsRingState rs{0, 0, 0};
    long long probe_acc[2] = {0, 0};
    long long* probe_w = nullptr;
    long long aprobe_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};       // attention sub-phases (probe builds)
    long long* aprobe_w = nullptr;
    DS_PROBE(if (timed) { probe_w = probe_acc; aprobe_w = aprobe_acc; })
    for (int oi = 0; oi < p.n_ops; ++oi) {
        const DsOp& op = p.ops[oi];
        if (op.type == DS_GEMV) {
            const int K = op.K;
            const int rpc = ds_rows_per_cta(op.N, G);
            const int j0 = min(op.N, cta * rpc), j1 = min(op.N, j0 + rpc);
            const int nloc = j1 - j0;
            if (nloc == 0) continue;          // no rows of this op: nothing to read, nothing to publish (the producers skip it too)
            const uint32_t tag_in = tag_of(oi - 1), tag_out = tag_of(oi);
            // the RMSNorm weights of this thread's elements are requested BEFORE the input is polled: loaded in the second pass they put an
            // L2 round trip on the chain of every normalising op (64 per step)
            uint4 nw_pre[2] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)};
            const bool nw_fit = op.pro != DSP_PLAIN && K <= 2 * kDsConsumerThreads * 8;
            if (nw_fit) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int k = (tid + i * kDsConsumerThreads) * 8;
                    if (k < K) nw_pre[i] = *reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(op.nw) + k);
                }
            }
            // ---- prologue: stage pro(x_v) for every stream; x_v is polled word by word from the previous op's exchange buffer
#pragma unroll 1
            for (int v = 0; v < NV; ++v) {
                T* xv = xs + static_cast<size_t>(v) * p.xcap;
                float s2 = 0.f;
                if (op.pro == DSP_EMBED_RMSNORM) {
                    const T* x0 = reinterpret_cast<const T*>(p.embed) + static_cast<size_t>(p.st[v].tok) * p.H;
                    for (int k = tid * 8; k < K; k += kDsConsumerThreads * 8) {
                        const uint4 u = *reinterpret_cast<const uint4*>(x0 + k);
                        *reinterpret_cast<uint4*>(xv + k) = u;
                        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) { const float2 f = Cvt<T>::unpack2(w[i]); s2 = fmaf(f.x, f.x, s2); s2 = fmaf(f.y, f.y, s2); }
                    }
                } else {
                    const unsigned long long* xl = op.xll + v * op.xll_stride;
                    uint2* xw = reinterpret_cast<uint2*>(xv);
                    const int nq = K / 4;                                               // word pairs = 4 elements
                    auto poll_batches = [&](auto ub_tag) {
                    constexpr int UB = decltype(ub_tag)::value;                         // 16-byte polls in flight per thread
                    for (int q0 = tid; q0 < nq; q0 += UB * kDsConsumerThreads) {
                        unsigned long long r[UB][2];
                        unsigned spins = 0;
                        for (bool ok = false; !ok;) {           // the whole batch is polled again until every tag matches (one round trip per try)
#pragma unroll
                            for (int u = 0; u < UB; ++u) {
                                const int q = q0 + u * kDsConsumerThreads;
                                if (q < nq) ds_ll_load2(xl + 2 * q, r[u][0], r[u][1]);
                            }
                            ok = true;
#pragma unroll
                            for (int u = 0; u < UB; ++u) {
                                const int q = q0 + u * kDsConsumerThreads;
                                if (q < nq) ok = ok && ds_ll_ok(r[u][0], tag_in) && ds_ll_ok(r[u][1], tag_in);
                            }
                            if (!ok) ds_ll_retry(spins, tag_in);
                        }
#pragma unroll
                        for (int u = 0; u < UB; ++u) {
                            const int q = q0 + u * kDsConsumerThreads;
                            if (q < nq) {
                                const uint32_t d0 = static_cast<uint32_t>(r[u][0]), d1 = static_cast<uint32_t>(r[u][1]);
                                xw[q] = make_uint2(d0, d1);
                                const float2 f = Cvt<T>::unpack2(d0), g = Cvt<T>::unpack2(d1);
                                s2 = fmaf(f.x, f.x, s2); s2 = fmaf(f.y, f.y, s2); s2 = fmaf(g.x, g.x, s2); s2 = fmaf(g.y, g.y, s2);
                            }
                        }
                    }
                    };
                    // every dependent poll round costs an L2 round trip: a vector wider than 4 polls per thread (down_proj's 14 336
                    // inputs = 14 per thread) is polled as two batches of 7 instead of four batches of 4 in a row (one batch of 14: same chain time, slower ring loop)
                    if (nq > 4 * kDsConsumerThreads) poll_batches(std::integral_constant<int, SMB_DS_WIDE_UB>{});
                    else poll_batches(std::integral_constant<int, 4>{});
                }
                if (op.pro != DSP_PLAIN) {
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
#pragma unroll 2
                    for (int k = tid * 8, i = 0; k < K; k += kDsConsumerThreads * 8, ++i) {
                        const uint4 u = *reinterpret_cast<const uint4*>(xv + k);
                        const uint4 g = nw_fit ? (i == 0 ? nw_pre[0] : nw_pre[1]) : *reinterpret_cast<const uint4*>(nw + k);
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
            const int P = op.P;
            ds_ring_phase<T, NV>(op, full_bar, empty_bar, p.n_slots, p.x_bytes, p.xcap, p.dbg_flags, j0, j1, &rs, probe_w);
            ds_consumer_sync();
            stamp(1);
            // ---- epilogue: one thread per output row (fixed-order sum of the parts); neighbouring lanes pair their rows into one
            // exchange word
            float best = -INFINITY;
            int best_i = 0x7fffffff;
#pragma unroll 1
            for (int v = 0; v < NV; ++v) {
                if (op.epi == DSE_LOGITS) { best = -INFINITY; best_i = 0x7fffffff; }
                for (int rb = 0; rb < nloc; rb += kDsConsumerThreads) {
                    const int r = rb + tid, n = j0 + r;
                    float out = 0.f;
                    if (r < nloc) {
                        float a0 = 0.f, a1 = 0.f;
                        for (int q = 0; q < P; ++q) a0 += part[(static_cast<size_t>(r) * P + q) * NV + v];
                        if (op.nmat == 2)
                            for (int q = 0; q < P; ++q) a1 += part[((static_cast<size_t>(nloc) + r) * P + q) * NV + v];
                        switch (op.epi) {
                            case DSE_STORE: out = rnd<T>(a0); break;
                            case DSE_RESID: out = rnd<T>(resid_own[v][r] + rnd<T>(a0)); resid_own[v][r] = out; break;
                            case DSE_SWIGLU: out = rnd<T>(rnd<T>(ds_silu(rnd<T>(a0))) * rnd<T>(a1)); break;
                            default: {      // DSE_LOGITS
                                out = rnd<T>(a0);
                                (op.logits + v * op.logits_stride)[n] = out;
                                if (out > best || (out == best && n < best_i)) { best = out; best_i = n; }
                                break;
                            }
                        }
                    }
                    const float hi = __shfl_xor_sync(0xffffffffu, out, 1);      // row r + 1 (nloc is even)
                    if (op.epi != DSE_LOGITS && r < nloc && !(r & 1))
                        ds_ll_store(op.yll + v * op.yll_stride + ((j0 + r) >> 1), Cvt<T>::pack2(out, hi), tag_out);
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
                        unsigned long long* cd = p.cand + (static_cast<size_t>(v) * G + cta) * 2;
                        ds_ll_store(cd, __float_as_uint(best), tag_out);
                        ds_ll_store(cd + 1, static_cast<uint32_t>(best_i), tag_out);
                    }
                    ds_consumer_sync();
                }
            }
            stamp(2);
        } else if (op.type == DS_ATTN) {
            ds_attention_op<T, NV>(op, p.st, p.att_part, xs, rope_cs, tag_of(oi - 1), tag_of(oi), ring, full_bar, empty_bar, p.n_slots, &rs, aprobe_w);
            stamp(4);
        } else {
            // ------------------------------------------------------------------ DS_FINAL: token selection (CTA 0)
            if (timed) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int cat = c == 3 ? 4 : c;
                    p.dbg[8 * (cta + 1) + cat] += t_acc[c];
                    if (cta == 0) p.dbg[cat] += t_acc[c];
                }
                DS_PROBE(p.dbg[8 * (cta + 1) + 6] += probe_acc[0]; p.dbg[8 * (cta + 1) + 7] += probe_acc[1];
                         if (cta == 0) { p.dbg[6] += probe_acc[0]; p.dbg[7] += probe_acc[1]; }
                         if (cta == 1) for (int i = 0; i < 8; ++i) p.dbg[8 * 149 + i] += aprobe_acc[i];)      // attention sub-phases of CTA 1
            }
            if (cta == 0 && warp == 0) {
                int all_done = 1;
                for (int v = 0; v < NV; ++v) {
                    float best = -INFINITY;
                    int best_i = 0x7fffffff;
                    const DsOp& head = p.ops[oi - 1];                                  // the lm_head GEMV: only CTAs with rows published a candidate
                    const int n_cand = (head.N + ds_rows_per_cta(head.N, G) - 1) / ds_rows_per_cta(head.N, G);
                    for (int c = lane; c < n_cand; c += 32) {
                        const unsigned long long* cd = p.cand + (static_cast<size_t>(v) * G + c) * 2;
                        const float cv = __uint_as_float(ds_ll_get(cd, tag_of(oi - 1)));
                        const int ci = static_cast<int>(ds_ll_get(cd + 1, tag_of(oi - 1)));
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
    return (static_cast<size_t>(group) * 136 + 3 * 128 + 2 * kDsGroupWarps * 8 + static_cast<size_t>(kDsGroupWarps) * group * 128 +
            static_cast<size_t>(kDsConsumerWarps) * (group > 4 ? group : 4) * 16) * sizeof(float);
}
inline size_t decode_stream_smem_bytes(int n_slots, int x_bytes, int part_cap) {
    return static_cast<size_t>(n_slots) * kDsSlotBytes + static_cast<size_t>(x_bytes) + static_cast<size_t>(part_cap) * sizeof(float);
}

}  // namespace smb
