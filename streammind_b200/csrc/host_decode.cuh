// Host side of the library, part 4: the persistent decode step (slot geometry, device op list, shared-memory plan, launch, timing).
// Fragment of the library's single translation unit: included by api.cu, in this order, inside nothing (it opens its own
// anonymous namespace where it needs one).
#pragma once

namespace {

// ------------------------------------------------------------------------------------------ persistent decode step
// Slot geometry of a GEMV op (decode_stream.cuh): R rows of K elements per 32 KB ring slot (a power of two <= 8),
// P = 8 / R parts per row, one (row, part) per consumer warp.
int ds_geometry(sm_handle* h, int K, int nmat, int* R, int* P) {
    const int fit = kDsSlotBytes / (K * 2);
    if (fit < nmat) return fail(h, "decode kernel: a row of K = %d elements does not fit a %d-byte ring slot", K, kDsSlotBytes);
    int r = kDsGroupWarps;
    while (r > fit) r >>= 1;
    const int p = kDsGroupWarps / r;
    if (K % p != 0 || (K / p) % 256 != 0) return fail(h, "decode kernel: K = %d cannot be split into %d parts of 256-weight block pairs", K, p);
    *R = r; *P = p;
    return 0;
}

// The op list of ONE decode step (hf MistralForCausalLM.forward for one new token per lane + greedy argmax), built once:
// per layer [qkv GEMV (RMSNorm prologue) | attention (RoPE, KV append, split-KV softmax, combine) | o_proj GEMV (+residual)
// | gate/up GEMV (RMSNorm prologue, SwiGLU epilogue) | down GEMV (+residual)], then lm_head (RMSNorm prologue, fp32
// logits, per-CTA argmax candidates) and the token selection.  embed_tokens is fused into the first layer.
int build_decode_ops(sm_handle* h) {
    const sm_config& c = h->cfg;
    const int H = c.llm_hidden, Hq = c.llm_heads, Hk = c.llm_kv_heads, D = c.llm_head_dim, F = c.llm_ffn, V = c.llm_vocab;
    const int QKV = (Hq + 2 * Hk) * D, G = h->num_sms;
    if (kDsMaxStreams * Hk > G) return fail(h, "decode kernel: %d lanes x %d kv heads exceed %d SMs", kDsMaxStreams, Hk, G);
    std::vector<DsOp> ops;
    int part_rows = 0, xcap = 0;
    // x / y: exchange buffers of the input and of what the op publishes, [lane][K / 2] and [lane][N / 2] words
    auto gemv = [&](const void* W0, const void* W1, int N, int K, int pro, int epi, const unsigned long long* x, const void* nw,
                    unsigned long long* y) -> int {
        if (N % 2 || K % 4) return fail(h, "decode kernel: GEMV shape %d x %d (rows must be even, columns a multiple of 4)", N, K);
        if (epi == DSE_RESID && (N != H || ds_rows_per_cta(H, G) > kDsResidRows))
            return fail(h, "decode kernel: residual rows per CTA exceed %d (hidden %d on %d SMs)", kDsResidRows, H, G);
        DsOp o{};
        o.type = DS_GEMV; o.W0 = W0; o.W1 = W1; o.nmat = W1 ? 2 : 1; o.N = N; o.K = K;
        if (ds_geometry(h, K, o.nmat, &o.R, &o.P)) return 1;
        o.pro = pro; o.epi = epi; o.xll = x; o.xll_stride = K / 2; o.nw = nw; o.eps = c.llm_eps; o.yll = y; o.yll_stride = N / 2;
        if (epi == DSE_LOGITS) { o.logits = h->ds_logits; o.logits_stride = N; }
        ops.push_back(o);
        part_rows = std::max(part_rows, o.nmat * ds_rows_per_cta(N, G) * o.P);
        xcap = std::max(xcap, (K + 7) & ~7);
        return 0;
    };
    if (D != 128) return fail(h, "decode kernel: head_dim %d (128 only)", D);
    for (int l = 0; l < c.llm_layers; ++l) {
        const MistralLayer& L = h->llm[l];
        if (gemv(L.wqkv, nullptr, QKV, H, l == 0 ? DSP_EMBED_RMSNORM : DSP_RMSNORM, DSE_STORE, h->ds_x_ll, L.in_ln, h->ds_qkv_ll)) return 1;
        DsOp a{};
        a.type = DS_ATTN; a.qkv_ll = h->ds_qkv_ll; a.qkv_ll_stride = QKV / 2; a.kc = h->kc[l]; a.vc = h->vc[l];
        a.kv_stream_stride = h->kv_stream_stride; a.att_ll = h->ds_att_ll; a.Hq = Hq; a.Hk = Hk; a.max_ctx = c.llm_max_ctx;
        a.rope_theta = c.llm_rope_theta;
        a.scale_log2e = static_cast<float>(1.4426950408889634 / std::sqrt(static_cast<double>(D)));
        ops.push_back(a);
        if (gemv(L.wo, nullptr, H, Hq * D, DSP_PLAIN, DSE_RESID, h->ds_att_ll, nullptr, h->ds_x_ll)) return 1;
        if (gemv(L.wgu, static_cast<const char*>(L.wgu) + static_cast<size_t>(F) * H * h->esz, F, H, DSP_RMSNORM, DSE_SWIGLU, h->ds_x_ll,
                 L.post_ln, h->ds_m_ll)) return 1;
        if (gemv(L.wd, nullptr, H, F, DSP_PLAIN, DSE_RESID, h->ds_m_ll, nullptr, h->ds_x_ll)) return 1;
    }
    if (gemv(h->lm_head, nullptr, V, H, DSP_RMSNORM, DSE_LOGITS, h->ds_x_ll, h->lm_norm, nullptr)) return 1;
    DsOp fin{};
    fin.type = DS_FINAL;
    ops.push_back(fin);
    h->ds_ops = static_cast<DsOp*>(dalloc(h, ops.size() * sizeof(DsOp)));
    if (!h->ds_ops) return fail(h, "decode kernel: out of device memory");
    CUDA_OK(h, cudaMemcpy(h->ds_ops, ops.data(), ops.size() * sizeof(DsOp), cudaMemcpyHostToDevice));
    h->ds_n_ops = static_cast<int>(ops.size());
    h->ds_xcap = xcap;
    h->ds_part_rows = part_rows;
    return 0;
}

// shared-memory plan of a launch with nv lanes: ring slots, staging region, partial sums
struct DsSmem { int n_slots, x_bytes, part_cap; size_t total; };
DsSmem ds_smem_plan(const sm_handle* h, int nv) {
    DsSmem m{};
    const int group = h->cfg.llm_heads / h->cfg.llm_kv_heads;
    m.x_bytes = static_cast<int>(std::max<size_t>(static_cast<size_t>(nv) * h->ds_xcap * 2, decode_stream_attn_scratch_bytes(group)));
    m.x_bytes = (m.x_bytes + 127) & ~127;
    m.part_cap = (h->ds_part_rows * nv + 31) & ~31;
    const long long budget = (227 - nv) * 1024 - (kDsGroups > 1 ? 2048 : 0) /* static shared memory of the kernel: <= 1 KB per stream */ - m.x_bytes - static_cast<long long>(m.part_cap) * 4;
    static const int env_slots = getenv("SMB_DS_SLOTS") ? atoi(getenv("SMB_DS_SLOTS")) : kDsMaxSlots;
    m.n_slots = static_cast<int>(std::max<long long>(0, std::min<long long>(std::min(env_slots, kDsMaxSlots), budget / kDsSlotBytes)));
    m.total = decode_stream_smem_bytes(m.n_slots, m.x_bytes, m.part_cap);
    return m;
}

template <typename T>
int launch_decode_step_t(sm_handle* h, int nv, cudaStream_t st) {
    const DsSmem m = ds_smem_plan(h, nv);
    if (m.n_slots < 2) return fail(h, "decode kernel: %d lanes leave no room for the weight ring", nv);
    DsParams p{};
    p.ops = h->ds_ops; p.n_ops = h->ds_n_ops; p.n_slots = m.n_slots; p.xcap = h->ds_xcap; p.x_bytes = m.x_bytes; p.part_cap = m.part_cap;
    p.sync = h->ds_sync; p.st = h->ds_state; p.out_ids = h->ds_out; p.out_stride = kDsMaxNew;
    p.stop = h->ds_stop; p.embed = h->lm_embed; p.H = h->cfg.llm_hidden; p.att_part = h->ds_att_part;
    p.cand = h->ds_cand; p.dbg = h->ds_dbg;
    {
        static const int flags = getenv("SMB_DS_DBG") ? atoi(getenv("SMB_DS_DBG")) : 0;
        p.dbg_flags = flags;
    }
    const dim3 grid(h->num_sms), block(ds_threads(nv));
    {
        static bool once = false;
        if (!once && getenv("SMB_DS_POLL_NS")) { const unsigned ns = atoi(getenv("SMB_DS_POLL_NS")); cudaMemcpyToSymbol(ds_poll_ns, &ns, sizeof ns); }
        once = true;
    }
    switch (nv) {
        case 1: decode_stream_kernel<T, 1><<<grid, block, m.total, st>>>(p); break;
        case 2: decode_stream_kernel<T, 2><<<grid, block, m.total, st>>>(p); break;
        case 3: decode_stream_kernel<T, 3><<<grid, block, m.total, st>>>(p); break;
        case 4: decode_stream_kernel<T, 4><<<grid, block, m.total, st>>>(p); break;
        default: return fail(h, "decode kernel: %d lanes not instantiated (1..%d)", nv, kDsMaxStreams);
    }
    count_launch(h);
    CUDA_OK(h, cudaGetLastError());
    return 0;
}
int launch_decode_step(sm_handle* h, int nv, cudaStream_t st) {
    DISPATCH_T(h, T, return launch_decode_step_t<T>(h, nv, st);)
}

// fold finished event pairs of earlier sm_llm_decode calls into the running totals (sm_decode_stats)
void ds_collect_timings(sm_handle* h, bool wait) {
    size_t k = 0;
    for (auto& t : h->ds_pending) {
        if (wait) cudaEventSynchronize(t.b);
        float ms = 0.f;
        if (cudaEventQuery(t.b) == cudaSuccess && cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
            h->ds_ms += ms; h->ds_steps += t.steps; h->ds_tokens += t.tokens; h->ds_ctx_sum += t.ctx_sum;
            cudaEventDestroy(t.a); cudaEventDestroy(t.b);
        } else {
            h->ds_pending[k++] = t;
        }
    }
    h->ds_pending.resize(k);
}


}  // namespace
