// Host side of the library, part 5: CUDA-graph capture and the pipelined frame path (tickets, tower batches, lanes, flush / join).
// Fragment of the library's single translation unit: included by api.cu, in this order, inside nothing (it opens its own
// anonymous namespace where it needs one).
#pragma once

namespace {

// Captures `body` (which must already have run once on a real stream, so tensor maps / plans exist) into a graph.
template <typename F>
int capture_graph(sm_handle* h, F&& body, cudaGraphExec_t* out, long long* n_launches, const char* what) {
    cudaGraph_t g;
    h->capturing = true;
    h->captured_launches = 0;
    CUDA_OK(h, cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
    const int rc = body(h->cap_stream);
    cudaError_t ce = cudaStreamEndCapture(h->cap_stream, &g);
    h->capturing = false;
    if (rc || ce != cudaSuccess) return fail(h, "%s: graph capture failed: %s", what, cudaGetErrorString(ce));
    CUDA_OK(h, cudaGraphInstantiate(out, g, 0));
    cudaGraphDestroy(g);
    *n_launches = h->captured_launches;
    return 0;
}

// Pipelined-path tile planner: with several towers in flight (or a chunk of frames) the SMs are kept busy by other
// work, so a GEMM is planned for bytes per flop (wide tiles, no split-K) instead of for its own latency.
// Measured at 4 lanes + gate batching (frames/s): split-K 4 -> 894, 2 -> 955, none -> 1025.
struct PipePlanScope {
    sm_handle* h; int div, ssm, msp, pre;
    explicit PipePlanScope(sm_handle* h_) : h(h_), div(h_->plan_div), ssm(h_->split_sms), msp(h_->max_split), pre(h_->gemm_pre) {
        static const int pdiv = getenv("SMB_PLAN_DIV") ? std::max(1, atoi(getenv("SMB_PLAN_DIV"))) : 8;
        static const int s_sm = getenv("SMB_SPLIT_SMS") ? atoi(getenv("SMB_SPLIT_SMS")) : 0;
        static const int s_k = getenv("SMB_PIPE_SPLITK") ? std::max(1, atoi(getenv("SMB_PIPE_SPLITK"))) : 1;
        static const int s_pre = getenv("SMB_PIPE_PRE") ? atoi(getenv("SMB_PIPE_PRE")) : 0;
        if (h->n_lanes > 1 || h->tower_batch > 1) { h->plan_div = pdiv; h->split_sms = s_sm; h->max_split = s_k; h->gemm_pre = s_pre; }
    }
    ~PipePlanScope() { select_lane(h, 0); h->plan_div = div; h->split_sms = ssm; h->max_split = msp; h->gemm_pre = pre; }
};

// run `body` on stream s: directly, or through a graph captured after the first (real) run
template <typename F>
int pipe_run_part(sm_handle* h, int key, cudaStream_t s, F&& body, const char* what) {
    if (!h->cfg.use_graphs) return body(s);
    auto it = h->frame_graphs.find(gkey(h, key));
    if (it == h->frame_graphs.end()) {
        if (body(s)) return 1;                       // real run: produces this call's outputs
        CUDA_OK(h, cudaStreamSynchronize(s));
        cudaGraphExec_t ge;
        long long n = 0;
        if (capture_graph(h, body, &ge, &n, what)) return 1;
        h->frame_graphs[gkey(h, key)] = ge;
        h->frame_graph_launches[gkey(h, key)] = n;
    } else {
        CUDA_OK(h, cudaGraphLaunch(it->second, s));
        h->launches += h->frame_graph_launches[gkey(h, key)];
    }
    return 0;
}

// Close the open batch of tickets.  Tower-batch mode: the towers of the pending single-frame tickets run as ONE chunk
// (pixels staged in px_ring) on the batch's lane; otherwise the towers were enqueued at submit time.  Then projector +
// gate for all pending frames as ONE batch on the gate stream: consecutive frames share every pass over the
// projector / gate weights (run_proj_gate).
int pipe_flush(sm_handle* h) {
    const int np = h->n_pending;
    if (np == 0) return 0;
    const sm_config& c = h->cfg;
    cudaStream_t gs = h->gate_stream;
    const long long first = h->first_pending;
    const int slot0 = static_cast<int>(first % kTicketRing);
    const int kf = static_cast<int>((h->kfilter & 0xFFFu) << 12);
    const size_t pooled_sz = static_cast<size_t>(c.vit_hidden) * h->esz, tok_sz = static_cast<size_t>(c.proj_d_model) * h->esz;
    // pending tickets are consecutive ring slots and (np > 1) single frames: their pooled vectors / pixels are contiguous
    char* pooled = static_cast<char*>(h->pooled_ring) + static_cast<size_t>(slot0) * c.max_frames * pooled_sz;
    int nframes = 0;
    for (int i = 0; i < np; ++i) nframes += h->pend[i].B;

    if (h->tower_batch > 1) {
        PipePlanScope plan(h);
        const int lane = 1 + static_cast<int>((first / h->tower_batch) % h->n_lanes);   // lanes 1..3 hold a chunk (lane 0: max_frames)
        if (np > h->lanes[lane].cap_frames) return fail(h, "pipe_flush: %d frames exceed lane capacity %d", np, h->lanes[lane].cap_frames);
        cudaStream_t vs = h->vit_streams[lane];
        select_lane(h, lane);
        bool want_feats = false;
        for (int i = 0; i < np; ++i) {
            CUDA_OK(h, cudaStreamWaitEvent(vs, h->ev_px[(first + i) % kTicketRing], 0));
            want_feats = want_feats || h->pend[i].feats_out != nullptr;
        }
        const size_t frame_px = static_cast<size_t>(3) * c.vit_image * c.vit_image * h->esz;
        const void* px = static_cast<const char*>(h->px_ring) + static_cast<size_t>(slot0) * frame_px;
        auto vit_body = [&](cudaStream_t s) -> int { return run_vit(h, px, np, want_feats ? h->ws_feats : nullptr, pooled, s); };
        if (pipe_run_part(h, np | (want_feats ? 1 << 8 : 0) | (1 << 9) | (1 << 10) | (lane << 28) | (slot0 << 24) | kf, vs, vit_body,
                          "sm_frame_submit(tower batch)")) return 1;
        const size_t feats_sz = static_cast<size_t>(h->P) * c.vit_hidden * h->esz;
        for (int i = 0; i < np; ++i)
            if (h->pend[i].feats_out)
                CUDA_OK(h, cudaMemcpyAsync(h->pend[i].feats_out, static_cast<const char*>(h->ws_feats) + i * feats_sz, feats_sz, cudaMemcpyDeviceToDevice, vs));
        CUDA_OK(h, cudaEventRecord(h->ev_vit[slot0], vs));
        CUDA_OK(h, cudaStreamWaitEvent(gs, h->ev_vit[slot0], 0));
    } else {
        for (int i = 0; i < np; ++i) CUDA_OK(h, cudaStreamWaitEvent(gs, h->ev_vit[(first + i) % kTicketRing], 0));
    }

    auto gate_body = [&](cudaStream_t s) -> int {
        return run_proj_gate(h, pooled, h->pj_toks, h->gt_logits, nframes, s);
    };
    if (pipe_run_part(h, nframes | (1 << 9) | (1 << 11) | (slot0 << 24) | kf, gs, gate_body, "sm_frame_submit(gate)")) return 1;
    int f0 = 0;
    for (int i = 0; i < np; ++i) {
        const sm_handle::PendingTicket& p = h->pend[i];
        const char* tk_src = static_cast<const char*>(h->pj_toks) + static_cast<size_t>(f0) * tok_sz;
        const float* lg_src = h->gt_logits + 2 * f0;
        if (p.toks_out) CUDA_OK(h, cudaMemcpyAsync(p.toks_out, tk_src, p.B * tok_sz, cudaMemcpyDeviceToDevice, gs));
        if (p.logits_out) CUDA_OK(h, cudaMemcpyAsync(p.logits_out, lg_src, static_cast<size_t>(p.B) * 2 * sizeof(float), cudaMemcpyDeviceToDevice, gs));
        if (p.logits_host) CUDA_OK(h, cudaMemcpyAsync(p.logits_host, lg_src, static_cast<size_t>(p.B) * 2 * sizeof(float), cudaMemcpyDeviceToHost, gs));
        f0 += p.B;
    }
    for (int i = 0; i < np; ++i) CUDA_OK(h, cudaEventRecord(h->ev_gate[(first + i) % kTicketRing], gs));
    h->first_pending = first + np;
    h->n_pending = 0;
    return 0;
}

// The serial entry points share the stream state (Mamba conv / SSM state, gate scratch) with the pipelined path:
// close the open gate batch and order the caller's stream after the last pipelined gate before touching it.
int pipe_join(sm_handle* h, cudaStream_t st) {
    if (!h->pipe_init || h->ticket == 0) return 0;
    if (pipe_flush(h)) return 1;
    CUDA_OK(h, cudaStreamWaitEvent(st, h->ev_gate[(h->ticket - 1) % kTicketRing], 0));
    return 0;
}


}  // namespace
