// Host side of the library, part 1: weight slots, per-layer pointers, the handle (struct sm_handle: configuration, weights, workspaces, per-stream
// state, pipeline state, graphs), error / allocation helpers, launch helpers (PDL, kernel classes, profiling scopes).
// Fragment of the library's single translation unit: included by api.cu, in this order, inside nothing (it opens its own
// anonymous namespace where it needs one).
#pragma once

namespace {

std::string g_create_error;

struct Slot {
    void* dst = nullptr;       // destination (device)
    size_t row_bytes = 0;      // bytes per source row
    size_t rows = 0;           // number of rows
    size_t dst_pitch = 0;      // destination pitch in bytes (== row_bytes unless re-pitched)
    size_t numel = 0;
    bool loaded = false;
    // pre-tiled GEMM weights (ViT): destination matrix base, first row of this block in it, k-blocks per n-tile
    bool tiled = false;
    void* tile_base = nullptr;
    int tile_row0 = 0, tile_kb = 0, cols = 0;
};

struct VitLayer {
    void *ln1_w, *ln1_b, *wqkv, *bqkv, *wo, *bo, *ln2_w, *ln2_b, *w1, *b1, *w2, *b2;
};
struct MistralLayer {
    void *in_ln, *wqkv /* gate: only the v rows */, *wo, *post_ln, *wgu /* [2F, H]: gate rows then up rows */, *wd;
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

constexpr int kMaxLanes = 8;      // concurrent vision towers of the pipelined path (SMB_LANES overrides)
constexpr int kTicketRing = 16;   // frames in flight (sm_frame_submit tickets)
constexpr int kTowerBatch = 8;    // single-frame tickets whose towers run as one chunk (pipelined path)
constexpr int kMaxHandleStreams = 16;   // video streams one handle can hold (sm_config.n_streams)
constexpr int kDsMaxNew = 4096;   // tokens one sm_llm_decode call can produce per stream

struct sm_handle {
    sm_config cfg{};
    int device = 0;
    int num_sms = 148;
    int esz = 2;
    std::string err;
    std::vector<void*> allocs;
    // frame preprocessing (sm_preprocess_frames): resample tables per padded side, grow-only staging / intermediate buffers
    struct PreTable { int ksize; int* bounds; int* kk; int* kk_t; };
    std::map<int, PreTable> pre_tables;
    void* pre_src = nullptr; size_t pre_src_bytes = 0;
    void* pre_tmp = nullptr; size_t pre_tmp_bytes = 0;
    void* pre_lut = nullptr; float pre_lut_key[6] = {0, 0, 0, 0, 0, 0};
    void* cog_buf = nullptr; size_t cog_bytes = 0;      // sm_cognition_sample scratch (indices + similarities), grow-only
    std::unordered_map<std::string, Slot> slots;
    std::map<std::tuple<const void*, int, int, int>, CUtensorMap> tmaps;
    PFN_encodeTiled encode = nullptr;
    long long launches = 0;
    bool use_pdl = true;
    int gemm_class = 0;               // kernel class of gemm_tc_kernel launches (KC_GEMM; run_gate_gemm: KC_GATE_GEMM)
    bool vit_tiled = true;            // ViT GEMM weights are stored pre-tiled (gemm_tc.cuh GemmArgs::w_tiled)
    int attn_mode = -1;               // debug (sm_debug_attention_mode): -1 = SMB_ATTN_TC / auto, 0 = mma.sync kernel, 2 = tcgen05 wherever supported
    unsigned kfilter = 0xFFFFFFFFu;   // debug: kernel classes that are actually launched (bench.py per-class timing)
    long long* gemm_dbg = nullptr;   // device buffer for sm_test_gemm_trace
    bool profiling = false;
    struct ProfRec { int cls; cudaEvent_t a, b; };
    std::vector<ProfRec> prof;
    bool capturing = false;
    long long captured_launches = 0;

    // ---- ViT
    int S = 0, P = 0, kpad = 0;
    void *vit_cls = nullptr, *vit_wpatch = nullptr, *vit_pos = nullptr, *vit_pre_w = nullptr, *vit_pre_b = nullptr;
    std::vector<VitLayer> vit;
    void *ws_im = nullptr, *ws_pemb = nullptr, *ws_x = nullptr, *ws_h = nullptr, *ws_qkv = nullptr, *ws_att = nullptr,
         *ws_mlp = nullptr, *ws_pooled = nullptr, *ws_pixels = nullptr, *ws_feats = nullptr;
    float* ws_part = nullptr;   // split-K partial sums [4][rows][C] fp32
    // Two complete sets of tower activations ("lanes"): sm_frame_submit runs the towers of consecutive frames on two
    // streams at the same time, so the kernels of one frame fill the SMs the other frame's small GEMMs leave idle.
    // The ws_* fields above always point at the lane selected by select_lane() (lane 0 outside sm_frame_submit).
    struct VitWs { void *ws_im, *ws_pemb, *ws_x, *ws_h, *ws_qkv, *ws_att, *ws_mlp, *ws_pixels, *ws_feats; float* ws_part;
                   int cap_frames, part_frames; };   // frames the activation buffers / the split-K partial buffer hold
    VitWs lanes[kMaxLanes] = {};
    int cur_lane = 0;
    // ---- projector
    int d_inner = 0, dt_rank = 0;
    void *pj_pre_w = nullptr, *pj_pre_b = nullptr, *pj_norm_w = nullptr, *pj_norm_b = nullptr, *pj_in = nullptr,
         *pj_conv_w = nullptr, *pj_conv_b = nullptr, *pj_xproj = nullptr, *pj_dt_w = nullptr, *pj_dt_b = nullptr,
         *pj_alog = nullptr, *pj_D = nullptr, *pj_out = nullptr, *pj_nf_w = nullptr, *pj_nf_b = nullptr,
         *pj_post_w = nullptr, *pj_post_b = nullptr;
    void *pj_h0 = nullptr, *pj_xc = nullptr, *pj_z = nullptr, *pj_xdb = nullptr, *pj_y = nullptr, *pj_r2 = nullptr,
         *pj_conv_state = nullptr, *pj_toks = nullptr;
    float* pj_ssm_state = nullptr;
    // ---- gate
    std::vector<MistralLayer> gate;
    void *gt_norm = nullptr, *gt_head = nullptr, *gt_h = nullptr, *gt_v = nullptr, *gt_m = nullptr;
    void *gg_h = nullptr, *gg_hn = nullptr, *gg_v = nullptr, *gg_ve = nullptr, *gg_gu = nullptr, *gg_m = nullptr;   // batched gate as GEMMs
    float* gg_part = nullptr;   // split-K partials [8][rows][H]
    int gate_gemm_cap = 0;   // rows (frames) the gg_* buffers hold
    float* gt_logits = nullptr;
    // ---- llm
    std::vector<MistralLayer> llm;
    void *lm_embed = nullptr, *lm_norm = nullptr, *lm_head = nullptr;
    std::vector<void*> kc, vc;
    int pmax = 0;
    void *lw_x = nullptr, *lw_hn = nullptr, *lw_qkv = nullptr, *lw_att = nullptr, *lw_gu = nullptr, *lw_m = nullptr;
    float *lw_akv_o = nullptr, *lw_akv_ml = nullptr;   // split-KV partials of the tcgen05 prefill attention: [lw_akv_rows][Hq][128] and [..][2]
    int lw_akv_rows = 0;
    float *lw_logits = nullptr, *lw_part2 = nullptr;   // lw_logits [n_streams][V]: last-position logits of each stream's prefill; lw_part2: split-K partials of the few-row prefill GEMMs
    // ---- per-stream state (multi-stream batching): the handle holds n_streams video streams that share
    // its weights; `cur` is the stream the single-stream entry points act on (sm_stream_select)
    int n_streams = 1, cur = 0;
    std::vector<int> kv_lens;          // KV length per stream
    long long kv_stream_stride = 0;    // elements between the caches of consecutive streams inside kc[l] / vc[l]
    // ---- persistent decode kernel (decode_stream.cuh)
    DsOp* ds_ops = nullptr;            // device op list of one decode step
    int ds_n_ops = 0, ds_xcap = 0, ds_part_rows = 0 /* max over ops of nmat * rows-per-CTA * P */;
    unsigned* ds_sync = nullptr;       // [1] epoch of the decode kernel (never reset: its exchange tags derive from it), [2] all-done flag
    DsStreamState* ds_state = nullptr; // [kDsMaxStreams]
    int *ds_out = nullptr, *ds_stop = nullptr;   // ds_out [kDsMaxStreams][kDsMaxNew]
    float* ds_logits = nullptr;
    // exchange buffers of the decode kernel (tagged 8-byte words, decode_stream.cuh): [kDsMaxStreams][elements / 2]
    unsigned long long *ds_x_ll = nullptr, *ds_qkv_ll = nullptr, *ds_att_ll = nullptr, *ds_m_ll = nullptr;
    unsigned long long *ds_att_part = nullptr, *ds_cand = nullptr;
    long long* ds_dbg = nullptr;       // sm_debug_decode_phases: per-phase ns of CTA 0
    double ds_ms = 0.0;                // device time of the decode steps since the last sm_decode_stats reset (CUDA events)
    long long ds_steps = 0, ds_tokens = 0, ds_ctx_sum = 0;
    struct DsTiming { cudaEvent_t a, b; long long steps, tokens, ctx_sum; };
    std::vector<DsTiming> ds_pending;
    // ---- pipelined frame path (sm_frame_submit): tower on vit_stream, projector + gate on gate_stream
    bool pipe_init = false;
    int n_lanes = 8;                   // towers of consecutive tickets run on this many streams / activation sets (<= kMaxLanes)
    int plan_div = 2;                  // GEMM tile planner: accept the widest tile that yields >= num_sms / plan_div CTAs
    int split_sms = 0;                 // SMs a split-K GEMM may fill (0 = all)
    int max_split = 4;                 // split-K cap of the residual GEMMs (serial path: fill the SMs; pipelined: 1)
    int gemm_pre = 1;                  // GemmArgs::pre_weights
    cudaStream_t vit_streams[kMaxLanes] = {}, gate_stream = nullptr;
    cudaEvent_t ev_in = nullptr, ev_vit[kTicketRing] = {}, ev_gate[kTicketRing] = {};
    long long ticket = 0;
    void* pooled_ring = nullptr;       // [kTicketRing][max_frames][C]: pooled patch means, one slot per ticket in flight
    struct PendingTicket { void* feats_out; void* toks_out; float* logits_out; float* logits_host; int B; };
    PendingTicket pend[kTowerBatch] = {};   // tickets of the open batch (towers enqueued or, in tower-batch mode, only copied in)
    int n_pending = 0, gate_batch = 4;
    long long first_pending = 0;
    int tower_batch = 1;               // > 1: the towers of this many consecutive single-frame tickets run as ONE chunk
    void* px_ring = nullptr;           // [kTicketRing][3*H*W] staged pixels of the tickets in flight (tower-batch mode)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_px[kTicketRing] = {};
    // ---- graphs
    std::map<long long, cudaGraphExec_t> frame_graphs;   // key: gkey(h, int key) = kernel filter << 32 | key   // key: B | flags<<8
    std::map<long long, long long> frame_graph_launches;
    cudaStream_t cap_stream = nullptr;   // capture happens on a private stream (the legacy default stream cannot capture)
};

namespace {

int fail(sm_handle* h, const char* fmt, ...) {
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return 1;
}

#define CUDA_OK(h, expr)                                                                               \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess) return fail(h, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                                            __FILE__, __LINE__);                                       \
    } while (0)

void* dalloc(sm_handle* h, size_t bytes) {
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    if (cudaMalloc(&p, (bytes + 255) & ~size_t(255)) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, bytes);
    h->allocs.push_back(p);
    return p;
}

void add_slot(sm_handle* h, const std::string& name, void* dst, size_t rows, size_t cols, size_t dst_pitch_elems = 0) {
    Slot s;
    s.dst = dst;
    s.rows = rows;
    s.row_bytes = cols * h->esz;
    s.dst_pitch = (dst_pitch_elems ? dst_pitch_elems : cols) * h->esz;
    s.numel = rows * cols;
    h->slots[name] = s;
}

void add_tiled_slot(sm_handle* h, const std::string& name, void* matrix_base, int row0, size_t rows, size_t cols,
                    int kb_total) {
    Slot s;
    s.dst = matrix_base; s.rows = rows; s.row_bytes = cols * h->esz; s.dst_pitch = s.row_bytes; s.numel = rows * cols;
    s.tiled = true; s.tile_base = matrix_base; s.tile_row0 = row0; s.tile_kb = kb_total; s.cols = static_cast<int>(cols);
    h->slots[name] = s;
}
inline size_t tiled_elems(int n, int k) { return static_cast<size_t>((n + 127) / 128) * ((k + 63) / 64) * 128 * 64; }

inline void select_lane(sm_handle* h, int lane) {
    const sm_handle::VitWs& w = h->lanes[lane];
    h->ws_im = w.ws_im; h->ws_pemb = w.ws_pemb; h->ws_x = w.ws_x; h->ws_h = w.ws_h; h->ws_qkv = w.ws_qkv;
    h->ws_att = w.ws_att; h->ws_mlp = w.ws_mlp; h->ws_pixels = w.ws_pixels; h->ws_feats = w.ws_feats; h->ws_part = w.ws_part;
    h->cur_lane = lane;
}

// graph cache key: captured graphs embed the kernel filter and the selected stream's state pointers
inline long long gkey(const sm_handle* h, int key) {
    return (static_cast<long long>(h->kfilter | (static_cast<unsigned>(h->cur) << 16)) << 32) | static_cast<unsigned int>(key);
}

inline void count_launch(sm_handle* h) {
    if (h->capturing) h->captured_launches++; else h->launches++;
}

// Launch with the programmatic-stream-serialization attribute (PDL): the kernel may start while its
// predecessor drains and synchronises itself with griddepcontrol.wait (ptx.cuh pdl_wait).
template <typename... KArgs, typename... Args>
cudaError_t launch_ex(sm_handle* h, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                      int cluster_y, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    int n = 0;
    if (h->use_pdl) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster_y > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = 1; at[n].val.clusterDim.y = cluster_y; at[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = at; cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(sm_handle* h, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                       Args&&... args) {
    return launch_ex(h, kern, grid, block, smem, st, 1, std::forward<Args>(args)...);
}

// per-kernel-class CUDA-event timing (bench.py's roofline pass; off on the timed path)
enum KClass { KC_GEMM = 0, KC_GEMV, KC_ATTN, KC_LAYERNORM, KC_IM2COL, KC_VIT_FINALIZE, KC_MAMBA_SCAN, KC_ROPE_APPEND,
              KC_DECODE_ATTN, KC_ARGMAX, KC_GATHER, KC_RMSNORM_ROWS, KC_SWIGLU_ROWS, KC_GATE_GEMM, KC_COUNT };
const char* kKClassNames[KC_COUNT] = {"gemm_tc_kernel", "gemv_kernel", "attention_kernel", "layernorm_kernel",
                                      "im2col_kernel", "vit_finalize_kernel", "mamba_scan_step_kernel",
                                      "rope_append_kernel", "decode_attn_kernels", "argmax_kernel",
                                      "gather_rows_kernel", "rmsnorm_rows_kernel", "swiglu_rows_kernel",
                                      "gate_gemm_kernel"};   // gemm_tc_kernel launches of the batched gate (weight streaming)
inline bool kon(const sm_handle* h, int cls) { return (h->kfilter >> cls) & 1u; }
struct ProfScope {
    sm_handle* h; cudaStream_t st; cudaEvent_t b = nullptr;
    ProfScope(sm_handle* h_, int cls, cudaStream_t st_) : h(h_), st(st_) {
        if (!h->profiling || h->capturing) return;
        cudaEvent_t a;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, st);
        h->prof.push_back({cls, a, b});
    }
    ~ProfScope() { if (b) cudaEventRecord(b, st); }
};


}  // namespace
