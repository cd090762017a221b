// C ABI of the B200-native StreamMind hot path (see include/streammind_b200.h for the contract and
// the reference interfaces each entry point replaces).  This file owns: weight slots (packed kernel
// layouts), per-stream state, workspaces, TMA tensor maps, launch sequences and CUDA-graph capture.
#include "../../include/streammind_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <unordered_map>
#include <utility>
#include <vector>

#include "attention.cuh"
#include "attention_tc.cuh"
#include "attention_kv_tc.cuh"
#include "preprocess.cuh"
#include "gemm_tc.cuh"
#include "gemv.cuh"
#include "decode_stream.cuh"
#include "misc_kernels.cuh"

using namespace smb;

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

// ------------------------------------------------------------------------------------------ tensor maps
const CUtensorMap* get_tmap(sm_handle* h, const void* ptr, int rows, int K, int box_rows) {
    auto key = std::make_tuple(ptr, rows, K, box_rows);
    auto it = h->tmaps.find(key);
    if (it != h->tmaps.end()) return &it->second;
    CUtensorMap m;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(K) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kGemmBK), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapDataType dt = h->cfg.dtype == SM_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUresult r = h->encode(&m, dt, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fail(h, "cuTensorMapEncodeTiled failed (%d) ptr=%p rows=%d K=%d box=%d", (int)r, ptr, rows, K, box_rows);
        return nullptr;
    }
    auto ins = h->tmaps.emplace(key, m);
    return &ins.first->second;
}

// ------------------------------------------------------------------------------------------ GEMM
struct GemmPlan { int swap, bn, bm2; };

// 256-row dual-accumulator tiles (GemmArgs::bm2): worth it when the tile count (x split-K) still feeds the machine --
// several towers in flight (plan_div >= 4) or a chunk of frames; they halve the weight bytes each SM ingests.
bool plan_bm2(int tokens, int feats, int split_k, int num_sms, int plan_div) {
    static const int mode = getenv("SMB_BM2") ? atoi(getenv("SMB_BM2")) : 0;   // measured: no gain (chunk 8: 962 vs 958 frames/s; B=1 x 4 lanes: 431 vs 633) -> opt-in
    if (mode == 0 || feats % 256 != 0 || tokens <= 256) return false;
    if (mode == 2) return true;
    const int tiles = ((tokens + 255) / 256) * (feats / 256) * std::max(1, split_k);
    return plan_div >= 4 ? tiles >= num_sms / (2 * plan_div) : tiles >= num_sms;
}

GemmPlan plan_gemm(int tokens, int feats, int K, int num_sms, int epi, int plan_div, int split_k = 1) {
    // Rule distilled from the graph-timed sweep in profiles/r01_gemm_plan_sweep.md: on B200 one tcgen05.mma
    // costs >= ~105 clocks whatever its N, so a CTA's mainloop lasts ~250 ns per K=64 slab for any tile width;
    // the best plan is the widest feature tile that still yields about half a wave of CTAs.  Transposed (swap)
    // tiles only pay off for a handful of tokens (weight rows fill the 128 MMA lanes, tokens ride on N >= 16).
    (void)K;
    const bool residual = epi == EPI_RESIDUAL || epi == EPI_STORE_F32;
    if (tokens <= 64 && !residual) return {1, std::max(16, (tokens + 15) / 16 * 16), 0};
    if (feats % 16 != 0) return {1, std::min(256, std::max(16, (tokens + 15) / 16 * 16)), 0};
    if (plan_bm2(tokens, feats, split_k, num_sms, plan_div)) return {0, 256, 1};
    const int mt = (tokens + 127) / 128;
    // measured (profiles/r01_gemm_plan_sweep.md, one streaming frame = 577 tokens): the wide fc1 GEMM is fastest with
    // weight rows on the MMA lanes and 160 tokens per tile (4 x 32 = 128 CTAs, no 2-byte-strided stores: TMA store)
    static const int fc1_swap = getenv("SMB_FC1_SWAP") ? atoi(getenv("SMB_FC1_SWAP")) : 1;
    if (fc1_swap && plan_div <= 2 && !residual && mt == 5 && feats >= 4096 && feats % 128 == 0) return {1, 160, 0};
    for (int bn : {256, 128, 64, 32})
        if (bn <= feats && mt * ((feats + bn - 1) / bn) >= num_sms / plan_div) return {0, bn, 0};
    return {0, std::min(32, feats), 0};
}

template <typename T>
int launch_gemm_t(sm_handle* h, const void* x, int tokens, const void* w, int feats, int K, const void* bias, void* out,
                  int ldo, int epi, cudaStream_t st, int force_swap = -1, int force_bn = 0, bool w_tiled = false,
                  int split_k = 1) {
    if (K % 8 != 0) return fail(h, "gemm: K=%d must be a multiple of 8", K);
    if (!kon(h, h->gemm_class)) return 0;
    GemmPlan p = plan_gemm(tokens, feats, K, h->num_sms, epi, h->plan_div, split_k);
    if (force_swap == 2) {                 // forced 256 x 256 dual-accumulator tile
        if (feats % 256 != 0) return fail(h, "gemm: the 256-row tile needs features %% 256 == 0");
        p.swap = 0; p.bn = 256; p.bm2 = 1;
    } else {
        if (force_swap >= 0) { p.swap = force_swap; p.bm2 = 0; }
        if (force_bn > 0) { p.bn = force_bn; p.bm2 = 0; }
    }
    if (!p.swap && (feats % 16 != 0)) return fail(h, "gemm: non-swapped layout needs features %% 16 == 0");
    const CUtensorMap *ta, *tb;
    GemmArgs a{};
    dim3 grid;
    const int w_kb = (K + kGemmBK - 1) / kGemmBK;
    const int w_rows_tiled = ((feats + 127) / 128) * w_kb * 128;
    a.w_tiled = w_tiled ? 1 : 0;
    a.w_kb = w_kb;
    if (!p.swap) {
        ta = get_tmap(h, x, tokens, K, kGemmBM);
        tb = w_tiled ? get_tmap(h, w, w_rows_tiled, kGemmBK, std::min(p.bn, kGemmBM)) : get_tmap(h, w, feats, K, p.bn);
        a.Ma = tokens; a.Nb = feats;
        grid = dim3((tokens + kGemmBM * (p.bm2 ? 2 : 1) - 1) / (kGemmBM * (p.bm2 ? 2 : 1)), (feats + p.bn - 1) / p.bn);
    } else {
        ta = w_tiled ? get_tmap(h, w, w_rows_tiled, kGemmBK, kGemmBM) : get_tmap(h, w, feats, K, kGemmBM);
        tb = get_tmap(h, x, tokens, K, p.bn);
        a.Ma = feats; a.Nb = tokens;
        grid = dim3((feats + kGemmBM - 1) / kGemmBM, (tokens + p.bn - 1) / p.bn);
    }
    // cluster along grid.y: the CTAs of a cluster share the A tile and multicast 128/CS-row slices of it
    int CS = 1;
    {
        static const int max_cs = getenv("SMB_GEMM_CLUSTER") ? atoi(getenv("SMB_GEMM_CLUSTER")) : 1;   // measured: no gain on B200 (the mainloop is MMA-issue bound)
        for (int c = 8; c > 1; c >>= 1)
            if (c <= max_cs && grid.y % c == 0) { CS = c; break; }
    }
    a.cluster_n = CS;
    if (CS > 1) {
        if (!p.swap) ta = get_tmap(h, x, tokens, K, kGemmBM / CS);
        else ta = w_tiled ? get_tmap(h, w, w_rows_tiled, kGemmBK, kGemmBM / CS) : get_tmap(h, w, feats, K, kGemmBM / CS);
    }
    if (split_k > 1) { grid.z = split_k; a.split_k = split_k; a.split_stride = static_cast<long long>(tokens) * feats; }
    if (!ta || !tb) return 1;
    a.K = K; a.bias = bias; a.out = out; a.ldo = ldo; a.swap = p.swap; a.bn = p.bn;
    a.bm2 = p.bm2;
    a.nstage = gemm_num_stages(p.bn, p.bm2); a.epi = epi;
    {
        static const int dm = getenv("SMB_GEMM_DBG_MODE") ? atoi(getenv("SMB_GEMM_DBG_MODE")) : 0;
        a.dbg_mode = dm;
    }
    a.pre_weights = h->gemm_pre;
    {
        static const int tp = getenv("SMB_GEMM_2PROD") ? atoi(getenv("SMB_GEMM_2PROD")) : 1;
        a.two_producers = tp;
    }
    a.dbg = h->gemm_dbg;
    if (h->gemm_dbg) h->gemm_dbg += 8;   // one 8-slot record per launch
    const CUtensorMap* tc = ta;  // placeholder when unused
    static const bool no_tma_store = getenv("SMB_NO_TMA_STORE") != nullptr;
    if (!p.swap) a.tma_store = (epi != EPI_STORE_F32 && p.bn % 64 == 0 && ldo == feats && !no_tma_store) ? 1 : 0;
    else a.tma_store = ((epi == EPI_STORE || epi == EPI_QUICK_GELU) && ldo == feats && feats % 8 == 0 && !no_tma_store) ? 1 : 0;
    if (a.tma_store) {
        tc = get_tmap(h, out, tokens, feats, p.swap ? p.bn : kGemmBM);
        if (!tc) return 1;
    }
    {
        static const int dbg_stages = getenv("SMB_GEMM_STAGES") ? atoi(getenv("SMB_GEMM_STAGES")) : 0;   // tuning knob
        if (dbg_stages > 0 && dbg_stages < a.nstage) a.nstage = dbg_stages;
    }
    const int smem = gemm_smem_bytes(p.bn, p.bm2);
    {
        ProfScope ps(h, h->gemm_class, st);
        CUDA_OK(h, launch_ex(h, gemm_tc_kernel<T>, grid, dim3(kGemmThreads2), smem, st, CS, *ta, *tb, *tc, a));
    }
    count_launch(h);
    CUDA_OK(h, cudaGetLastError());
    return 0;
}

int launch_gemm(sm_handle* h, const void* x, int tokens, const void* w, int feats, int K, const void* bias, void* out,
                int ldo, int epi, cudaStream_t st, int force_swap = -1, int force_bn = 0, bool w_tiled = false,
                int split_k = 1) {
    if (h->cfg.dtype == SM_DTYPE_BF16)
        return launch_gemm_t<__nv_bfloat16>(h, x, tokens, w, feats, K, bias, out, ldo, epi, st, force_swap, force_bn, w_tiled, split_k);
    return launch_gemm_t<__half>(h, x, tokens, w, feats, K, bias, out, ldo, epi, st, force_swap, force_bn, w_tiled, split_k);
}

// Split-K factor for a residual GEMM whose 128x128 tiles alone cannot fill the SMs (streaming B = 1):
// every MMA instruction costs >= ~105 clocks whatever its N (profiles/r01_gemm_phases.md), so a CTA's time is
// ~ k-blocks x 250 ns and the only way to shorten it is to give each CTA fewer k-blocks.
int splitk_factor(const sm_handle* h, int tokens, int feats, int K, bool bm2 = false) {
    const int max_split = h->max_split;
    const int split_sms = h->split_sms;
    if (h->S > 0 && tokens > h->lanes[h->cur_lane].part_frames * h->S) return 1;   // partial-sum buffer of this lane is smaller
    const int tiles = bm2 ? ((tokens + 255) / 256) * ((feats + 255) / 256) : ((tokens + 127) / 128) * ((feats + 127) / 128);
    const int kb = (K + kGemmBK - 1) / kGemmBK;
    int s = std::min({max_split, (split_sms > 0 ? split_sms : h->num_sms) / std::max(1, tiles), kb / 4});
    while (s > 1 && (s - 1) * ((kb + s - 1) / s) >= kb) --s;   // every split gets at least one k-block
    return s < 2 ? 1 : s;
}

// ------------------------------------------------------------------------------------------ GEMV
template <typename T>
int launch_gemv_t(sm_handle* h, GemvArgs a, int nmat, cudaStream_t st) {
    if (a.K % 8 != 0) return fail(h, "gemv: K=%d must be a multiple of 8", a.K);
    if (!kon(h, KC_GEMV)) return 0;
    const int nv = std::max(1, a.nv_host);
    const int nvt = nv;
    a.seg_len = nv == 1 ? (nmat == 1 ? SMB_GEMV_UNR1 * 256 : 1024) : 2048;   // one batch of loads covers a segment
    int grid = std::min(a.N, (nv == 1 ? 2 : 1) * h->num_sms);
    const int rows_per_cta = (a.N + grid - 1) / grid;
    grid = (a.N + rows_per_cta - 1) / rows_per_cta;
    const int nseg = (a.K + a.seg_len - 1) / a.seg_len;
    const int xpitch = (a.K + 7) & ~7;
    const size_t smem = ((static_cast<size_t>(nvt) * xpitch * 2 + 15) & ~size_t(15)) +
                        static_cast<size_t>(nmat) * nvt * rows_per_cta * nseg * sizeof(float);
    if (smem > (nv == 1 ? 100u : 200u) * 1024) return fail(h, "gemv: K=%d x %d vectors too large for the staging buffer", a.K, nv);
    {
        ProfScope ps(h, KC_GEMV, st);
        const dim3 g(grid), b(kGemvThreads);
#define SMB_GEMV_CASE(NM, NVV) CUDA_OK(h, launch_pdl(h, gemv_kernel<T, NM, NVV>, g, b, smem, st, a))
        switch (nmat * 10 + nvt) {
            case 11: SMB_GEMV_CASE(1, 1); break;
            case 12: SMB_GEMV_CASE(1, 2); break;
            case 13: SMB_GEMV_CASE(1, 3); break;
            case 14: SMB_GEMV_CASE(1, 4); break;
            case 21: SMB_GEMV_CASE(2, 1); break;
            case 22: SMB_GEMV_CASE(2, 2); break;
            case 23: SMB_GEMV_CASE(2, 3); break;
            case 24: SMB_GEMV_CASE(2, 4); break;
            default: return fail(h, "gemv: %d matrices x %d vectors not instantiated", nmat, nv);
        }
#undef SMB_GEMV_CASE
    }
    count_launch(h);
    CUDA_OK(h, cudaGetLastError());
    return 0;
}
int launch_gemv(sm_handle* h, const GemvArgs& a, int nmat, cudaStream_t st) {
    if (h->cfg.dtype == SM_DTYPE_BF16) return launch_gemv_t<__nv_bfloat16>(h, a, nmat, st);
    return launch_gemv_t<__half>(h, a, nmat, st);
}
GemvArgs gv(const void* W, int N, int K, int pro, const void* x0, int epi, void* y) {
    GemvArgs a{};
    a.W0 = W; a.N = N; a.K = K; a.pro = pro; a.x0 = x0; a.epi = epi; a.y = y;
    return a;
}

#define DISPATCH_T(h, T, ...)                          \
    if ((h)->cfg.dtype == SM_DTYPE_BF16) {             \
        using T = __nv_bfloat16;                       \
        __VA_ARGS__                                    \
    } else {                                           \
        using T = __half;                              \
        __VA_ARGS__                                    \
    }

// ------------------------------------------------------------------------------------------ attention
template <typename T, int D>
int launch_attn_t(sm_handle* h, const AttnArgs& a, int q_tiles, int heads, int batch, cudaStream_t st) {
    constexpr int smem = attn_smem_bytes<D>();
    if (!kon(h, KC_ATTN)) return 0;
    {
        ProfScope ps(h, KC_ATTN, st);
        CUDA_OK(h, launch_pdl(h, attention_kernel<T, D>, dim3(q_tiles, heads, batch), dim3(kAttnThreads), smem, st, a));
    }
    count_launch(h);
    CUDA_OK(h, cudaGetLastError());
    return 0;
}
// tcgen05 attention (attention_tc.cuh): d = 64, non-causal, q / k / v packed in one row-major matrix, batch items
// contiguous (the vision tower's qkv activation)
template <typename T>
int launch_attn_tc_t(sm_handle* h, const AttnArgs& a, int heads, int batch, cudaStream_t st) {
    if (!kon(h, KC_ATTN)) return 0;
    const int pitch = static_cast<int>(a.q_ss);                 // elements per packed row (3C)
    const CUtensorMap* tm = get_tmap(h, a.q, batch * a.q_len, pitch, kAtcTile);
    if (!tm) return 1;
    AttnTcArgs t{};
    t.o = a.o; t.o_ss = a.o_ss; t.S = a.q_len; t.col_q = 0;
    t.col_k = static_cast<int>((static_cast<const char*>(a.k) - static_cast<const char*>(a.q)) / 2);
    t.col_v = static_cast<int>((static_cast<const char*>(a.v) - static_cast<const char*>(a.q)) / 2);
    t.scale_log2e = a.scale_log2e;
    t.dbg = h->gemm_dbg;
    {
        ProfScope ps(h, KC_ATTN, st);
        const dim3 grid((a.q_len + kAtcTile - 1) / kAtcTile, heads, batch);
        CUDA_OK(h, launch_pdl(h, attention_tc_kernel<T>, grid, dim3(kAtcThreads), static_cast<size_t>(attn_tc_smem_bytes()), st, *tm, t));
    }
    count_launch(h);
    CUDA_OK(h, cudaGetLastError());
    return 0;
}

// tcgen05 prefill attention (attention_kv_tc.cuh): causal GQA, d = 128, queries = rows of a packed (already rotated) activation,
// keys / values = one stream's cache [Hk][max_ctx][128].  n_splits: 0 = planned (about one wave of CTAs), > 0 forced (tests).
int plan_attn_kv_splits(const sm_handle* h, int P, int pos0, int group, int Hk) {
    const int TB = 128 / group, q_tiles = (P + TB - 1) / TB, nb_max = (pos0 + P + 127) / 128;
    return std::max(1, std::min(nb_max, h->num_sms / (q_tiles * Hk)));
}
bool attn_kv_tc_ok(const sm_handle* h, int D, int Hq, int Hk) {
    static const int env_on = getenv("SMB_PREFILL_ATTN_TC") ? atoi(getenv("SMB_PREFILL_ATTN_TC")) : 1;
    const int on = h->attn_mode >= 0 ? (h->attn_mode != 0) : env_on;
    if (!on || D != 128 || Hq % Hk != 0) return false;
    const int group = Hq / Hk;
    return 128 % group == 0 && (128 / group) % 8 == 0;
}
template <typename T>
int launch_attn_kv_tc_t(sm_handle* h, const void* q, int q_rows, int q_pitch, int col_q, const void* kc, const void* vc, int max_ctx, void* o,
                        int o_ss, int P, int pos0, int Hq, int Hk, float scale_log2e, int n_splits, cudaStream_t st) {
    if (!kon(h, KC_ATTN)) return 0;
    const int group = Hq / Hk, TB = 128 / group;
    if (n_splits <= 0) n_splits = plan_attn_kv_splits(h, P, pos0, group, Hk);
    if (n_splits > 1 && (h->lw_akv_o == nullptr || static_cast<long long>(n_splits) * P > h->lw_akv_rows))
        return fail(h, "prefill attention: %d splits x %d positions exceed the partial buffer (%d rows)", n_splits, P, h->lw_akv_rows);
    const CUtensorMap* tq = get_tmap(h, q, q_rows, q_pitch, TB);
    const CUtensorMap* tk = get_tmap(h, kc, Hk * max_ctx, 128, 128);
    const CUtensorMap* tv = get_tmap(h, vc, Hk * max_ctx, 128, 128);
    if (!tq || !tk || !tv) return 1;
    AttnKvArgs a{};
    a.o = o; a.o_ss = o_ss; a.P = P; a.pos0 = pos0; a.group = group; a.Hq = Hq; a.max_ctx = max_ctx; a.col_q = col_q;
    a.n_splits = n_splits; a.scale_log2e = scale_log2e; a.ws_o = h->lw_akv_o; a.ws_ml = h->lw_akv_ml;
    {
        ProfScope ps(h, KC_ATTN, st);
        const dim3 grid((P + TB - 1) / TB, Hk, n_splits);
        CUDA_OK(h, launch_pdl(h, attention_kv_tc_kernel<T>, grid, dim3(kAkvThreads), static_cast<size_t>(attn_kv_smem_bytes()), st, *tq, *tk, *tv, a));
        count_launch(h);
        if (n_splits > 1) {
            CUDA_OK(h, launch_pdl(h, attention_kv_merge_kernel<T>, dim3(P, Hq), dim3(128), 0, st, a));
            count_launch(h);
        }
    }
    CUDA_OK(h, cudaGetLastError());
    return 0;
}

int launch_attn(sm_handle* h, const AttnArgs& a, int D, int heads, int batch, cudaStream_t st) {
    // one frame alone is only 80 CTAs of 128 query rows: the 64-row mma.sync kernel (160 CTAs) fills the machine better
    static const int env_tc = getenv("SMB_ATTN_TC") ? atoi(getenv("SMB_ATTN_TC")) : 1;
    const int use_tc = h->attn_mode >= 0 ? h->attn_mode : env_tc;
    static const int tc_min_ctas = getenv("SMB_ATTN_TC_MINCTAS") ? atoi(getenv("SMB_ATTN_TC_MINCTAS")) : 148;
    if (use_tc && ((a.q_len + kAtcTile - 1) / kAtcTile) * heads * batch >= (use_tc == 2 ? 0 : tc_min_ctas) && D == 64 && !a.causal && a.group == 1 && a.q_len == a.kv_len && a.q_ss == a.k_ss && a.q_ss == a.v_ss &&
        a.k_hs == 64 && a.v_hs == 64 && a.q_bs == static_cast<long long>(a.q_len) * a.q_ss && a.k_bs == a.q_bs && a.v_bs == a.q_bs &&
        a.o_bs == static_cast<long long>(a.q_len) * a.o_ss && (a.q_ss * 2) % 16 == 0) {
        DISPATCH_T(h, T, return launch_attn_tc_t<T>(h, a, heads, batch, st);)
    }
    const int q_tiles = (a.q_len + kAttnBQ - 1) / kAttnBQ;
    if (D == 64) { DISPATCH_T(h, T, return launch_attn_t<T, 64>(h, a, q_tiles, heads, batch, st);) }
    if (D == 128) { DISPATCH_T(h, T, return launch_attn_t<T, 128>(h, a, q_tiles, heads, batch, st);) }
    return fail(h, "attention: head_dim %d not supported (64 or 128)", D);
}

template <typename T>
int init_kernel_attrs_t(sm_handle* h) {
    CUDA_OK(h, cudaFuncSetAttribute(gemm_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(h, cudaFuncSetAttribute(gemv_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    {
        auto big = [&](const void* f) { return cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); };
        CUDA_OK(h, big((const void*)gemv_kernel<T, 1, 2>)); CUDA_OK(h, big((const void*)gemv_kernel<T, 1, 3>));
        CUDA_OK(h, big((const void*)gemv_kernel<T, 1, 4>)); CUDA_OK(h, big((const void*)gemv_kernel<T, 2, 2>));
        CUDA_OK(h, big((const void*)gemv_kernel<T, 2, 3>)); CUDA_OK(h, big((const void*)gemv_kernel<T, 2, 4>));
    }
    CUDA_OK(h, cudaFuncSetAttribute(gemv_kernel<T, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CUDA_OK(h, cudaFuncSetAttribute(attention_kernel<T, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bytes<64>()));
    CUDA_OK(h, cudaFuncSetAttribute(attention_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_tc_smem_bytes()));
    CUDA_OK(h, cudaFuncSetAttribute(attention_kernel<T, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bytes<128>()));
    CUDA_OK(h, cudaFuncSetAttribute(attention_kv_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_kv_smem_bytes()));
    CUDA_OK(h, cudaFuncSetAttribute(decode_stream_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (227 - 1) * 1024 - (kDsGroups > 1 ? 2048 : 0)));   // static: <= 1 KB per stream
    CUDA_OK(h, cudaFuncSetAttribute(decode_stream_kernel<T, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (227 - 2) * 1024 - (kDsGroups > 1 ? 2048 : 0)));   // static: <= 1 KB per stream
    CUDA_OK(h, cudaFuncSetAttribute(decode_stream_kernel<T, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (227 - 3) * 1024 - (kDsGroups > 1 ? 2048 : 0)));   // static: <= 1 KB per stream
    CUDA_OK(h, cudaFuncSetAttribute(decode_stream_kernel<T, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (227 - 4) * 1024 - (kDsGroups > 1 ? 2048 : 0)));   // static: <= 1 KB per stream
    return 0;
}

// ------------------------------------------------------------------------------------------ sub-model runners
int run_vit(sm_handle* h, const void* pixels, int B, void* feats_out, void* pooled_out, cudaStream_t st) {
    const sm_config& c = h->cfg;
    const int C = c.vit_hidden, S = h->S, P = h->P, rows = B * S, F = c.vit_ffn;
    DISPATCH_T(h, T, {
        const long long n = static_cast<long long>(B) * P * (3 * c.vit_patch + 1);
        ProfScope ps_kc_im2col(h, KC_IM2COL, st);
        if (kon(h, KC_IM2COL)) {
        CUDA_OK(h, launch_pdl(h, im2col_kernel<T>, dim3(static_cast<int>(std::min<long long>((n + 255) / 256, 4096))), dim3(256), 0, st,
            reinterpret_cast<const T*>(pixels), reinterpret_cast<T*>(h->ws_im), B, c.vit_image, c.vit_patch, h->kpad));
        }
        count_launch(h);
    })
    if (launch_gemm(h, h->ws_im, B * P, h->vit_wpatch, C, h->kpad, nullptr, h->ws_pemb, C, EPI_STORE, st, -1, 0, h->vit_tiled)) return 1;
    const int warps_per_block = 8;
    const int ln_blocks = (rows + warps_per_block - 1) / warps_per_block;
    DISPATCH_T(h, T, {
        ProfScope ps_kc_layernorm(h, KC_LAYERNORM, st);
        if (kon(h, KC_LAYERNORM)) {
        CUDA_OK(h, launch_pdl(h, vit_embed_ln_kernel<T, 32>, dim3(ln_blocks), dim3(warps_per_block * 32), 0, st,
            (const T*)h->ws_pemb, (const T*)h->vit_cls, (const T*)h->vit_pos, (const T*)h->vit_pre_w,
            (const T*)h->vit_pre_b, (const T*)h->vit[0].ln1_w, (const T*)h->vit[0].ln1_b, (T*)h->ws_x, (T*)h->ws_h, rows,
            S, C, c.vit_eps));
        }
        count_launch(h);
    })
    const int D = C / c.vit_heads;
    // x += W a + bias, then h = LN(x) (ln_w == nullptr: no LN).  Small token counts: split-K GEMM into fp32
    // partials whose fixed-order sum, the residual add and the LayerNorm run in one row kernel.
    auto residual_gemm_ln = [&](const void* a_in, const void* W, int K, const void* bias, const void* ln_w,
                                const void* ln_b) -> int {
        const bool bm2 = h->vit_tiled && plan_bm2(rows, C, 3, h->num_sms, h->plan_div);
        const int S = ((C & 255) == 0 && C <= 1024) ? splitk_factor(h, rows, C, K, bm2) : 1;
        if (S > 1) {
            // forced plan: 128-wide tiles, or (force_swap = 2) the 256 x 256 dual-accumulator tile
            if (launch_gemm(h, a_in, rows, W, C, K, nullptr, h->ws_part, C, EPI_STORE_F32, st, bm2 ? 2 : 0, bm2 ? 256 : 128, h->vit_tiled, S)) return 1;
            DISPATCH_T(h, T, {
                ProfScope ps_kc_layernorm(h, KC_LAYERNORM, st);
                if (kon(h, KC_LAYERNORM)) {
                CUDA_OK(h, launch_pdl(h, splitk_residual_ln_kernel<T>, dim3(rows), dim3(C / 8), 0, st,
                    (const float*)h->ws_part, S, static_cast<long long>(rows) * C, (const T*)bias, (T*)h->ws_x, (const T*)ln_w,
                    (const T*)ln_b, (T*)h->ws_h, rows, C, c.vit_eps));
                }
                count_launch(h);
            })
        } else {
            if (launch_gemm(h, a_in, rows, W, C, K, bias, h->ws_x, C, EPI_RESIDUAL, st, -1, 0, h->vit_tiled)) return 1;
            if (ln_w != nullptr) {
                DISPATCH_T(h, T, {
                    ProfScope ps_kc_layernorm(h, KC_LAYERNORM, st);
                    if (kon(h, KC_LAYERNORM)) {
                    CUDA_OK(h, launch_pdl(h, layernorm_kernel<T>, dim3(ln_blocks), dim3(warps_per_block * 32), 0, st,
                        (const T*)h->ws_x, (const T*)ln_w, (const T*)ln_b, (T*)h->ws_h, rows, C, c.vit_eps));
                    }
                    count_launch(h);
                })
            }
        }
        return 0;
    };
    for (int l = 0; l < c.vit_layers; ++l) {
        const VitLayer& L = h->vit[l];
        if (launch_gemm(h, h->ws_h, rows, L.wqkv, 3 * C, C, L.bqkv, h->ws_qkv, 3 * C, EPI_STORE, st, -1, 0, h->vit_tiled)) return 1;
        AttnArgs a{};
        a.q = h->ws_qkv;
        a.k = reinterpret_cast<const char*>(h->ws_qkv) + static_cast<size_t>(C) * 2;
        a.v = reinterpret_cast<const char*>(h->ws_qkv) + static_cast<size_t>(2 * C) * 2;
        a.o = h->ws_att;
        a.q_bs = a.k_bs = a.v_bs = static_cast<long long>(S) * 3 * C;
        a.q_ss = a.k_ss = a.v_ss = 3 * C;
        a.k_hs = a.v_hs = D;
        a.o_bs = static_cast<long long>(S) * C;
        a.o_ss = C;
        a.q_len = S; a.kv_len = S; a.q_pos0 = 0; a.causal = 0; a.group = 1;
        a.scale_log2e = static_cast<float>(1.4426950408889634 / std::sqrt(static_cast<double>(D)));
        if (launch_attn(h, a, D, c.vit_heads, B, st)) return 1;
        if (residual_gemm_ln(h->ws_att, L.wo, C, L.bo, L.ln2_w, L.ln2_b)) return 1;
        if (launch_gemm(h, h->ws_h, rows, L.w1, F, C, L.b1, h->ws_mlp, F, EPI_QUICK_GELU, st, -1, 0, h->vit_tiled)) return 1;
        const bool last = l + 1 == c.vit_layers;
        if (residual_gemm_ln(h->ws_mlp, L.w2, F, L.b2, last ? nullptr : h->vit[l + 1].ln1_w,
                             last ? nullptr : h->vit[l + 1].ln1_b)) return 1;
    }
    DISPATCH_T(h, T, {
        ProfScope ps_kc_vit_finalize(h, KC_VIT_FINALIZE, st);
        if (kon(h, KC_VIT_FINALIZE)) {
        CUDA_OK(h, launch_pdl(h, vit_finalize_kernel<T>, dim3((C / 8 + 3) / 4, B), dim3(128), 0, st, (const T*)h->ws_x, (T*)feats_out,
                                                                         (T*)pooled_out, S, C));
        }
        count_launch(h);
    })
    CUDA_OK(h, cudaGetLastError());
    return 0;
}

// nv consecutive frames (<= kGemvBatch) in one pass over the weights: every GEMV takes nv input vectors, the two
// sequential pieces (conv window in the in_proj epilogue, SSM state in the scan kernel) walk the frames in order.
constexpr int kGemvBatch = 4;

// multi = true: the nv frames belong to the nv consecutive stream slots starting at h->cur (one frame each) instead of being
// nv consecutive frames of stream h->cur
int run_projector(sm_handle* h, const void* pooled, void* tok_out, int nv, cudaStream_t st, bool multi = false) {
    const sm_config& c = h->cfg;
    const int Dm = c.proj_d_model, Di = h->d_inner, R = h->dt_rank, N = c.proj_d_state, C = c.vit_hidden;
    const int nxp = (R + 2 * N + 7) & ~7;
    auto batched = [&](GemvArgs& a, long long xs, long long ys, long long rs = 0, long long zs = 0) {
        a.nv_host = nv; a.x_stride = xs; a.y_stride = ys; a.resid_stride = rs; a.z_stride = zs;
    };
    GemvArgs a = gv(h->pj_pre_w, Dm, C, PRO_PLAIN, pooled, GEPI_LEAKY, h->pj_h0);
    a.bias = h->pj_pre_b;
    batched(a, C, Dm);
    if (launch_gemv(h, a, 1, st)) return 1;
    a = gv(h->pj_in, 2 * Di, Dm, PRO_LAYERNORM, h->pj_h0, GEPI_MAMBA_CONV, h->pj_xc);
    a.nw = h->pj_norm_w; a.nb = h->pj_norm_b; a.eps = c.proj_eps;
    a.conv_state = static_cast<char*>(h->pj_conv_state) + static_cast<size_t>(h->cur) * Di * c.proj_d_conv * h->esz; a.conv_w = h->pj_conv_w; a.conv_b = h->pj_conv_b; a.z_out = h->pj_z;
    a.d_inner = Di; a.d_conv = c.proj_d_conv;
    a.conv_state_stride = multi ? static_cast<long long>(Di) * c.proj_d_conv : 0;
    batched(a, Dm, Di, 0, Di);
    if (launch_gemv(h, a, 1, st)) return 1;
    a = gv(h->pj_xproj, R + 2 * N, Di, PRO_PLAIN, h->pj_xc, GEPI_STORE, h->pj_xdb);
    batched(a, Di, nxp);
    if (launch_gemv(h, a, 1, st)) return 1;
    ScanArgs s{};
    s.W_dt = h->pj_dt_w; s.b_dt = h->pj_dt_b; s.A_log = h->pj_alog; s.D = h->pj_D; s.xdb = h->pj_xdb; s.x = h->pj_xc;
    s.z = h->pj_z; s.state = h->pj_ssm_state + static_cast<size_t>(h->cur) * Di * N; s.y = h->pj_y; s.d_inner = Di; s.dt_rank = R; s.d_state = N;
    s.nv = nv; s.xdb_stride = nxp; s.x_stride = Di; s.z_stride = Di; s.y_stride = Di;
    s.state_stride = multi ? static_cast<long long>(Di) * N : 0;
    const int scan_smem = (nv * nxp * 2 + 15) & ~15;
    DISPATCH_T(h, T, {
        ProfScope ps_kc_mamba_scan(h, KC_MAMBA_SCAN, st);
        if (kon(h, KC_MAMBA_SCAN)) {
        CUDA_OK(h, launch_pdl(h, mamba_scan_step_kernel<T>, dim3(std::min((Di + 7) / 8, 8 * h->num_sms)), dim3(256), scan_smem, st, s));
        }
        count_launch(h);
    })
    a = gv(h->pj_out, Dm, Di, PRO_PLAIN, h->pj_y, GEPI_ADD_TO, h->pj_r2);
    a.resid = h->pj_h0;
    batched(a, Di, Dm, Dm);
    if (launch_gemv(h, a, 1, st)) return 1;
    a = gv(h->pj_post_w, Dm, Dm, PRO_LN_LEAKY, h->pj_r2, GEPI_STORE, tok_out);
    a.nw = h->pj_nf_w; a.nb = h->pj_nf_b; a.eps = c.proj_eps; a.bias = h->pj_post_b;
    batched(a, Dm, Dm);
    if (launch_gemv(h, a, 1, st)) return 1;
    return 0;
}

int run_gate(sm_handle* h, const void* tok, float* logits_out, int nv, cudaStream_t st) {
    const sm_config& c = h->cfg;
    const int H = c.proj_d_model, Hq = c.gate_heads, Hk = c.gate_kv_heads, D = c.gate_head_dim, F = c.gate_ffn;
    auto batched = [&](GemvArgs& a, long long xs, long long ys, long long rs = 0) {
        a.nv_host = nv; a.x_stride = xs; a.y_stride = ys; a.resid_stride = rs;
    };
    CUDA_OK(h, cudaMemcpyAsync(h->gt_h, tok, static_cast<size_t>(nv) * H * h->esz, cudaMemcpyDeviceToDevice, st));
    for (int l = 0; l < c.gate_layers; ++l) {
        const MistralLayer& L = h->gate[l];
        GemvArgs a = gv(L.wqkv, Hk * D, H, PRO_RMSNORM, h->gt_h, GEPI_STORE, h->gt_v);
        a.nw = L.in_ln; a.eps = c.gate_eps;
        batched(a, H, Hk * D);
        if (launch_gemv(h, a, 1, st)) return 1;
        a = gv(L.wo, H, Hq * D, PRO_GQA_EXPAND, h->gt_v, GEPI_RESID, nullptr);
        a.resid = h->gt_h; a.gqa_rep = Hq / Hk; a.head_dim = D;
        batched(a, Hk * D, 0, H);
        if (launch_gemv(h, a, 1, st)) return 1;
        a = gv(L.wgu, F, H, PRO_RMSNORM, h->gt_h, GEPI_SWIGLU, h->gt_m);
        a.W1 = reinterpret_cast<const char*>(L.wgu) + static_cast<size_t>(F) * H * h->esz;
        a.nw = L.post_ln; a.eps = c.gate_eps;
        batched(a, H, F);
        if (launch_gemv(h, a, 2, st)) return 1;
        a = gv(L.wd, H, F, PRO_PLAIN, h->gt_m, GEPI_RESID, nullptr);
        a.resid = h->gt_h;
        batched(a, F, 0, H);
        if (launch_gemv(h, a, 1, st)) return 1;
    }
    GemvArgs a = gv(h->gt_head, 2, H, PRO_RMSNORM, h->gt_h, GEPI_F32, logits_out);
    a.nw = h->gt_norm; a.eps = c.gate_eps;
    batched(a, H, 2);
    return launch_gemv(h, a, 1, st);
}

// out[n, N] (= or +=) x[n, K] . W[N, K]^T for a few rows (n <= 64) with the weight rows on the MMA lanes (swap plan).
// N / 128 CTAs alone cannot pull HBM bandwidth for narrow outputs (32-48 tiles for the 4096 / 6144-wide projections), so
// K is split until about one CTA per SM streams weights; the fp32 partials are summed in fixed order by
// splitk_rows_kernel (T(resid + T(sum)) for the in-place residual stream).  part: [8][n][N] floats.
int gemm_few_rows(sm_handle* h, const void* x, int n, const void* W, int N, int K, void* out, bool residual, float* part,
                  cudaStream_t st) {
    const int tiles = (N + 127) / 128, kb = (K + 63) / 64;
    const int bn = std::max(16, (n + 15) / 16 * 16);
    int split = std::min({8, std::max(1, h->num_sms / tiles), std::max(1, kb / 8)});
    while (split > 1 && (split - 1) * ((kb + split - 1) / split) >= kb) --split;
    if (split < 2 || part == nullptr)
        return launch_gemm(h, x, n, W, N, K, nullptr, out, N, residual ? EPI_RESIDUAL : EPI_STORE, st, 1, bn);
    if (launch_gemm(h, x, n, W, N, K, nullptr, part, N, EPI_STORE_F32, st, 1, bn, false, split)) return 1;
    DISPATCH_T(h, T, {
        const long long tot = static_cast<long long>(n) * N;
        if (kon(h, h->gemm_class)) {
        CUDA_OK(h, launch_pdl(h, splitk_rows_kernel<T>, dim3(static_cast<int>(std::min<long long>((tot + 255) / 256, 1024))), dim3(256), 0, st,
                              (const float*)part, split, tot, residual ? (const T*)out : (const T*)nullptr, (T*)out, tot));
        }
        count_launch(h);
    })
    return 0;
}

// The gate for n >= gate_gemm_min frames as tensor-core GEMMs (the gate at L = 1 is a token-wise MLP stack, so n
// frames are n independent rows): weights on the 128 MMA lanes (swap plan), the n rows on the MMA N dimension, every
// weight byte streamed once for all n frames by TMA.  Same rounding points as the GEMV chain (rmsnorm rows, T outputs,
// T(silu) * up, in-place residual); the 2-row lm_head stays a GEMV.
int run_gate_gemm(sm_handle* h, const void* toks, float* logits_out, int n, cudaStream_t st) {
    const sm_config& c = h->cfg;
    const int H = c.proj_d_model, Hq = c.gate_heads, Hk = c.gate_kv_heads, D = c.gate_head_dim, F = c.gate_ffn;
    if (n > h->gate_gemm_cap) return fail(h, "run_gate_gemm: %d rows exceed capacity %d", n, h->gate_gemm_cap);
    struct ClassScope { sm_handle* h; ~ClassScope() { h->gemm_class = KC_GEMM; } } class_scope{h};
    h->gemm_class = KC_GATE_GEMM;
    if (!kon(h, KC_GATE_GEMM)) return 0;
    CUDA_OK(h, cudaMemcpyAsync(h->gg_h, toks, static_cast<size_t>(n) * H * h->esz, cudaMemcpyDeviceToDevice, st));
    const int nb = (n + 7) / 8;
    auto rms = [&](const void* nw) -> int {
        DISPATCH_T(h, T, {
            CUDA_OK(h, launch_pdl(h, rmsnorm_rows_kernel<T>, dim3(nb), dim3(256), 0, st, (const T*)h->gg_h, (const T*)nw, (T*)h->gg_hn, n, H, c.gate_eps));
            count_launch(h);
        })
        return 0;
    };
    auto mm = [&](const void* x, const void* W, int N, int K, void* out, bool residual) -> int {
        return gemm_few_rows(h, x, n, W, N, K, out, residual, h->gg_part, st);
    };
    for (int l = 0; l < c.gate_layers; ++l) {
        const MistralLayer& L = h->gate[l];
        if (rms(L.in_ln)) return 1;
        if (mm(h->gg_hn, L.wqkv, Hk * D, H, h->gg_v, false)) return 1;
        DISPATCH_T(h, T, {
            const long long tot = static_cast<long long>(n) * Hq * D / 8;
            CUDA_OK(h, launch_pdl(h, gqa_expand_rows_kernel<T>, dim3(static_cast<int>(std::min<long long>((tot + 255) / 256, 1024))), dim3(256), 0, st,
                                  (const T*)h->gg_v, (T*)h->gg_ve, n, Hq, Hk, D));
            count_launch(h);
        })
        if (mm(h->gg_ve, L.wo, H, Hq * D, h->gg_h, true)) return 1;
        if (rms(L.post_ln)) return 1;
        if (mm(h->gg_hn, L.wgu, 2 * F, H, h->gg_gu, false)) return 1;
        DISPATCH_T(h, T, {
            const long long tot = static_cast<long long>(n) * F;
            CUDA_OK(h, launch_pdl(h, swiglu_rows_kernel<T>, dim3(static_cast<int>(std::min<long long>((tot + 255) / 256, 4096))), dim3(256), 0, st,
                                  (const T*)h->gg_gu, (T*)h->gg_m, n, F));
            count_launch(h);
        })
        if (mm(h->gg_m, L.wd, H, F, h->gg_h, true)) return 1;
    }
    // final RMSNorm + lm_head [2, H]: GEMV over the rows, kGemvBatch at a time
    for (int i = 0; i < n; i += kGemvBatch) {
        const int nv = std::min(kGemvBatch, n - i);
        GemvArgs a = gv(h->gt_head, 2, H, PRO_RMSNORM, static_cast<const char*>(h->gg_h) + static_cast<size_t>(i) * H * h->esz, GEPI_F32, logits_out + 2 * i);
        a.nw = h->gt_norm; a.eps = c.gate_eps;
        a.nv_host = nv; a.x_stride = H; a.y_stride = 2;
        if (launch_gemv(h, a, 1, st)) return 1;
    }
    return 0;
}

// projector + gate for n frames: batches of <= kGemvBatch frames share every weight pass
int run_proj_gate(sm_handle* h, const void* pooled, void* toks, float* logits, int n, cudaStream_t st) {
    const sm_config& c = h->cfg;
    static const int max_batch = getenv("SMB_GEMV_BATCH") ? std::max(1, std::min(kGemvBatch, atoi(getenv("SMB_GEMV_BATCH")))) : kGemvBatch;
    static const int gemm_min = getenv("SMB_GATE_GEMM") ? atoi(getenv("SMB_GATE_GEMM")) : 5;   // frames from which the gate runs as GEMMs (0 = never)
    const bool gate_gemm = gemm_min > 0 && n >= gemm_min && n <= h->gate_gemm_cap && c.proj_d_model % 64 == 0 && c.gate_ffn % 64 == 0;
    for (int i = 0; i < n; i += max_batch) {
        const int nv = std::min(max_batch, n - i);
        char* tok = static_cast<char*>(toks) + static_cast<size_t>(i) * c.proj_d_model * h->esz;
        if (run_projector(h, static_cast<const char*>(pooled) + static_cast<size_t>(i) * c.vit_hidden * h->esz, tok, nv, st)) return 1;
        if (!gate_gemm && run_gate(h, tok, logits + 2 * i, nv, st)) return 1;
    }
    if (gate_gemm) return run_gate_gemm(h, toks, logits, n, st);
    return 0;
}

int run_prefill_chunk(sm_handle* h, const void* embeds, int P, int pos0, cudaStream_t st) {
    const sm_config& c = h->cfg;
    const int H = c.llm_hidden, Hq = c.llm_heads, Hk = c.llm_kv_heads, D = c.llm_head_dim, F = c.llm_ffn;
    const int QKV = (Hq + 2 * Hk) * D;
    CUDA_OK(h, cudaMemcpyAsync(h->lw_x, embeds, static_cast<size_t>(P) * H * h->esz, cudaMemcpyDeviceToDevice, st));
    const int nb = (P + 7) / 8;
    // short dialogue suffixes (a fire prefills 11-74 new tokens): the narrow projections are split along K
    static const bool few_on = getenv("SMB_PREFILL_SPLITK") ? atoi(getenv("SMB_PREFILL_SPLITK")) != 0 : true;
    const bool few = few_on && P <= 64 && h->lw_part2 != nullptr;
    for (int l = 0; l < c.llm_layers; ++l) {
        const MistralLayer& L = h->llm[l];
        DISPATCH_T(h, T, {
            ProfScope ps_kc_rmsnorm_rows(h, KC_RMSNORM_ROWS, st);
            if (kon(h, KC_RMSNORM_ROWS)) {
            rmsnorm_rows_kernel<T><<<nb, 256, 0, st>>>((const T*)h->lw_x, (const T*)L.in_ln, (T*)h->lw_hn, P, H, c.llm_eps);
            }
            count_launch(h);
        })
        if (few) { if (gemm_few_rows(h, h->lw_hn, P, L.wqkv, QKV, H, h->lw_qkv, false, h->lw_part2, st)) return 1; }
        else if (launch_gemm(h, h->lw_hn, P, L.wqkv, QKV, H, nullptr, h->lw_qkv, QKV, EPI_STORE, st)) return 1;
        DISPATCH_T(h, T, {
            const long long tot = static_cast<long long>(P) * ((Hq + Hk) * (D / 2) + Hk * D);
            rope_append_kernel<T><<<static_cast<int>(std::min<long long>((tot + 255) / 256, 2048)), 256, 0, st>>>(
                (T*)h->lw_qkv, (T*)h->kc[l] + h->cur * h->kv_stream_stride, (T*)h->vc[l] + h->cur * h->kv_stream_stride, P, Hq, Hk, D, c.llm_max_ctx, nullptr, pos0, c.llm_rope_theta);
            count_launch(h);
        })
        AttnArgs a{};
        a.q = h->lw_qkv; a.o = h->lw_att;
        a.k = static_cast<char*>(h->kc[l]) + h->cur * h->kv_stream_stride * h->esz;
        a.v = static_cast<char*>(h->vc[l]) + h->cur * h->kv_stream_stride * h->esz;
        a.q_bs = 0; a.q_ss = QKV;
        a.k_bs = a.v_bs = 0; a.k_hs = a.v_hs = static_cast<long long>(c.llm_max_ctx) * D; a.k_ss = a.v_ss = D;
        a.o_bs = 0; a.o_ss = Hq * D;
        a.q_len = P; a.kv_len = pos0 + P; a.q_pos0 = pos0; a.causal = 1; a.group = Hq / Hk;
        a.scale_log2e = static_cast<float>(1.4426950408889634 / std::sqrt(static_cast<double>(D)));
        if (attn_kv_tc_ok(h, D, Hq, Hk)) {
            DISPATCH_T(h, T, { if (launch_attn_kv_tc_t<T>(h, h->lw_qkv, h->pmax, QKV, 0, a.k, a.v, c.llm_max_ctx, h->lw_att, Hq * D, P, pos0, Hq, Hk,
                                                          a.scale_log2e, 0, st)) return 1; })
        } else if (launch_attn(h, a, D, Hq, 1, st)) return 1;
        if (few) { if (gemm_few_rows(h, h->lw_att, P, L.wo, H, Hq * D, h->lw_x, true, h->lw_part2, st)) return 1; }
        else if (launch_gemm(h, h->lw_att, P, L.wo, H, Hq * D, nullptr, h->lw_x, H, EPI_RESIDUAL, st)) return 1;
        DISPATCH_T(h, T, {
            ProfScope ps_kc_rmsnorm_rows(h, KC_RMSNORM_ROWS, st);
            if (kon(h, KC_RMSNORM_ROWS)) {
            rmsnorm_rows_kernel<T><<<nb, 256, 0, st>>>((const T*)h->lw_x, (const T*)L.post_ln, (T*)h->lw_hn, P, H, c.llm_eps);
            }
            count_launch(h);
        })
        if (launch_gemm(h, h->lw_hn, P, L.wgu, 2 * F, H, nullptr, h->lw_gu, 2 * F, EPI_STORE, st)) return 1;
        DISPATCH_T(h, T, {
            const long long tot = static_cast<long long>(P) * F;
            ProfScope ps_kc_swiglu_rows(h, KC_SWIGLU_ROWS, st);
            if (kon(h, KC_SWIGLU_ROWS)) {
            swiglu_rows_kernel<T><<<static_cast<int>(std::min<long long>((tot + 255) / 256, 4096)), 256, 0, st>>>(
                (const T*)h->lw_gu, (T*)h->lw_m, P, F);
            }
            count_launch(h);
        })
        if (few) { if (gemm_few_rows(h, h->lw_m, P, L.wd, H, F, h->lw_x, true, h->lw_part2, st)) return 1; }
        else if (launch_gemm(h, h->lw_m, P, L.wd, H, F, nullptr, h->lw_x, H, EPI_RESIDUAL, st)) return 1;
    }
    CUDA_OK(h, cudaGetLastError());
    return 0;
}


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

// =========================================================================================== C ABI
extern "C" {

const char* sm_last_error(const sm_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int sm_create(sm_handle** out, int device, const sm_config* cfg) {
    if (!out || !cfg) return fail(nullptr, "sm_create: null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, "sm_create: no CUDA device visible (the CUDA path is the only path; there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(nullptr, "sm_create: device %d out of range (%d devices)", device, ndev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, "cudaGetDeviceProperties failed");
    if (prop.major != 10) return fail(nullptr, "sm_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    if (cfg->dtype != SM_DTYPE_F16 && cfg->dtype != SM_DTYPE_BF16) return fail(nullptr, "sm_create: dtype must be fp16 or bf16");
    cudaSetDevice(device);
    sm_handle* h = new sm_handle();
    h->cfg = *cfg;
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    const sm_config& c = h->cfg;
    if (c.max_frames < 1) h->cfg.max_frames = 1;
    h->n_streams = std::max(1, c.n_streams);
    if (h->n_streams > kMaxHandleStreams) { delete h; return fail(nullptr, "sm_create: n_streams %d > %d", c.n_streams, kMaxHandleStreams); }
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
        delete h;
        return fail(nullptr, "sm_create: cuTensorMapEncodeTiled not available from the driver");
    }
    h->encode = reinterpret_cast<PFN_encodeTiled>(fn);
    h->use_pdl = getenv("SMB_NO_PDL") == nullptr;
    h->max_split = getenv("SMB_SPLITK") ? std::max(1, atoi(getenv("SMB_SPLITK"))) : 4;
    h->gemm_pre = getenv("SMB_GEMM_PRE") ? atoi(getenv("SMB_GEMM_PRE")) : 1;
    if (cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return fail(nullptr, "sm_create: cudaStreamCreate failed");
    }
    {
        int rc = 0;
        DISPATCH_T(h, T, rc = init_kernel_attrs_t<T>(h);)
        if (rc) { g_create_error = h->err; delete h; return 1; }
    }
    const size_t e = h->esz;
    const int Bm = h->cfg.max_frames;
    bool oom = false;
    auto A = [&](size_t n) { void* p = dalloc(h, n); if (!p) oom = true; return p; };

    if (c.vit_image > 0 && c.vit_patch > 0) {
        const int gw0 = c.vit_image / c.vit_patch;
        h->P = gw0 * gw0;
        h->S = h->P + 1;
    }
    // ---------------- ViT
    if (c.vit_layers > 0) {
        if (c.vit_hidden % c.vit_heads || c.vit_hidden / c.vit_heads != 64 || c.vit_hidden % 64 || c.vit_ffn % 64 ||
            c.vit_hidden > 1024) {
            delete h;
            return fail(nullptr, "sm_create: ViT needs head_dim 64, hidden %% 64 == 0, hidden <= 1024, ffn %% 64 == 0");
        }
        const int C = c.vit_hidden, F = c.vit_ffn, gw = c.vit_image / c.vit_patch;
        h->P = gw * gw;
        h->S = h->P + 1;
        const int kreal = 3 * c.vit_patch * c.vit_patch;
        h->kpad = (kreal + 63) / 64 * 64;
        const std::string p = "model.vision_tower.vision_tower.vision_model.";
        h->vit_cls = A(C * e); add_slot(h, p + "embeddings.class_embedding", h->vit_cls, 1, C);
        h->vit_tiled = getenv("SMB_NO_TILED_WEIGHTS") == nullptr && C % 128 == 0 && F % 128 == 0;
        const int kbC = (C + 63) / 64, kbF = (F + 63) / 64;
        if (h->vit_tiled) {
            h->vit_wpatch = A(tiled_elems(C, h->kpad) * e);
            add_tiled_slot(h, p + "embeddings.patch_embedding.weight", h->vit_wpatch, 0, C, kreal, h->kpad / 64);
        } else {
            h->vit_wpatch = A(static_cast<size_t>(C) * h->kpad * e);
            add_slot(h, p + "embeddings.patch_embedding.weight", h->vit_wpatch, C, kreal, h->kpad);
        }
        h->vit_pos = A(static_cast<size_t>(h->S) * C * e);
        add_slot(h, p + "embeddings.position_embedding.weight", h->vit_pos, h->S, C);
        h->vit_pre_w = A(C * e); add_slot(h, p + "pre_layrnorm.weight", h->vit_pre_w, 1, C);
        h->vit_pre_b = A(C * e); add_slot(h, p + "pre_layrnorm.bias", h->vit_pre_b, 1, C);
        h->vit.resize(c.vit_layers);
        for (int l = 0; l < c.vit_layers; ++l) {
            VitLayer& L = h->vit[l];
            const std::string lp = p + "encoder.layers." + std::to_string(l) + ".";
            L.ln1_w = A(C * e); add_slot(h, lp + "layer_norm1.weight", L.ln1_w, 1, C);
            L.ln1_b = A(C * e); add_slot(h, lp + "layer_norm1.bias", L.ln1_b, 1, C);
            L.ln2_w = A(C * e); add_slot(h, lp + "layer_norm2.weight", L.ln2_w, 1, C);
            L.ln2_b = A(C * e); add_slot(h, lp + "layer_norm2.bias", L.ln2_b, 1, C);
            L.wqkv = A(tiled_elems(3 * C, C) * e);
            L.bqkv = A(static_cast<size_t>(3) * C * e);
            const char* nm[3] = {"q_proj", "k_proj", "v_proj"};
            for (int j = 0; j < 3; ++j) {
                if (h->vit_tiled) add_tiled_slot(h, lp + "self_attn." + nm[j] + ".weight", L.wqkv, j * C, C, C, kbC);
                else add_slot(h, lp + "self_attn." + nm[j] + ".weight", (char*)L.wqkv + static_cast<size_t>(j) * C * C * e, C, C);
                add_slot(h, lp + "self_attn." + nm[j] + ".bias", (char*)L.bqkv + static_cast<size_t>(j) * C * e, 1, C);
            }
            L.wo = A(tiled_elems(C, C) * e);
            if (h->vit_tiled) add_tiled_slot(h, lp + "self_attn.out_proj.weight", L.wo, 0, C, C, kbC);
            else add_slot(h, lp + "self_attn.out_proj.weight", L.wo, C, C);
            L.bo = A(C * e); add_slot(h, lp + "self_attn.out_proj.bias", L.bo, 1, C);
            L.w1 = A(tiled_elems(F, C) * e);
            if (h->vit_tiled) add_tiled_slot(h, lp + "mlp.fc1.weight", L.w1, 0, F, C, kbC);
            else add_slot(h, lp + "mlp.fc1.weight", L.w1, F, C);
            L.b1 = A(F * e); add_slot(h, lp + "mlp.fc1.bias", L.b1, 1, F);
            L.w2 = A(tiled_elems(C, F) * e);
            if (h->vit_tiled) add_tiled_slot(h, lp + "mlp.fc2.weight", L.w2, 0, C, F, kbF);
            else add_slot(h, lp + "mlp.fc2.weight", L.w2, C, F);
            L.b2 = A(C * e); add_slot(h, lp + "mlp.fc2.bias", L.b2, 1, C);
        }
        const size_t rows = static_cast<size_t>(Bm) * h->S;
        h->ws_pixels = A(static_cast<size_t>(Bm) * 3 * c.vit_image * c.vit_image * e);
        h->ws_im = A(static_cast<size_t>(Bm) * h->P * h->kpad * e);
        h->ws_pemb = A(static_cast<size_t>(Bm) * h->P * C * e);
        h->ws_x = A(rows * C * e);
        h->ws_h = A(rows * C * e);
        h->ws_qkv = A(rows * 3 * C * e);
        h->ws_att = A(rows * C * e);
        h->ws_mlp = A(rows * F * e);
        h->ws_pooled = A(static_cast<size_t>(Bm) * C * e);
        h->pooled_ring = A(static_cast<size_t>(kTicketRing) * Bm * C * e);
        h->ws_feats = A(static_cast<size_t>(Bm) * h->P * C * e);
        h->ws_part = static_cast<float*>(A(static_cast<size_t>(4) * rows * C * sizeof(float)));
        h->lanes[0] = {h->ws_im, h->ws_pemb, h->ws_x, h->ws_h, h->ws_qkv, h->ws_att, h->ws_mlp, h->ws_pixels, h->ws_feats, h->ws_part, Bm, Bm};
        for (int ln = 1; ln < kMaxLanes; ++ln) {
            sm_handle::VitWs& w = h->lanes[ln];
            // lanes 1..3 can also hold a chunk of kTowerBatch single-frame tickets (tower-batch mode of the pipelined path)
            const int capf = (ln <= 3 && Bm < kTowerBatch) ? kTowerBatch : Bm;
            const size_t lrows = static_cast<size_t>(capf) * h->S;
            w.cap_frames = capf; w.part_frames = Bm;
            w.ws_pixels = A(static_cast<size_t>(Bm) * 3 * c.vit_image * c.vit_image * e);
            w.ws_im = A(static_cast<size_t>(capf) * h->P * h->kpad * e);
            w.ws_pemb = A(static_cast<size_t>(capf) * h->P * C * e);
            w.ws_x = A(lrows * C * e);
            w.ws_h = A(lrows * C * e);
            w.ws_qkv = A(lrows * 3 * C * e);
            w.ws_att = A(lrows * C * e);
            w.ws_mlp = A(lrows * F * e);
            w.ws_feats = A(static_cast<size_t>(capf) * h->P * C * e);
            w.ws_part = static_cast<float*>(A(static_cast<size_t>(4) * rows * C * sizeof(float)));
        }
        h->px_ring = A(static_cast<size_t>(kTicketRing) * 3 * c.vit_image * c.vit_image * e);
    }
    // ---------------- projector
    if (c.proj_d_model > 0) {
        const int Dm = c.proj_d_model, C = c.vit_hidden, N = c.proj_d_state, W = c.proj_d_conv;
        h->d_inner = c.proj_expand * Dm;
        h->dt_rank = (Dm + 15) / 16;
        const int Di = h->d_inner, R = h->dt_rank;
        if (Dm % 8 || C % 8 || Di % 8 || R % 8) {
            delete h;
            return fail(nullptr, "sm_create: projector dims must be multiples of 8 (d_model %d, dt_rank %d)", Dm, R);
        }
        const std::string p = "model.mm_projector.", mp = p + "mamba_model.ssms.0.";
        h->pj_pre_w = A(static_cast<size_t>(Dm) * C * e); add_slot(h, p + "pre_net.fc3.weight", h->pj_pre_w, Dm, C);
        h->pj_pre_b = A(Dm * e); add_slot(h, p + "pre_net.fc3.bias", h->pj_pre_b, 1, Dm);
        h->pj_norm_w = A(Dm * e); add_slot(h, mp + "norm.weight", h->pj_norm_w, 1, Dm);
        h->pj_norm_b = A(Dm * e); add_slot(h, mp + "norm.bias", h->pj_norm_b, 1, Dm);
        h->pj_in = A(static_cast<size_t>(2) * Di * Dm * e); add_slot(h, mp + "mixer.in_proj.weight", h->pj_in, 2 * Di, Dm);
        h->pj_conv_w = A(static_cast<size_t>(Di) * W * e); add_slot(h, mp + "mixer.conv1d.weight", h->pj_conv_w, Di, W);
        h->pj_conv_b = A(Di * e); add_slot(h, mp + "mixer.conv1d.bias", h->pj_conv_b, 1, Di);
        h->pj_xproj = A(static_cast<size_t>(R + 2 * N) * Di * e); add_slot(h, mp + "mixer.x_proj.weight", h->pj_xproj, R + 2 * N, Di);
        h->pj_dt_w = A(static_cast<size_t>(Di) * R * e); add_slot(h, mp + "mixer.dt_proj.weight", h->pj_dt_w, Di, R);
        h->pj_dt_b = A(Di * e); add_slot(h, mp + "mixer.dt_proj.bias", h->pj_dt_b, 1, Di);
        h->pj_alog = A(static_cast<size_t>(Di) * N * e); add_slot(h, mp + "mixer.A_log", h->pj_alog, Di, N);
        h->pj_D = A(Di * e); add_slot(h, mp + "mixer.D", h->pj_D, 1, Di);
        h->pj_out = A(static_cast<size_t>(Dm) * Di * e); add_slot(h, mp + "mixer.out_proj.weight", h->pj_out, Dm, Di);
        h->pj_nf_w = A(Dm * e); add_slot(h, p + "mamba_model.norm_fn.weight", h->pj_nf_w, 1, Dm);
        h->pj_nf_b = A(Dm * e); add_slot(h, p + "mamba_model.norm_fn.bias", h->pj_nf_b, 1, Dm);
        h->pj_post_w = A(static_cast<size_t>(Dm) * Dm * e); add_slot(h, p + "post_net.fc3.weight", h->pj_post_w, Dm, Dm);
        h->pj_post_b = A(Dm * e); add_slot(h, p + "post_net.fc3.bias", h->pj_post_b, 1, Dm);
        constexpr int NB = 4;   // kGemvBatch frames share one pass over the weights
        if (N > 32) { delete h; return fail(nullptr, "sm_create: projector d_state %d > 32 not supported", N); }
        h->pj_h0 = A(NB * Dm * e); h->pj_xc = A(NB * Di * e); h->pj_z = A(NB * Di * e); h->pj_xdb = A(NB * ((R + 2 * N + 7) & ~7) * e + 64);
        h->pj_y = A(NB * Di * e); h->pj_r2 = A(NB * Dm * e);
        h->pj_conv_state = A(static_cast<size_t>(h->n_streams) * Di * W * e);                           // per stream
        h->pj_ssm_state = static_cast<float*>(A(static_cast<size_t>(h->n_streams) * Di * N * sizeof(float)));
        h->pj_toks = A(static_cast<size_t>(std::max(Bm, kTowerBatch)) * Dm * e);
    }
    // ---------------- gate
    if (c.gate_layers > 0) {
        const int H = c.proj_d_model, Hq = c.gate_heads, Hk = c.gate_kv_heads, D = c.gate_head_dim, F = c.gate_ffn;
        if (H % 8 || F % 8 || (Hq * D) % 8 || Hq % Hk) {
            delete h;
            return fail(nullptr, "sm_create: gate dims must be multiples of 8 and heads %% kv_heads == 0");
        }
        const std::string p = "model.mm_projector.cls_net.cls_model.";
        h->gate.resize(c.gate_layers);
        for (int l = 0; l < c.gate_layers; ++l) {
            MistralLayer& L = h->gate[l];
            const std::string lp = p + "model.layers." + std::to_string(l) + ".";
            L.in_ln = A(H * e); add_slot(h, lp + "input_layernorm.weight", L.in_ln, 1, H);
            L.post_ln = A(H * e); add_slot(h, lp + "post_attention_layernorm.weight", L.post_ln, 1, H);
            L.wqkv = A(static_cast<size_t>(Hk) * D * H * e); add_slot(h, lp + "self_attn.v_proj.weight", L.wqkv, Hk * D, H);
            L.wo = A(static_cast<size_t>(H) * Hq * D * e); add_slot(h, lp + "self_attn.o_proj.weight", L.wo, H, Hq * D);
            L.wgu = A(static_cast<size_t>(2) * F * H * e);
            add_slot(h, lp + "mlp.gate_proj.weight", L.wgu, F, H);
            add_slot(h, lp + "mlp.up_proj.weight", (char*)L.wgu + static_cast<size_t>(F) * H * e, F, H);
            L.wd = A(static_cast<size_t>(H) * F * e); add_slot(h, lp + "mlp.down_proj.weight", L.wd, H, F);
        }
        h->gt_norm = A(H * e); add_slot(h, p + "model.norm.weight", h->gt_norm, 1, H);
        h->gt_head = A(static_cast<size_t>(2) * H * e); add_slot(h, p + "lm_head.weight", h->gt_head, 2, H);
        h->gt_h = A(4 * H * e); h->gt_v = A(static_cast<size_t>(4) * Hk * D * e); h->gt_m = A(static_cast<size_t>(4) * F * e);
        h->gate_gemm_cap = std::max(Bm, kTowerBatch);
        {
            const size_t R = h->gate_gemm_cap;
            h->gg_h = A(R * H * e); h->gg_hn = A(R * H * e); h->gg_v = A(R * Hk * D * e); h->gg_ve = A(R * Hq * D * e);
            h->gg_gu = A(R * 2 * F * e); h->gg_m = A(R * F * e);
            h->gg_part = static_cast<float*>(A(static_cast<size_t>(8) * R * std::max(H, Hk * D) * sizeof(float)));
        }
        h->gt_logits = static_cast<float*>(A(static_cast<size_t>(std::max(Bm, kTowerBatch)) * 2 * sizeof(float)));
    }
    // ---------------- LLM
    if (c.llm_layers > 0) {
        const int H = c.llm_hidden, Hq = c.llm_heads, Hk = c.llm_kv_heads, D = c.llm_head_dim, F = c.llm_ffn, V = c.llm_vocab;
        if (D != 128 || H % 64 || F % 64 || Hq % Hk || Hq / Hk > 8) {
            delete h;
            return fail(nullptr, "sm_create: LLM needs head_dim 128, hidden/ffn %% 64 == 0, GQA group <= 8");
        }
        const int QKV = (Hq + 2 * Hk) * D;
        h->lm_embed = A(static_cast<size_t>(V) * H * e); add_slot(h, "model.embed_tokens.weight", h->lm_embed, V, H);
        h->llm.resize(c.llm_layers);
        h->kc.resize(c.llm_layers);
        h->vc.resize(c.llm_layers);
        for (int l = 0; l < c.llm_layers; ++l) {
            MistralLayer& L = h->llm[l];
            const std::string lp = "model.layers." + std::to_string(l) + ".";
            L.in_ln = A(H * e); add_slot(h, lp + "input_layernorm.weight", L.in_ln, 1, H);
            L.post_ln = A(H * e); add_slot(h, lp + "post_attention_layernorm.weight", L.post_ln, 1, H);
            L.wqkv = A(static_cast<size_t>(QKV) * H * e);
            add_slot(h, lp + "self_attn.q_proj.weight", L.wqkv, Hq * D, H);
            add_slot(h, lp + "self_attn.k_proj.weight", (char*)L.wqkv + static_cast<size_t>(Hq) * D * H * e, Hk * D, H);
            add_slot(h, lp + "self_attn.v_proj.weight", (char*)L.wqkv + static_cast<size_t>(Hq + Hk) * D * H * e, Hk * D, H);
            L.wo = A(static_cast<size_t>(H) * Hq * D * e); add_slot(h, lp + "self_attn.o_proj.weight", L.wo, H, Hq * D);
            L.wgu = A(static_cast<size_t>(2) * F * H * e);
            add_slot(h, lp + "mlp.gate_proj.weight", L.wgu, F, H);
            add_slot(h, lp + "mlp.up_proj.weight", (char*)L.wgu + static_cast<size_t>(F) * H * e, F, H);
            L.wd = A(static_cast<size_t>(H) * F * e); add_slot(h, lp + "mlp.down_proj.weight", L.wd, H, F);
            h->kc[l] = A(static_cast<size_t>(h->n_streams) * Hk * c.llm_max_ctx * D * e);                // [stream][Hk][max_ctx][D]
            h->vc[l] = A(static_cast<size_t>(h->n_streams) * Hk * c.llm_max_ctx * D * e);
        }
        h->lm_norm = A(H * e); add_slot(h, "model.norm.weight", h->lm_norm, 1, H);
        h->lm_head = A(static_cast<size_t>(V) * H * e); add_slot(h, "lm_head.weight", h->lm_head, V, H);
        h->pmax = 512;
        const size_t Pm = h->pmax;
        h->lw_x = A(Pm * H * e); h->lw_hn = A(Pm * H * e); h->lw_qkv = A(Pm * QKV * e); h->lw_att = A(Pm * Hq * D * e);
        h->lw_gu = A(Pm * 2 * F * e); h->lw_m = A(Pm * F * e);
        h->lw_logits = static_cast<float*>(A(static_cast<size_t>(h->n_streams) * V * sizeof(float)));
        h->kv_stream_stride = static_cast<long long>(Hk) * c.llm_max_ctx * D;
        h->lw_part2 = static_cast<float*>(A(static_cast<size_t>(8) * 64 * std::max(QKV, H) * sizeof(float)));
        if (D == 128 && Hq % Hk == 0 && 128 % (Hq / Hk) == 0 && (128 / (Hq / Hk)) % 8 == 0) {
            const int TB = 128 / (Hq / Hk);
            h->lw_akv_rows = std::max(h->pmax, h->num_sms * TB / Hk + TB);      // n_splits x P never exceeds this (plan_attn_kv_splits)
            h->lw_akv_o = static_cast<float*>(A(static_cast<size_t>(h->lw_akv_rows) * Hq * 128 * sizeof(float)));
            h->lw_akv_ml = static_cast<float*>(A(static_cast<size_t>(h->lw_akv_rows) * Hq * 2 * sizeof(float)));
        }
        // persistent decode kernel: activations of up to kDsMaxStreams lanes, state, split-KV partials, argmax candidates
        const size_t NL = kDsMaxStreams;
        auto LLA = [&](size_t words) { return static_cast<unsigned long long*>(A(words * sizeof(unsigned long long))); };   // zeroed: tag 0 never matches
        h->ds_x_ll = LLA(NL * H / 2); h->ds_qkv_ll = LLA(NL * QKV / 2); h->ds_att_ll = LLA(NL * Hq * D / 2); h->ds_m_ll = LLA(NL * F / 2);
        h->ds_logits = static_cast<float*>(A(NL * V * sizeof(float)));
        h->ds_sync = static_cast<unsigned*>(A(8 * sizeof(unsigned)));
        h->ds_state = static_cast<DsStreamState*>(A(NL * sizeof(DsStreamState)));
        h->ds_out = static_cast<int*>(A(NL * kDsMaxNew * sizeof(int)));
        h->ds_stop = static_cast<int*>(A(64 * sizeof(int)));
        h->ds_att_part = LLA(NL * Hq * h->num_sms * (D + 2));
        h->ds_cand = LLA(NL * h->num_sms * 2);
        h->kv_lens.assign(h->n_streams, 0);
    }
    if (oom) {
        std::string msg = "sm_create: cudaMalloc failed (out of device memory)";
        sm_destroy(h);
        return fail(nullptr, "%s", msg.c_str());
    }
    if (c.llm_layers > 0 && build_decode_ops(h)) {
        g_create_error = h->err;
        sm_destroy(h);
        return 1;
    }
    cudaDeviceSynchronize();
    *out = h;
    return 0;
}

void sm_destroy(sm_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto& g : h->frame_graphs) cudaGraphExecDestroy(g.second);
    for (auto& t : h->ds_pending) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    for (auto st : h->vit_streams) if (st) cudaStreamDestroy(st);
    if (h->gate_stream) cudaStreamDestroy(h->gate_stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (auto e : h->ev_px) if (e) cudaEventDestroy(e);
    if (h->ev_in) cudaEventDestroy(h->ev_in);
    for (auto e : h->ev_vit) if (e) cudaEventDestroy(e);
    for (auto e : h->ev_gate) if (e) cudaEventDestroy(e);
    for (void* p : h->allocs) cudaFree(p);
    delete h;
}

int sm_load_weight(sm_handle* h, const char* name, const void* data, int data_on_host, int dtype, int ndim,
                   const int64_t* shape) {
    if (!h || !name || !data) return fail(h, "sm_load_weight: null argument");
    const std::string n(name);
    auto it = h->slots.find(n);
    if (it == h->slots.end()) {
        // keys that exist in the reference state_dict but do not influence the path
        auto has = [&](const char* t) { return n.find(t) != std::string::npos; };
        if (has("post_layernorm") || has("position_ids") || has("rotary_emb") || has("inv_freq")) return 0;
        if (has("cls_net") && (has("q_proj") || has("k_proj") || has("embed_tokens"))) return 0;
        const size_t lp = n.find("vision_model.encoder.layers.");
        if (lp != std::string::npos && atoi(n.c_str() + lp + strlen("vision_model.encoder.layers.")) >= h->cfg.vit_layers)
            return 0;  // layers after hidden_states[select_layer] (clip_encoder.py:32)
        return fail(h, "sm_load_weight: unknown key '%s'", name);
    }
    if (dtype != h->cfg.dtype) return fail(h, "sm_load_weight: '%s' has dtype %d, handle dtype is %d", name, dtype, h->cfg.dtype);
    Slot& s = it->second;
    size_t numel = 1;
    for (int i = 0; i < ndim; ++i) numel *= static_cast<size_t>(shape[i]);
    if (numel != s.numel) return fail(h, "sm_load_weight: '%s' has %zu elements, expected %zu", name, numel, s.numel);
    cudaSetDevice(h->device);
    const cudaMemcpyKind kind = data_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if (s.tiled) {
        const void* src = data;
        void* tmp = nullptr;
        if (data_on_host) {
            CUDA_OK(h, cudaMalloc(&tmp, numel * h->esz));
            CUDA_OK(h, cudaMemcpy(tmp, data, numel * h->esz, cudaMemcpyHostToDevice));
            src = tmp;
        }
        const int blocks = static_cast<int>(std::min<size_t>((numel + 255) / 256, 8192));
        DISPATCH_T(h, T, {
            retile_weight_kernel<T><<<blocks, 256>>>((const T*)src, (T*)s.tile_base, static_cast<int>(s.rows), s.cols,
                                                     s.tile_row0, s.tile_kb);
        })
        CUDA_OK(h, cudaDeviceSynchronize());
        if (tmp) cudaFree(tmp);
        s.loaded = true;
        return 0;
    }
    CUDA_OK(h, cudaMemcpy2D(s.dst, s.dst_pitch, data, s.row_bytes, s.row_bytes, s.rows, kind));
    s.loaded = true;
    return 0;
}

int sm_finalize_weights(sm_handle* h) {
    if (!h) return 1;
    std::string missing;
    int n = 0;
    for (auto& kv : h->slots)
        if (!kv.second.loaded) {
            if (n < 12) missing += (n ? ", " : "") + kv.first;
            ++n;
        }
    if (n) return fail(h, "sm_finalize_weights: %d weights missing: %s%s", n, missing.c_str(), n > 12 ? ", ..." : "");
    CUDA_OK(h, cudaDeviceSynchronize());
    return 0;
}

int sm_stream_reset(sm_handle* h, void* stream) {
    if (!h) return 1;
    cudaSetDevice(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (pipe_join(h, st)) return 1;            // frames still in flight on the internal streams belong to the old stream state
    if (h->pj_conv_state) {
        const size_t cb = static_cast<size_t>(h->d_inner) * h->cfg.proj_d_conv * h->esz, sb = static_cast<size_t>(h->d_inner) * h->cfg.proj_d_state * sizeof(float);
        CUDA_OK(h, cudaMemsetAsync(static_cast<char*>(h->pj_conv_state) + h->cur * cb, 0, cb, st));
        CUDA_OK(h, cudaMemsetAsync(reinterpret_cast<char*>(h->pj_ssm_state) + h->cur * sb, 0, sb, st));
    }
    if (!h->kv_lens.empty()) h->kv_lens[h->cur] = 0;
    return 0;
}

int sm_stream_select(sm_handle* h, int stream_id) {
    if (!h) return 1;
    if (stream_id < 0 || stream_id >= h->n_streams) return fail(h, "sm_stream_select: stream %d outside [0, %d)", stream_id, h->n_streams);
    if (stream_id == h->cur) return 0;
    cudaSetDevice(h->device);
    if (h->pipe_init && pipe_flush(h)) return 1;     // an open batch of tickets never mixes streams
    h->cur = stream_id;
    return 0;
}

int sm_num_streams(const sm_handle* h) { return h ? h->n_streams : -1; }

int sm_vit_encode(sm_handle* h, const void* pixels, int B, void* feats_out, void* pooled_out, void* stream) {
    if (!h || h->cfg.vit_layers <= 0) return fail(h, "sm_vit_encode: vision tower not configured");
    if (B < 1 || B > h->cfg.max_frames) return fail(h, "sm_vit_encode: B=%d outside [1, max_frames=%d]", B, h->cfg.max_frames);
    cudaSetDevice(h->device);
    if (pipe_join(h, static_cast<cudaStream_t>(stream))) return 1;
    return run_vit(h, pixels, B, feats_out, pooled_out, static_cast<cudaStream_t>(stream));
}

int sm_preprocess_frames(sm_handle* h, const unsigned char* frames, int n, int H, int W, int frames_on_device, const float* mean,
                         const float* std_, const int* background, void* pixels_out, void* stream) {
    if (!h || !frames || !mean || !std_ || !background || !pixels_out) return fail(h, "sm_preprocess_frames: null argument");
    if (n < 1 || H < 1 || W < 1) return fail(h, "sm_preprocess_frames: n=%d H=%d W=%d", n, H, W);
    const int out = h->cfg.vit_image;
    if (out <= 0) return fail(h, "sm_preprocess_frames: vision tower not configured");
    cudaSetDevice(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int S = std::max(H, W);
    // buffers grow only; a grown buffer is replaced after the stream has drained (earlier calls may still read it)
    auto grow = [&](void*& p, size_t& have, size_t need) -> int {
        if (need <= have) return 0;
        CUDA_OK(h, cudaStreamSynchronize(st));
        if (p) {
            CUDA_OK(h, cudaFree(p));
            auto it = std::find(h->allocs.begin(), h->allocs.end(), p);
            if (it != h->allocs.end()) h->allocs.erase(it);
            p = nullptr; have = 0;
        }
        CUDA_OK(h, cudaMalloc(&p, need));
        h->allocs.push_back(p);
        have = need;
        return 0;
    };
    auto tab = h->pre_tables.find(S);
    if (tab == h->pre_tables.end()) {
        std::vector<int> bounds, kk;
        sm_handle::PreTable t{};
        t.ksize = pre_build_table(S, out, bounds, kk);
        CUDA_OK(h, cudaMalloc(reinterpret_cast<void**>(&t.bounds), bounds.size() * sizeof(int)));
        CUDA_OK(h, cudaMalloc(reinterpret_cast<void**>(&t.kk), kk.size() * sizeof(int)));
        CUDA_OK(h, cudaMalloc(reinterpret_cast<void**>(&t.kk_t), kk.size() * sizeof(int)));
        h->allocs.push_back(t.bounds); h->allocs.push_back(t.kk); h->allocs.push_back(t.kk_t);
        std::vector<int> kk_t(kk.size());
        for (int xx = 0; xx < out; ++xx)
            for (int x = 0; x < t.ksize; ++x) kk_t[static_cast<size_t>(x) * out + xx] = kk[static_cast<size_t>(xx) * t.ksize + x];
        CUDA_OK(h, cudaMemcpy(t.kk_t, kk_t.data(), kk_t.size() * sizeof(int), cudaMemcpyHostToDevice));
        CUDA_OK(h, cudaMemcpy(t.bounds, bounds.data(), bounds.size() * sizeof(int), cudaMemcpyHostToDevice));
        CUDA_OK(h, cudaMemcpy(t.kk, kk.data(), kk.size() * sizeof(int), cudaMemcpyHostToDevice));
        tab = h->pre_tables.emplace(S, t).first;
    }
    // rescale (uint8 * (1/255) in double -> float) + normalize ((x - mean) / std in float) + rounding to the model dtype,
    // as transformers' image_transforms.rescale / normalize and the caller's .half() do: 3 x 256 possible results
    const float key[6] = {mean[0], mean[1], mean[2], std_[0], std_[1], std_[2]};
    if (!h->pre_lut || memcmp(key, h->pre_lut_key, sizeof key) != 0) {
        std::vector<uint16_t> lut(3 * 256);
        for (int c = 0; c < 3; ++c)
            for (int v = 0; v < 256; ++v) {
                const float x = static_cast<float>(static_cast<double>(v) * (1.0 / 255));
                const float y = (x - mean[c]) / std_[c];
                if (h->cfg.dtype == SM_DTYPE_BF16) { const __nv_bfloat16 t = __float2bfloat16_rn(y); memcpy(&lut[c * 256 + v], &t, 2); }
                else { const __half t = __float2half_rn(y); memcpy(&lut[c * 256 + v], &t, 2); }
            }
        CUDA_OK(h, cudaStreamSynchronize(st));
        if (!h->pre_lut) { CUDA_OK(h, cudaMalloc(&h->pre_lut, lut.size() * 2)); h->allocs.push_back(h->pre_lut); }
        CUDA_OK(h, cudaMemcpy(h->pre_lut, lut.data(), lut.size() * 2, cudaMemcpyHostToDevice));
        memcpy(h->pre_lut_key, key, sizeof key);
    }
    const size_t src_bytes = static_cast<size_t>(n) * H * W * 3;
    const uint8_t* src = frames;
    if (!frames_on_device) {
        if (grow(h->pre_src, h->pre_src_bytes, src_bytes)) return 1;
        CUDA_OK(h, cudaMemcpyAsync(h->pre_src, frames, src_bytes, cudaMemcpyHostToDevice, st));
        src = static_cast<const uint8_t*>(h->pre_src);
    }
    if (grow(h->pre_tmp, h->pre_tmp_bytes, static_cast<size_t>(n) * S * out * 3)) return 1;
    PreArgs a{};
    a.src = src; a.tmp = static_cast<uint8_t*>(h->pre_tmp); a.bounds = tab->second.bounds; a.kk = tab->second.kk; a.kk_t = tab->second.kk_t;
    a.H = H; a.W = W; a.S = S; a.out = out; a.ksize = tab->second.ksize;
    a.pad_x = H > W ? (H - W) / 2 : 0;       // expand2square: paste at ((height - width) // 2, 0) / (0, (width - height) // 2)
    a.pad_y = W > H ? (W - H) / 2 : 0;
    a.bg0 = background[0]; a.bg1 = background[1]; a.bg2 = background[2];
    const size_t row_smem = (static_cast<size_t>(W) * 3 + 15) & ~size_t(15);
    if (row_smem > 48 * 1024) return fail(h, "sm_preprocess_frames: frames wider than 16384 pixels are not supported (W=%d)", W);
    preprocess_h_kernel<<<dim3(S, n), 128, row_smem, st>>>(a);
    count_launch(h);
    DISPATCH_T(h, T, {
        if (out % 4 == 0) {
            preprocess_v4_kernel<T><<<dim3((out * 3 / 4 + 127) / 128, out, n), 128, 0, st>>>(a, static_cast<const T*>(h->pre_lut), static_cast<T*>(pixels_out));
        } else {
            preprocess_v_kernel<T><<<dim3((out + 127) / 128, out, n), 128, 0, st>>>(a, static_cast<const T*>(h->pre_lut), static_cast<T*>(pixels_out));
        }
        count_launch(h);
    })
    CUDA_OK(h, cudaGetLastError());
    return 0;
}

int sm_resample_table(int in_size, int out_size, int* ksize_out, int* bounds_out, int* kk_out, long long kk_capacity) {
    if (in_size < 1 || out_size < 1 || !ksize_out) return 1;
    std::vector<int> bounds, kk;
    *ksize_out = pre_build_table(in_size, out_size, bounds, kk);
    if (bounds_out) memcpy(bounds_out, bounds.data(), bounds.size() * sizeof(int));
    if (kk_out) {
        if (kk_capacity < static_cast<long long>(kk.size())) return 2;
        memcpy(kk_out, kk.data(), kk.size() * sizeof(int));
    }
    return 0;
}

int sm_pool_features(sm_handle* h, const void* feats, int n, void* pooled_out, void* stream) {
    if (!h || h->cfg.vit_hidden <= 0) return fail(h, "sm_pool_features: not configured");
    cudaSetDevice(h->device);
    const int C = h->cfg.vit_hidden, P = h->P;
    DISPATCH_T(h, T, {
        pool_kernel<T><<<dim3((C + 127) / 128, n), 128, 0, static_cast<cudaStream_t>(stream)>>>((const T*)feats, (T*)pooled_out, P, C);
        count_launch(h);
    })
    CUDA_OK(h, cudaGetLastError());
    return 0;
}

int sm_projector_step(sm_handle* h, const void* pooled, int n, void* tok_out, void* stream) {
    if (!h || h->cfg.proj_d_model <= 0) return fail(h, "sm_projector_step: projector not configured");
    cudaSetDevice(h->device);
    if (pipe_join(h, static_cast<cudaStream_t>(stream))) return 1;
    for (int i = 0; i < n; i += kGemvBatch) {
        const char* src = static_cast<const char*>(pooled) + static_cast<size_t>(i) * h->cfg.vit_hidden * h->esz;
        char* dst = static_cast<char*>(tok_out) + static_cast<size_t>(i) * h->cfg.proj_d_model * h->esz;
        if (run_projector(h, src, dst, std::min(kGemvBatch, n - i), static_cast<cudaStream_t>(stream))) return 1;
    }
    return 0;
}

int sm_gate_score(sm_handle* h, const void* tok, float* logits_out, void* stream) {
    if (!h || h->cfg.gate_layers <= 0) return fail(h, "sm_gate_score: gate not configured");
    cudaSetDevice(h->device);
    if (pipe_join(h, static_cast<cudaStream_t>(stream))) return 1;
    return run_gate(h, tok, logits_out, 1, static_cast<cudaStream_t>(stream));
}

int sm_frame_step(sm_handle* h, const void* pixels, int pixels_on_host, int B, void* feats_out, void* toks_out,
                  float* logits_out, float* logits_host, void* stream) {
    if (!h || h->cfg.vit_layers <= 0 || h->cfg.proj_d_model <= 0 || h->cfg.gate_layers <= 0)
        return fail(h, "sm_frame_step: needs vision tower + projector + gate");
    if (B < 1 || B > h->cfg.max_frames) return fail(h, "sm_frame_step: B=%d outside [1, max_frames=%d]", B, h->cfg.max_frames);
    cudaSetDevice(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (pipe_join(h, st)) return 1;
    const sm_config& c = h->cfg;
    const size_t px_bytes = static_cast<size_t>(B) * 3 * c.vit_image * c.vit_image * h->esz;
    CUDA_OK(h, cudaMemcpyAsync(h->ws_pixels, pixels, px_bytes, pixels_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
    const bool want_feats = feats_out != nullptr;
    auto body = [&](cudaStream_t s) -> int {
        if (run_vit(h, h->ws_pixels, B, want_feats ? h->ws_feats : nullptr, h->ws_pooled, s)) return 1;
        return run_proj_gate(h, h->ws_pooled, h->pj_toks, h->gt_logits, B, s);
    };
    if (c.use_graphs) {
        const int key = B | (want_feats ? 1 << 8 : 0) | static_cast<int>((h->kfilter & 0xFFFFu) << 9);
        auto it = h->frame_graphs.find(gkey(h, key));
        if (it == h->frame_graphs.end()) {
            // warm run outside capture: fills the tensor-map cache and sets function attributes
            if (body(st)) return 1;
            CUDA_OK(h, cudaStreamSynchronize(st));
            // the warm run advanced the Mamba state; rewind is the caller's job only on the very first
            // call, so capture must not run the kernels: stream capture records without executing.
            cudaGraph_t g;
            h->capturing = true;
            h->captured_launches = 0;
            CUDA_OK(h, cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
            const int rc = body(h->cap_stream);
            cudaError_t ce = cudaStreamEndCapture(h->cap_stream, &g);
            h->capturing = false;
            if (rc || ce != cudaSuccess) return fail(h, "sm_frame_step: graph capture failed: %s", cudaGetErrorString(ce));
            cudaGraphExec_t ge;
            CUDA_OK(h, cudaGraphInstantiate(&ge, g, 0));
            cudaGraphDestroy(g);
            h->frame_graphs[gkey(h, key)] = ge;
            h->frame_graph_launches[gkey(h, key)] = h->captured_launches;
            // the warm run already produced this call's outputs
        } else {
            CUDA_OK(h, cudaGraphLaunch(it->second, st));
            h->launches += h->frame_graph_launches[gkey(h, key)];
        }
    } else {
        if (body(st)) return 1;
    }
    if (feats_out) CUDA_OK(h, cudaMemcpyAsync(feats_out, h->ws_feats, static_cast<size_t>(B) * h->P * c.vit_hidden * h->esz, cudaMemcpyDeviceToDevice, st));
    if (toks_out) CUDA_OK(h, cudaMemcpyAsync(toks_out, h->pj_toks, static_cast<size_t>(B) * c.proj_d_model * h->esz, cudaMemcpyDeviceToDevice, st));
    if (logits_out) CUDA_OK(h, cudaMemcpyAsync(logits_out, h->gt_logits, static_cast<size_t>(B) * 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (logits_host) CUDA_OK(h, cudaMemcpyAsync(logits_host, h->gt_logits, static_cast<size_t>(B) * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
    return 0;
}

int sm_frame_step_multi(sm_handle* h, const void* pixels, int pixels_on_host, int n, int first_stream, void* toks_out,
                        float* logits_out, float* logits_host, void* stream) {
    if (!h || h->cfg.vit_layers <= 0 || h->cfg.proj_d_model <= 0 || h->cfg.gate_layers <= 0)
        return fail(h, "sm_frame_step_multi: needs vision tower + projector + gate");
    if (n < 1 || n > h->cfg.max_frames) return fail(h, "sm_frame_step_multi: n=%d outside [1, max_frames=%d]", n, h->cfg.max_frames);
    if (first_stream < 0 || first_stream + n > h->n_streams) return fail(h, "sm_frame_step_multi: streams [%d, %d) outside [0, %d)", first_stream, first_stream + n, h->n_streams);
    cudaSetDevice(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (pipe_join(h, st)) return 1;
    const sm_config& c = h->cfg;
    const size_t px_bytes = static_cast<size_t>(n) * 3 * c.vit_image * c.vit_image * h->esz;
    CUDA_OK(h, cudaMemcpyAsync(h->ws_pixels, pixels, px_bytes, pixels_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
    if (run_vit(h, h->ws_pixels, n, nullptr, h->ws_pooled, st)) return 1;
    const int saved = h->cur;
    int rc = 0;
    for (int i = 0; i < n && !rc; i += kGemvBatch) {          // <= 4 streams share each pass over the projector weights
        const int nv = std::min(kGemvBatch, n - i);
        h->cur = first_stream + i;
        rc = run_projector(h, static_cast<const char*>(h->ws_pooled) + static_cast<size_t>(i) * c.vit_hidden * h->esz,
                           static_cast<char*>(h->pj_toks) + static_cast<size_t>(i) * c.proj_d_model * h->esz, nv, st, true);
    }
    h->cur = saved;
    if (rc) return 1;
    // the gate is stateless: the n frames are n rows of one weight pass (GEMMs from 5 rows on, else the batched GEMV chain)
    static const int gemm_min = getenv("SMB_GATE_GEMM") ? atoi(getenv("SMB_GATE_GEMM")) : 5;
    if (gemm_min > 0 && n >= gemm_min && n <= h->gate_gemm_cap) {
        if (run_gate_gemm(h, h->pj_toks, h->gt_logits, n, st)) return 1;
    } else {
        for (int i = 0; i < n; i += kGemvBatch)
            if (run_gate(h, static_cast<const char*>(h->pj_toks) + static_cast<size_t>(i) * c.proj_d_model * h->esz, h->gt_logits + 2 * i,
                         std::min(kGemvBatch, n - i), st)) return 1;
    }
    if (toks_out) CUDA_OK(h, cudaMemcpyAsync(toks_out, h->pj_toks, static_cast<size_t>(n) * c.proj_d_model * h->esz, cudaMemcpyDeviceToDevice, st));
    if (logits_out) CUDA_OK(h, cudaMemcpyAsync(logits_out, h->gt_logits, static_cast<size_t>(n) * 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (logits_host) CUDA_OK(h, cudaMemcpyAsync(logits_host, h->gt_logits, static_cast<size_t>(n) * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
    return 0;
}

int sm_frame_submit(sm_handle* h, const void* pixels, int pixels_on_host, int B, void* feats_out, void* toks_out,
                    float* logits_out, float* logits_host, void* stream, long long* ticket_out) {
    if (!h || h->cfg.vit_layers <= 0 || h->cfg.proj_d_model <= 0 || h->cfg.gate_layers <= 0)
        return fail(h, "sm_frame_submit: needs vision tower + projector + gate");
    if (B < 1 || B > h->cfg.max_frames) return fail(h, "sm_frame_submit: B=%d outside [1, max_frames=%d]", B, h->cfg.max_frames);
    cudaSetDevice(h->device);
    const sm_config& c = h->cfg;
    if (!h->pipe_init) {
        int lo = 0, hi = 0;
        CUDA_OK(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));   // lo = least, hi = greatest priority
        for (auto& vst : h->vit_streams) CUDA_OK(h, cudaStreamCreateWithPriority(&vst, cudaStreamNonBlocking, hi));
        CUDA_OK(h, cudaStreamCreateWithPriority(&h->copy_stream, cudaStreamNonBlocking, hi));
        CUDA_OK(h, cudaStreamCreateWithPriority(&h->gate_stream, cudaStreamNonBlocking, lo));
        CUDA_OK(h, cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
        for (auto& e : h->ev_vit) CUDA_OK(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : h->ev_gate) CUDA_OK(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : h->ev_px) CUDA_OK(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        // streaming engines (max_frames == 1): the towers of tower_batch consecutive tickets run as one chunk on one of
        // (up to 4) lanes; otherwise every ticket runs its own tower on one of (up to 8) lanes
        h->tower_batch = c.max_frames == 1 ? std::max(1, std::min(kTowerBatch, getenv("SMB_TOWER_BATCH") ? atoi(getenv("SMB_TOWER_BATCH")) : kTowerBatch)) : 1;
        if (h->tower_batch == 3 || (h->tower_batch > 4 && h->tower_batch < 8)) h->tower_batch = 4;   // groups must tile the ticket ring
        const int lanes_default = h->tower_batch > 1 ? 2 : 8;
        h->n_lanes = std::max(1, std::min(h->tower_batch > 1 ? 3 : kMaxLanes, getenv("SMB_LANES") ? atoi(getenv("SMB_LANES")) : lanes_default));
        h->gate_batch = std::max(1, std::min(kGemvBatch, getenv("SMB_GATE_BATCH") ? atoi(getenv("SMB_GATE_BATCH")) : kGemvBatch));
        if (h->gate_batch == 3) h->gate_batch = 2;   // groups must tile the ticket ring
        h->pipe_init = true;
    }
    const long long tk = h->ticket;
    const int ring = static_cast<int>(tk % kTicketRing);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the ring slot (events, pooled vector, staged pixels) of ticket tk - kTicketRing is reused: its gate must be done
    if (tk >= kTicketRing) {
        if (h->n_pending > 0 && h->first_pending <= tk - kTicketRing && pipe_flush(h)) return 1;
        CUDA_OK(h, cudaEventSynchronize(h->ev_gate[ring]));
    }
    CUDA_OK(h, cudaEventRecord(h->ev_in, st));
    if (h->n_pending == 0) CUDA_OK(h, cudaStreamWaitEvent(h->gate_stream, h->ev_in, 0));   // earlier serial calls on `stream` precede this batch
    const size_t px_bytes = static_cast<size_t>(B) * 3 * c.vit_image * c.vit_image * h->esz;
    const cudaMemcpyKind kind = pixels_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if (h->n_pending == 0) h->first_pending = tk;

    if (h->tower_batch > 1) {
        // ---- tower-batch mode: stage the pixels; the tower runs when the batch closes (pipe_flush)
        CUDA_OK(h, cudaStreamWaitEvent(h->copy_stream, h->ev_in, 0));
        CUDA_OK(h, cudaMemcpyAsync(static_cast<char*>(h->px_ring) + static_cast<size_t>(ring) * px_bytes, pixels, px_bytes, kind, h->copy_stream));
        CUDA_OK(h, cudaEventRecord(h->ev_px[ring], h->copy_stream));
        h->pend[h->n_pending++] = {feats_out, toks_out, logits_out, logits_host, B};
        if (ticket_out) *ticket_out = tk;
        h->ticket = tk + 1;
        if (h->n_pending >= h->tower_batch || (tk + 1) % h->tower_batch == 0) return pipe_flush(h);
        return 0;
    }

    // ---- one tower per ticket, on lane tk mod n_lanes
    const int lane = static_cast<int>(tk % h->n_lanes);
    cudaStream_t vs = h->vit_streams[lane];
    {
        PipePlanScope plan(h);
        select_lane(h, lane);
        CUDA_OK(h, cudaStreamWaitEvent(vs, h->ev_in, 0));                            // inputs are ready on the caller's stream
        CUDA_OK(h, cudaMemcpyAsync(h->ws_pixels, pixels, px_bytes, kind, vs));
        const bool want_feats = feats_out != nullptr;
        void* pooled = static_cast<char*>(h->pooled_ring) + static_cast<size_t>(ring) * c.max_frames * c.vit_hidden * h->esz;
        auto vit_body = [&](cudaStream_t s) -> int { return run_vit(h, h->ws_pixels, B, want_feats ? h->ws_feats : nullptr, pooled, s); };
        const int key = B | (want_feats ? 1 << 8 : 0) | (1 << 9) | (lane << 28) | (ring << 24) | static_cast<int>((h->kfilter & 0xFFFu) << 12);
        if (pipe_run_part(h, key, vs, vit_body, "sm_frame_submit(tower)")) return 1;
        if (feats_out) CUDA_OK(h, cudaMemcpyAsync(feats_out, h->ws_feats, static_cast<size_t>(B) * h->P * c.vit_hidden * h->esz, cudaMemcpyDeviceToDevice, vs));
        CUDA_OK(h, cudaEventRecord(h->ev_vit[ring], vs));
    }
    // projector + gate: batched over up to gate_batch consecutive single-frame tickets (aligned groups, so a batch
    // never wraps around the ring); multi-frame tickets are batched inside the call
    h->pend[h->n_pending++] = {nullptr, toks_out, logits_out, logits_host, B};
    if (ticket_out) *ticket_out = tk;
    h->ticket = tk + 1;
    const bool batchable = B == 1 && c.max_frames == 1 && h->gate_batch > 1;   // ring slots are then contiguous single vectors
    if (!batchable || h->n_pending >= h->gate_batch || (tk + 1) % h->gate_batch == 0) {
        if (pipe_flush(h)) return 1;
    }
    return 0;
}

int sm_frame_wait(sm_handle* h, long long ticket, void* stream, int block_host) {
    if (!h || !h->pipe_init) return fail(h, "sm_frame_wait: nothing submitted");
    if (ticket < 0 || ticket >= h->ticket || ticket + kTicketRing < h->ticket)
        return fail(h, "sm_frame_wait: ticket %lld is not in flight (next ticket %lld, ring of %d)", ticket, h->ticket, kTicketRing);
    cudaSetDevice(h->device);
    if (h->n_pending > 0 && ticket >= h->first_pending && pipe_flush(h)) return 1;   // its gate batch is still open
    cudaEvent_t ev = h->ev_gate[ticket % kTicketRing];
    if (stream != nullptr || !block_host) CUDA_OK(h, cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), ev, 0));
    if (block_host) CUDA_OK(h, cudaEventSynchronize(ev));
    return 0;
}

// torch.linspace(0, n - 1, steps).int() as the reference gets it: the call has no device argument, so it is ATen's CPU
// kernel (aten/src/ATen/native/cpu/RangeFactoriesKernel.cpp) whatever device the model is on.  In float32:
// step = (end - start) / (steps - 1); element i is start + step * i below steps / 2 and end - step * (steps - i - 1) from
// there on, truncated toward zero by .int().  ATen's AVX2 / AVX512 builds contract each multiply-add into ONE fused
// multiply-add, its DEFAULT build (no FMA hardware) rounds twice -- the two differ whenever the product lands within an
// ulp of an integer (n = 17, steps = 15: index 7 is 7 with FMA, 8 without), so the host's capability is part of the rule.
static void linspace_indices(int n, int steps, bool fused, int* out) {
    const float start = 0.f, end = static_cast<float>(n - 1);
    if (steps == 1) { out[0] = 0; return; }
    const float step = (end - start) / static_cast<float>(steps - 1);
    const int halfway = steps / 2;
    for (int i = 0; i < steps; ++i) {
        const float a = i < halfway ? step : -step, b = static_cast<float>(i < halfway ? i : steps - i - 1), c = i < halfway ? start : end;
        volatile float prod = a * b;                       // volatile: keeps the unfused product a separately rounded float
        out[i] = static_cast<int>(fused ? fmaf(a, b, c) : prod + c);
    }
}

// does ATen on this host run its AVX2 / AVX512 (FMA) kernels?  (torch.backends.cpu.get_cpu_capability() != "DEFAULT")
static bool host_fused_multiply_add() {
    __builtin_cpu_init();
    return __builtin_cpu_supports("fma") && __builtin_cpu_supports("avx2");
}

int sm_cognition_count(int n, double percentage, int mode) {
    if (n < 1) return 0;
    const int k = static_cast<int>(percentage * n);           // Python: int(percentage * n), float64 product
    return mode == 0 ? (k == 0 ? 1 : k) : std::max(k, 1);
}

int sm_linspace_indices(int n, int steps, int fused, int* out) {
    if (n < 1 || steps < 1 || !out || fused < -1 || fused > 1) return 1;
    linspace_indices(n, steps, fused < 0 ? host_fused_multiply_add() : fused == 1, out);
    return 0;
}

int sm_cognition_sample(sm_handle* h, const void* toks, int n, int d, int mode, double percentage, void* out, int32_t* indices_out,
                        void* stream) {
    if (!h || !toks || !out) return fail(h, "sm_cognition_sample: null argument");
    if (n < 1 || d < 1) return fail(h, "sm_cognition_sample: n=%d d=%d", n, d);
    if (mode != 0 && mode != 1) return fail(h, "sm_cognition_sample: mode %d (0 = linspace, 1 = similarity)", mode);
    cudaSetDevice(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int k = sm_cognition_count(n, percentage, mode);
    // scratch: indices [n] + similarities [n]
    const size_t need = static_cast<size_t>(n) * (sizeof(int) + sizeof(float));
    if (need > h->cog_bytes) {
        CUDA_OK(h, cudaStreamSynchronize(st));
        if (h->cog_buf) { cudaFree(h->cog_buf); h->allocs.erase(std::find(h->allocs.begin(), h->allocs.end(), h->cog_buf)); }
        CUDA_OK(h, cudaMalloc(&h->cog_buf, need));
        h->allocs.push_back(h->cog_buf);
        h->cog_bytes = need;
    }
    int* d_idx = static_cast<int*>(h->cog_buf);
    float* d_sim = reinterpret_cast<float*>(d_idx + n);
    if (mode == 0) {
        std::vector<int> idx(k);
        linspace_indices(n, k, host_fused_multiply_add(), idx.data());
        CUDA_OK(h, cudaMemcpyAsync(d_idx, idx.data(), sizeof(int) * k, cudaMemcpyHostToDevice, st));
        CUDA_OK(h, cudaStreamSynchronize(st));                 // idx is a stack-lifetime host buffer
    } else {
        if (static_cast<size_t>(n) * sizeof(int) > 200 * 1024) return fail(h, "sm_cognition_sample: at most %d tokens", 200 * 1024 / 4);
        DISPATCH_T(h, T, {
            cos_sim_rows_kernel<T><<<(n + 7) / 8, 256, 0, st>>>((const T*)toks, n, d, 1e-8f, d_sim);
            count_launch(h);
        })
        CUDA_OK(h, cudaFuncSetAttribute(topk_keep_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        topk_keep_order_kernel<<<1, 1024, static_cast<size_t>(n) * sizeof(int), st>>>(d_sim, n, k, d_idx);
        count_launch(h);
    }
    DISPATCH_T(h, T, {
        gather_rows_kernel<T><<<k, 256, 0, st>>>((const T*)toks, d_idx, (T*)out, k, d);
        count_launch(h);
    })
    if (indices_out) CUDA_OK(h, cudaMemcpyAsync(indices_out, d_idx, sizeof(int) * k, cudaMemcpyDeviceToDevice, st));
    CUDA_OK(h, cudaGetLastError());
    return 0;
}

int sm_embed_tokens(sm_handle* h, const int32_t* ids, int n, void* out, void* stream) {
    if (!h || h->cfg.llm_layers <= 0) return fail(h, "sm_embed_tokens: LLM not configured");
    cudaSetDevice(h->device);
    DISPATCH_T(h, T, {
        gather_rows_kernel<T><<<n, 256, 0, static_cast<cudaStream_t>(stream)>>>((const T*)h->lm_embed, ids, (T*)out, n, h->cfg.llm_hidden);
        count_launch(h);
    })
    CUDA_OK(h, cudaGetLastError());
    return 0;
}

int sm_llm_prefill(sm_handle* h, const void* embeds, int P, float* last_logits, void* stream) {
    if (!h || h->cfg.llm_layers <= 0) return fail(h, "sm_llm_prefill: LLM not configured");
    if (P < 1) return fail(h, "sm_llm_prefill: P must be >= 1");
    int& kv_len = h->kv_lens[h->cur];
    if (kv_len + P > h->cfg.llm_max_ctx) return fail(h, "sm_llm_prefill: %d + %d exceeds llm_max_ctx %d", kv_len, P, h->cfg.llm_max_ctx);
    cudaSetDevice(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const sm_config& c = h->cfg;
    int done = 0, last_chunk = 0;
    while (done < P) {
        const int n = std::min(h->pmax, P - done);
        const char* src = static_cast<const char*>(embeds) + static_cast<size_t>(done) * c.llm_hidden * h->esz;
        if (run_prefill_chunk(h, src, n, kv_len, st)) return 1;
        kv_len += n;
        done += n;
        last_chunk = n;
    }
    // final norm + lm_head on the last position only (generate() needs nothing else)
    const char* xlast = static_cast<const char*>(h->lw_x) + static_cast<size_t>(last_chunk - 1) * c.llm_hidden * h->esz;
    float* logits = h->lw_logits + static_cast<size_t>(h->cur) * c.llm_vocab;      // kept per stream until its decode call
    GemvArgs a = gv(h->lm_head, c.llm_vocab, c.llm_hidden, PRO_RMSNORM, xlast, GEPI_F32, logits);
    a.nw = h->lm_norm; a.eps = c.llm_eps;
    if (launch_gemv(h, a, 1, st)) return 1;
    if (last_logits) CUDA_OK(h, cudaMemcpyAsync(last_logits, logits, static_cast<size_t>(c.llm_vocab) * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

int sm_llm_decode_multi(sm_handle* h, int n, const int* stream_ids, const int* max_new, const int32_t* stop_ids, int n_stop,
                        int32_t* ids_out_host, int out_stride, int32_t* n_out_host, void* stream) {
    if (!h || h->cfg.llm_layers <= 0) return fail(h, "sm_llm_decode: LLM not configured");
    if (n < 1 || n > kDsMaxStreams) return fail(h, "sm_llm_decode: %d streams per pass outside [1, %d]", n, kDsMaxStreams);
    if (!stream_ids || !max_new || !ids_out_host || !n_out_host) return fail(h, "sm_llm_decode: null argument");
    if (n_stop < 0 || n_stop > 63) return fail(h, "sm_llm_decode: at most 63 stop ids");
    cudaSetDevice(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const sm_config& c = h->cfg;
    DsStreamState hs[kDsMaxStreams] = {};
    int steps = 0;
    long long ctx_sum = 0, tok_sum = 0;
    for (int i = 0; i < n; ++i) {
        const int s = stream_ids[i];
        if (s < 0 || s >= h->n_streams) return fail(h, "sm_llm_decode: stream %d outside [0, %d)", s, h->n_streams);
        for (int j = 0; j < i; ++j)
            if (stream_ids[j] == s) return fail(h, "sm_llm_decode: stream %d listed twice", s);
        if (max_new[i] < 1) return fail(h, "sm_llm_decode: max_new must be >= 1");
        if (h->kv_lens[s] < 1) return fail(h, "sm_llm_decode: stream %d has no prefilled context (call sm_llm_prefill first)", s);
        // a stream at the end of its cache produces what still fits (the token fed at position p needs slot p)
        const int room = c.llm_max_ctx - h->kv_lens[s] + 1;
        const int mn = std::min({max_new[i], room, kDsMaxNew, out_stride});
        if (mn < 1) return fail(h, "sm_llm_decode: KV cache of stream %d is full (%d of %d)", s, h->kv_lens[s], c.llm_max_ctx);
        hs[i].pos = h->kv_lens[s]; hs[i].max_new = mn; hs[i].kv_slot = s;
        steps = std::max(steps, mn - 1);
    }
    int stopbuf[64] = {};
    stopbuf[0] = n_stop;
    for (int i = 0; i < n_stop; ++i) stopbuf[1 + i] = stop_ids[i];
    CUDA_OK(h, cudaMemcpyAsync(h->ds_stop, stopbuf, sizeof stopbuf, cudaMemcpyHostToDevice, st));
    CUDA_OK(h, cudaMemcpyAsync(h->ds_state, hs, sizeof(DsStreamState) * n, cudaMemcpyHostToDevice, st));
    CUDA_OK(h, cudaMemsetAsync(h->ds_sync + 2, 0, sizeof(unsigned), st));      // the all-done flag; the epoch in [1] keeps counting
    ds_first_token_kernel<<<1, 1024, 0, st>>>(h->lw_logits, c.llm_vocab, c.llm_vocab, n, h->ds_state, h->ds_out, kDsMaxNew, h->ds_stop, h->ds_sync);
    count_launch(h);
    CUDA_OK(h, cudaGetLastError());
    // one launch per token; with stop ids the host looks at the all-done flag every 16 steps (a finished call turns the
    // remaining launches into no-ops: the kernel returns at its first instruction)
    sm_handle::DsTiming tm{};
    CUDA_OK(h, cudaEventCreate(&tm.a));
    CUDA_OK(h, cudaEventCreate(&tm.b));
    CUDA_OK(h, cudaEventRecord(tm.a, st));
    int launched = 0;
    unsigned host_done = 0;
    const int check_every = n_stop > 0 ? 16 : steps;
    while (launched < steps && !host_done) {
        const int burst = std::min(check_every, steps - launched);
        for (int i = 0; i < burst; ++i)
            if (launch_decode_step(h, n, st)) return 1;
        launched += burst;
        if (n_stop > 0 && launched < steps) {
            CUDA_OK(h, cudaMemcpyAsync(&host_done, h->ds_sync + 2, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            CUDA_OK(h, cudaStreamSynchronize(st));
        }
    }
    CUDA_OK(h, cudaEventRecord(tm.b, st));
    CUDA_OK(h, cudaMemcpyAsync(hs, h->ds_state, sizeof(DsStreamState) * n, cudaMemcpyDeviceToHost, st));
    CUDA_OK(h, cudaStreamSynchronize(st));
    for (int i = 0; i < n; ++i) {
        const int nout = std::min(hs[i].n_out, hs[i].max_new);
        CUDA_OK(h, cudaMemcpy(ids_out_host + static_cast<size_t>(i) * out_stride, h->ds_out + static_cast<size_t>(i) * kDsMaxNew,
                              static_cast<size_t>(nout) * sizeof(int), cudaMemcpyDeviceToHost));
        n_out_host[i] = nout;
        ctx_sum += static_cast<long long>(nout - 1) * (h->kv_lens[hs[i].kv_slot] + hs[i].pos + 1) / 2;   // sum over the nout - 1 steps of the KV length each one read
        tok_sum += nout - 1;
        h->kv_lens[hs[i].kv_slot] = hs[i].pos;
    }
    int steps_run = 0;
    for (int i = 0; i < n; ++i) steps_run = std::max(steps_run, std::min(hs[i].n_out, hs[i].max_new) - 1);
    tm.steps = steps_run; tm.tokens = tok_sum; tm.ctx_sum = ctx_sum;
    h->ds_pending.push_back(tm);
    ds_collect_timings(h, false);
    return 0;
}

int sm_llm_decode(sm_handle* h, int max_new, const int32_t* stop_ids, int n_stop, int32_t* ids_out_host,
                  int32_t* n_out_host, void* stream) {
    if (!h) return 1;
    if (max_new < 1 || max_new > kDsMaxNew) return fail(h, "sm_llm_decode: max_new must be in [1, %d]", kDsMaxNew);
    const int sid = h->cur;
    return sm_llm_decode_multi(h, 1, &sid, &max_new, stop_ids, n_stop, ids_out_host, max_new, n_out_host, stream);
}

int sm_decode_stats(sm_handle* h, double* ms, long long* steps, long long* tokens, long long* ctx_sum, int reset) {
    if (!h) return 1;
    cudaSetDevice(h->device);
    ds_collect_timings(h, true);
    if (ms) *ms = h->ds_ms;
    if (steps) *steps = h->ds_steps;
    if (tokens) *tokens = h->ds_tokens;
    if (ctx_sum) *ctx_sum = h->ds_ctx_sum;
    if (reset) { h->ds_ms = 0.0; h->ds_steps = h->ds_tokens = h->ds_ctx_sum = 0; }
    return 0;
}

int sm_debug_decode_phases(sm_handle* h, long long* device_buf) {
    if (!h) return 1;
    h->ds_dbg = device_buf;
    return 0;
}

int sm_debug_decode_logits(sm_handle* h, int lane, float* logits_out, void* stream) {
    if (!h || !h->ds_logits || lane < 0 || lane >= kDsMaxStreams || !logits_out) return fail(h, "sm_debug_decode_logits: bad argument");
    cudaSetDevice(h->device);
    CUDA_OK(h, cudaMemcpyAsync(logits_out, h->ds_logits + static_cast<size_t>(lane) * h->cfg.llm_vocab,
                               static_cast<size_t>(h->cfg.llm_vocab) * sizeof(float), cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    return 0;
}

int sm_kv_len(const sm_handle* h) { return (h && !h->kv_lens.empty()) ? h->kv_lens[h->cur] : -1; }

int sm_kv_set_len(sm_handle* h, int len) {
    if (!h || h->kv_lens.empty()) return 1;
    if (len < 0 || len > h->kv_lens[h->cur]) return fail(h, "sm_kv_set_len: %d outside [0, %d]", len, h->kv_lens[h->cur]);
    h->kv_lens[h->cur] = len;      // host-side bookkeeping only: positions >= len are simply overwritten by the next prefill / decode
    return 0;
}

int sm_test_gemm(sm_handle* h, const void* x, const void* w, const void* bias, void* out, int M, int N, int K, int epi,
                 int force_swap, int force_bn, void* stream) {
    if (!h) return 1;
    cudaSetDevice(h->device);
    return launch_gemm(h, x, M, w, N, K, bias, out, N, epi, static_cast<cudaStream_t>(stream), force_swap, force_bn);
}

int sm_test_attention(sm_handle* h, const void* qkv, void* out, int B, int S, int H, int D, void* stream) {
    if (!h) return 1;
    cudaSetDevice(h->device);
    const int C = H * D;
    AttnArgs a{};
    a.q = qkv;
    a.k = reinterpret_cast<const char*>(qkv) + static_cast<size_t>(C) * 2;
    a.v = reinterpret_cast<const char*>(qkv) + static_cast<size_t>(2 * C) * 2;
    a.o = out;
    a.q_bs = a.k_bs = a.v_bs = static_cast<long long>(S) * 3 * C;
    a.q_ss = a.k_ss = a.v_ss = 3 * C;
    a.k_hs = a.v_hs = D;
    a.o_bs = static_cast<long long>(S) * C;
    a.o_ss = C;
    a.q_len = S; a.kv_len = S; a.q_pos0 = 0; a.causal = 0; a.group = 1;
    a.scale_log2e = static_cast<float>(1.4426950408889634 / std::sqrt(static_cast<double>(D)));
    return launch_attn(h, a, D, H, B, static_cast<cudaStream_t>(stream));
}

int sm_test_kv_attention(sm_handle* h, const void* q, int q_pitch, const void* kcache, const void* vcache, int max_ctx, void* out, int P,
                         int pos0, int Hq, int Hk, int n_splits, void* stream) {
    if (!h) return 1;
    cudaSetDevice(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int D = 128;
    const float scale = static_cast<float>(1.4426950408889634 / std::sqrt(static_cast<double>(D)));
    if (n_splits >= 0) {
        if (Hk < 1 || Hq % Hk != 0 || 128 % (Hq / Hk) != 0 || (128 / (Hq / Hk)) % 8 != 0) return fail(h, "sm_test_kv_attention: unsupported head grouping %d / %d", Hq, Hk);
        if (h->lw_akv_o == nullptr) {        // a handle without an LLM (unit tests): partial buffers on first use
            const int TB = 128 / (Hq / Hk);
            h->lw_akv_rows = std::max(512, h->num_sms * TB / Hk + TB);
            h->lw_akv_o = static_cast<float*>(dalloc(h, static_cast<size_t>(h->lw_akv_rows) * Hq * 128 * sizeof(float)));
            h->lw_akv_ml = static_cast<float*>(dalloc(h, static_cast<size_t>(h->lw_akv_rows) * Hq * 2 * sizeof(float)));
            if (!h->lw_akv_o || !h->lw_akv_ml) return fail(h, "sm_test_kv_attention: out of device memory");
        }
        DISPATCH_T(h, T, return launch_attn_kv_tc_t<T>(h, q, P, q_pitch, 0, kcache, vcache, max_ctx, out, Hq * D, P, pos0, Hq, Hk, scale, n_splits, st);)
    }
    AttnArgs a{};
    a.q = q; a.o = out; a.k = kcache; a.v = vcache;
    a.q_bs = 0; a.q_ss = q_pitch;
    a.k_bs = a.v_bs = 0; a.k_hs = a.v_hs = static_cast<long long>(max_ctx) * D; a.k_ss = a.v_ss = D;
    a.o_bs = 0; a.o_ss = Hq * D;
    a.q_len = P; a.kv_len = pos0 + P; a.q_pos0 = pos0; a.causal = 1; a.group = Hq / Hk;
    a.scale_log2e = scale;
    return launch_attn(h, a, D, Hq, 1, st);
}

int sm_debug_attention_mode(sm_handle* h, int mode) {
    if (!h) return 1;
    h->attn_mode = mode;
    return 0;
}

int sm_test_gemm_trace(sm_handle* h, long long* device_buf) {
    if (!h) return 1;
    h->gemm_dbg = device_buf;
    return 0;
}

int sm_debug_kernel_filter(sm_handle* h, unsigned mask) {
    if (!h) return 1;
    h->kfilter = mask;
    return 0;
}

int sm_profile_enable(sm_handle* h, int on) {
    if (!h) return 1;
    h->profiling = on != 0;
    return 0;
}

int sm_profile_read(sm_handle* h, int max_classes, double* ms_by_class, long long* launches_by_class) {
    if (!h) return -1;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < max_classes; ++i) { ms_by_class[i] = 0.0; launches_by_class[i] = 0; }
    for (auto& r : h->prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess && r.cls < max_classes) {
            ms_by_class[r.cls] += ms;
            launches_by_class[r.cls] += 1;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    h->prof.clear();
    return KC_COUNT;
}

const char* sm_profile_class_name(int cls) { return (cls >= 0 && cls < KC_COUNT) ? kKClassNames[cls] : ""; }

long long sm_launch_count(sm_handle* h, int reset) {
    if (!h) return 0;
    const long long v = h->launches;
    if (reset) h->launches = 0;
    return v;
}

}  // extern "C"
