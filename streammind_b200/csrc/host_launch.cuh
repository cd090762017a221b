// Host side of the library, part 2: TMA tensor maps, GEMM tile planner and launcher, GEMV and attention launchers, kernel attributes.
// Fragment of the library's single translation unit: included by api.cu, in this order, inside nothing (it opens its own
// anonymous namespace where it needs one).
#pragma once

namespace {

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


}  // namespace
