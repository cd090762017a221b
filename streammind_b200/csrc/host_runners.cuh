// Host side of the library, part 3: the launch sequences of the sub-models (vision tower, projector, gate as GEMVs / GEMMs, LLM prefill chunk).
// Fragment of the library's single translation unit: included by api.cu, in this order, inside nothing (it opens its own
// anonymous namespace where it needs one).
#pragma once

namespace {

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
// deferred != nullptr: when K is split, the partials are left for the caller's next kernel to fold (rmsnorm_rows_kernel does it for the
// residual stream) and *deferred = the number of splits; 0 when the GEMM wrote `out` itself.
int gemm_few_rows(sm_handle* h, const void* x, int n, const void* W, int N, int K, void* out, bool residual, float* part,
                  cudaStream_t st, int* deferred = nullptr) {
    if (deferred) *deferred = 0;
    const int tiles = (N + 127) / 128, kb = (K + 63) / 64;
    const int bn = std::max(16, (n + 15) / 16 * 16);
    int split = std::min({8, std::max(1, h->num_sms / tiles), std::max(1, kb / 8)});
    while (split > 1 && (split - 1) * ((kb + split - 1) / split) >= kb) --split;
    if (split < 2 || part == nullptr)
        return launch_gemm(h, x, n, W, N, K, nullptr, out, N, residual ? EPI_RESIDUAL : EPI_STORE, st, 1, bn);
    if (launch_gemm(h, x, n, W, N, K, nullptr, part, N, EPI_STORE_F32, st, 1, bn, false, split)) return 1;
    if (deferred) { *deferred = split; return 0; }
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
    auto rms = [&](const void* nw) -> int {
        DISPATCH_T(h, T, {
            CUDA_OK(h, launch_pdl(h, rmsnorm_rows_kernel<T>, dim3(n), dim3(kRmsRowsThreads), 0, st, (const T*)h->gg_h, (const T*)nw, (T*)h->gg_hn, n, H, c.gate_eps, (const float*)nullptr, 0, 0LL, (T*)nullptr));
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
    // short dialogue suffixes (a fire prefills 11-74 new tokens): the narrow projections are split along K
    static const bool few_on = getenv("SMB_PREFILL_SPLITK") ? atoi(getenv("SMB_PREFILL_SPLITK")) != 0 : true;
    const bool few = few_on && P <= 64 && h->lw_part2 != nullptr;
    // split-K partials of a residual projection (o_proj, down_proj) are folded into the residual stream by the RMSNorm kernel that
    // reads it next, not by a kernel of their own; `pending` = splits of the partials in lw_part2 that lw_x is still waiting for
    int pending = 0;
    const long long part_stride = static_cast<long long>(P) * H;
    int* const defer = H <= 8 * kRmsRowsThreads * 8 ? &pending : nullptr;      // the norm kernel folds partials for rows it holds in registers
    for (int l = 0; l < c.llm_layers; ++l) {
        const MistralLayer& L = h->llm[l];
        DISPATCH_T(h, T, {
            ProfScope ps_kc_rmsnorm_rows(h, KC_RMSNORM_ROWS, st);
            if (kon(h, KC_RMSNORM_ROWS)) {
            rmsnorm_rows_kernel<T><<<P, kRmsRowsThreads, 0, st>>>((const T*)h->lw_x, (const T*)L.in_ln, (T*)h->lw_hn, P, H, c.llm_eps,
                                                                 (const float*)h->lw_part2, pending, part_stride, (T*)h->lw_x);
            }
            count_launch(h);
        })
        pending = 0;
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
        if (few) { if (gemm_few_rows(h, h->lw_att, P, L.wo, H, Hq * D, h->lw_x, true, h->lw_part2, st, defer)) return 1; }
        else if (launch_gemm(h, h->lw_att, P, L.wo, H, Hq * D, nullptr, h->lw_x, H, EPI_RESIDUAL, st)) return 1;
        DISPATCH_T(h, T, {
            ProfScope ps_kc_rmsnorm_rows(h, KC_RMSNORM_ROWS, st);
            if (kon(h, KC_RMSNORM_ROWS)) {
            rmsnorm_rows_kernel<T><<<P, kRmsRowsThreads, 0, st>>>((const T*)h->lw_x, (const T*)L.post_ln, (T*)h->lw_hn, P, H, c.llm_eps,
                                                                 (const float*)h->lw_part2, pending, part_stride, (T*)h->lw_x);
            }
            count_launch(h);
        })
        pending = 0;
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
        if (few) { if (gemm_few_rows(h, h->lw_m, P, L.wd, H, F, h->lw_x, true, h->lw_part2, st, defer)) return 1; }
        else if (launch_gemm(h, h->lw_m, P, L.wd, H, F, nullptr, h->lw_x, H, EPI_RESIDUAL, st)) return 1;
    }
    if (pending > 0) {        // the last down_proj: nothing normalises the residual stream after it inside this chunk
        DISPATCH_T(h, T, {
            CUDA_OK(h, launch_pdl(h, splitk_rows_kernel<T>, dim3(static_cast<int>(std::min<long long>((part_stride + 255) / 256, 1024))), dim3(256), 0, st,
                                  (const float*)h->lw_part2, pending, part_stride, (const T*)h->lw_x, (T*)h->lw_x, part_stride));
            count_launch(h);
        })
    }
    CUDA_OK(h, cudaGetLastError());
    return 0;
}



}  // namespace
