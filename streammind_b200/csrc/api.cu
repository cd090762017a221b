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

#include "host_state.cuh"      // slots, struct sm_handle, error / allocation / launch helpers
#include "host_launch.cuh"     // tensor maps, GEMM planner, kernel launchers
#include "host_runners.cuh"    // launch sequences of the sub-models
#include "host_decode.cuh"     // persistent decode step
#include "host_pipeline.cuh"   // graph capture, pipelined frame path


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

