// Small bandwidth kernels around the tensor-core tiles: patch gather (im2col), embedding assembly +
// LayerNorms, patch mean-pool, RoPE + KV-cache append, split-KV decode attention, argmax, row gather.
#pragma once
#include "ptx.cuh"

namespace smb {

// ------------------------------------------------------------------------------------------
// ViT patch gather: pixels [B,3,H,W] (NCHW) -> A [B*P, Kpad], column k = c*p*p + i*p + j (the order of
// Conv2d weight [C_out, 3, p, p] flattened), zero padded to Kpad (TMA needs 16-byte row pitch).
// hf CLIPVisionEmbeddings.patch_embedding (modeling_clip.py:148-154,209-210)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void im2col_kernel(const T* __restrict__ px, T* __restrict__ out, int B, int img, int patch, int kpad) {
    pdl_trigger();   // programmatic dependent launch: the next kernel may start its prologue now ...
    pdl_wait();      // ... and this one touches its predecessor's outputs only from here on
    // work item = (output row, segment): segments 0 .. 3*patch-1 are one patch line each (patch contiguous pixels ->
    // patch contiguous columns), the last segment zero-fills the K padding
    const int gw = img / patch, P = gw * gw, kreal = 3 * patch * patch, nseg = 3 * patch + 1;
    const long long total = static_cast<long long>(B) * P * nseg;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int seg = static_cast<int>(i % nseg);
        const long long row = i / nseg;
        T* o = out + row * kpad;
        if (seg == 3 * patch) {
            for (int k = kreal; k < kpad; ++k) o[k] = Cvt<T>::from_f(0.f);
        } else {
            const int b = static_cast<int>(row / P), p = static_cast<int>(row % P);
            const int py = p / gw, pxx = p % gw;
            const int c = seg / patch, ii = seg % patch;
            const T* src = px + ((static_cast<long long>(b) * 3 + c) * img + (py * patch + ii)) * img + pxx * patch;
            T* d = o + c * patch * patch + ii * patch;
            for (int jj = 0; jj < patch; ++jj) d[jj] = src[jj];
        }
    }
}

// Re-tile a row-major weight block [rows, cols] into the GEMM's HBM operand layout
// dst[((r / 128) * KB + c / 64) * 128 + r % 128][c % 64]  (r counted from row0 of the packed matrix).
template <typename T>
__global__ void retile_weight_kernel(const T* __restrict__ src, T* __restrict__ dst, int rows, int cols, int row0,
                                     int kb_total) {
    const long long total = static_cast<long long>(rows) * cols;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / cols) + row0, c = static_cast<int>(i % cols);
        const long long tile = static_cast<long long>(r / 128) * kb_total + c / 64;
        dst[(tile * 128 + r % 128) * 64 + c % 64] = src[i];
    }
}

// warp-per-row LayerNorm helper: C <= 32*VPL elements, fp32 statistics, returns normalised values in v[]
template <typename T, int VPL>
__device__ __forceinline__ void warp_layernorm(float (&v)[VPL], int C, int lane, const T* w, const T* b, float eps) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) s += v[i];
    s = warp_sum(s);
    const float mean = s / C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = lane + i * 32;
        const float d = c < C ? v[i] - mean : 0.f;
        sq += d * d;
    }
    sq = warp_sum(sq);
    const float r = rsqrtf(sq / C + eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = lane + i * 32;
        if (c < C) v[i] = rnd<T>((v[i] - mean) * r * Cvt<T>::to_f(w[c]) + Cvt<T>::to_f(b[c]));
    }
}

// x[row] = pre_layrnorm( T( (p==0 ? cls : patch_emb[f*P + p-1]) + pos[p] ) );  h[row] = LN1_layer0(x[row])
// (modeling_clip.py:212-218 cat + position add, :677 pre_layrnorm, :363-366 layer_norm1)
template <typename T, int VPL>
__global__ void vit_embed_ln_kernel(const T* __restrict__ patch_emb, const T* __restrict__ cls,
                                    const T* __restrict__ pos, const T* pre_w, const T* pre_b, const T* ln_w,
                                    const T* ln_b, T* __restrict__ x, T* __restrict__ h, int rows, int S, int C,
                                    float eps) {
    pdl_trigger();   // programmatic dependent launch: the next kernel may start its prologue now ...
    pdl_wait();      // ... and this one touches its predecessor's outputs only from here on
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const int f = warp / S, p = warp % S;
    float v[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = lane + i * 32;
        if (c < C) {
            const float e = p == 0 ? Cvt<T>::to_f(cls[c])
                                   : Cvt<T>::to_f(patch_emb[(static_cast<long long>(f) * (S - 1) + p - 1) * C + c]);
            v[i] = rnd<T>(e + Cvt<T>::to_f(pos[static_cast<long long>(p) * C + c]));
        } else {
            v[i] = 0.f;
        }
    }
    warp_layernorm<T, VPL>(v, C, lane, pre_w, pre_b, eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = lane + i * 32;
        if (c < C) x[static_cast<long long>(warp) * C + c] = Cvt<T>::from_f(v[i]);
    }
    warp_layernorm<T, VPL>(v, C, lane, ln_w, ln_b, eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = lane + i * 32;
        if (c < C) h[static_cast<long long>(warp) * C + c] = Cvt<T>::from_f(v[i]);
    }
}

// LayerNorm rows: one warp per row, fp32 two-pass statistics held in registers.  Fast path: C a multiple of
// 256 up to 1024 -> each lane owns C/256 chunks of 8 contiguous elements (16-byte loads / stores).
template <typename T>
__global__ void layernorm_kernel(const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ b,
                                 T* __restrict__ h, int rows, int C, float eps) {
    pdl_trigger();   // programmatic dependent launch: the next kernel may start its prologue now ...
    pdl_wait();      // ... and this one touches its predecessor's outputs only from here on
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const T* xr = x + static_cast<long long>(warp) * C;
    T* hr = h + static_cast<long long>(warp) * C;
    if ((C & 255) == 0 && C <= 1024) {
        float v[4][8];
        const int nch = C >> 8;
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < nch) {
                const uint4 u = *reinterpret_cast<const uint4*>(xr + j * 256 + lane * 8);
                const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = Cvt<T>::unpack2(uw[i]);
                    v[j][2 * i] = f.x; v[j][2 * i + 1] = f.y;
                    s += f.x + f.y;
                }
            }
        }
        s = warp_sum(s);
        const float mean = s / C;
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < nch)
#pragma unroll
                for (int i = 0; i < 8; ++i) { const float d = v[j][i] - mean; sq += d * d; }
        sq = warp_sum(sq);
        const float r = rsqrtf(sq / C + eps);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < nch) {
                const uint4 wu = *reinterpret_cast<const uint4*>(w + j * 256 + lane * 8);
                const uint4 bu = *reinterpret_cast<const uint4*>(b + j * 256 + lane * 8);
                const uint32_t ww[4] = {wu.x, wu.y, wu.z, wu.w}, bw[4] = {bu.x, bu.y, bu.z, bu.w};
                uint32_t o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 wf = Cvt<T>::unpack2(ww[i]), bf = Cvt<T>::unpack2(bw[i]);
                    o[i] = Cvt<T>::pack2((v[j][2 * i] - mean) * r * wf.x + bf.x, (v[j][2 * i + 1] - mean) * r * wf.y + bf.y);
                }
                *reinterpret_cast<uint4*>(hr + j * 256 + lane * 8) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
    } else {
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += Cvt<T>::to_f(xr[c]);
        s = warp_sum(s);
        const float mean = s / C;
        float sq = 0.f;
        for (int c = lane; c < C; c += 32) { const float d = Cvt<T>::to_f(xr[c]) - mean; sq += d * d; }
        sq = warp_sum(sq);
        const float r = rsqrtf(sq / C + eps);
        for (int c = lane; c < C; c += 32)
            hr[c] = Cvt<T>::from_f((Cvt<T>::to_f(xr[c]) - mean) * r * Cvt<T>::to_f(w[c]) + Cvt<T>::to_f(b[c]));
    }
}

// Consumer of a split-K residual GEMM, fused with the LayerNorm that follows it:
//   v = T(sum_z part[z][row][:] + bias)   (fixed order -> deterministic);  x[row] = T(x[row] + v);
//   h[row] = LN(x[row]) * w + b           (skipped when w == nullptr: last layer)
// One CTA per row, one thread per 8 contiguous elements (blockDim = C/8 <= 128): every thread issues all
// of its partial-sum loads at once, so the kernel costs about one L2 round trip plus two block reductions.
template <typename T>
__global__ void __launch_bounds__(128) splitk_residual_ln_kernel(
    const float* __restrict__ part, int nsplit, long long split_stride, const T* __restrict__ bias, T* __restrict__ x,
    const T* __restrict__ w, const T* __restrict__ b, T* __restrict__ h, int rows, int C, float eps) {
    pdl_trigger();   // programmatic dependent launch: the next kernel may start its prologue now ...
    pdl_wait();      // ... and this one touches its predecessor's outputs only from here on
    __shared__ float red[8];
    const int row = blockIdx.x, c0 = threadIdx.x * 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    T* xr = x + static_cast<long long>(row) * C;
    const float* p = part + static_cast<long long>(row) * C + c0;
    float4 pa[4], pb[4];
#pragma unroll
    for (int z = 0; z < 4; ++z) {
        if (z < nsplit) {
            pa[z] = *reinterpret_cast<const float4*>(p + z * split_stride);
            pb[z] = *reinterpret_cast<const float4*>(p + z * split_stride + 4);
        }
    }
    const uint4 bu = *reinterpret_cast<const uint4*>(bias + c0);
    const uint4 xu = *reinterpret_cast<const uint4*>(xr + c0);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int z = 0; z < 4; ++z) {
        if (z < nsplit) {
            acc[0] += pa[z].x; acc[1] += pa[z].y; acc[2] += pa[z].z; acc[3] += pa[z].w;
            acc[4] += pb[z].x; acc[5] += pb[z].y; acc[6] += pb[z].z; acc[7] += pb[z].w;
        }
    }
    const uint32_t bw[4] = {bu.x, bu.y, bu.z, bu.w}, xw[4] = {xu.x, xu.y, xu.z, xu.w};
    float v[8];
    uint32_t o[4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 bf = Cvt<T>::unpack2(bw[i]), xf = Cvt<T>::unpack2(xw[i]);
        v[2 * i] = rnd<T>(xf.x + rnd<T>(acc[2 * i] + bf.x));
        v[2 * i + 1] = rnd<T>(xf.y + rnd<T>(acc[2 * i + 1] + bf.y));
        s += v[2 * i] + v[2 * i + 1];
        o[i] = Cvt<T>::pack2(v[2 * i], v[2 * i + 1]);
    }
    *reinterpret_cast<uint4*>(xr + c0) = make_uint4(o[0], o[1], o[2], o[3]);
    if (w == nullptr) return;
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < nwarp; ++i) tot += red[i];
    const float mean = tot / C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; sq += d * d; }
    sq = warp_sum(sq);
    if (lane == 0) red[4 + warp] = sq;
    __syncthreads();
    float var = 0.f;
    for (int i = 0; i < nwarp; ++i) var += red[4 + i];
    const float r = rsqrtf(var / C + eps);
    const uint4 wu = *reinterpret_cast<const uint4*>(w + c0);
    const uint4 b2 = *reinterpret_cast<const uint4*>(b + c0);
    const uint32_t ww[4] = {wu.x, wu.y, wu.z, wu.w}, b2w[4] = {b2.x, b2.y, b2.z, b2.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 wf = Cvt<T>::unpack2(ww[i]), bf = Cvt<T>::unpack2(b2w[i]);
        o[i] = Cvt<T>::pack2((v[2 * i] - mean) * r * wf.x + bf.x, (v[2 * i + 1] - mean) * r * wf.y + bf.y);
    }
    *reinterpret_cast<uint4*>(h + static_cast<long long>(row) * C + c0) = make_uint4(o[0], o[1], o[2], o[3]);
}

// feature_select('patch') + mean over patches: feats[f, p, :] = x[f*S + 1 + p, :] (CLS dropped,
// clip_encoder.py:31-35); pooled[f, :] = T(mean_p feats[f, p, :]) (multimodal_projector/builder.py:405)
template <typename T>
__global__ void __launch_bounds__(128) vit_finalize_kernel(const T* __restrict__ x, T* __restrict__ feats,
                                                           T* __restrict__ pooled, int S, int C) {
    pdl_trigger();   // programmatic dependent launch: the next kernel may start its prologue now ...
    pdl_wait();      // ... and this one touches its predecessor's outputs only from here on
    // one warp per (frame, 8-column chunk): lanes stride over the patches with 16-byte loads, fp32 partial sums,
    // one butterfly reduction per chunk (the old one-thread-per-column loop was a 576-step serial chain: 30 us)
    const int f = blockIdx.y;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int c0 = warp * 8;
    if (c0 >= C) return;
    const int P = S - 1;
    const T* src = x + (static_cast<long long>(f) * S + 1) * C + c0;
    T* dst = feats ? feats + static_cast<long long>(f) * P * C + c0 : nullptr;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int p = lane; p < P; p += 32) {
        const uint4 u = *reinterpret_cast<const uint4*>(src + static_cast<long long>(p) * C);
        if (dst) *reinterpret_cast<uint4*>(dst + static_cast<long long>(p) * C) = u;
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 fv = Cvt<T>::unpack2(w[i]);
            acc[2 * i] += fv.x; acc[2 * i + 1] += fv.y;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = warp_sum(acc[i]);
    if (pooled && lane == 0) {
        uint4 o;
        o.x = Cvt<T>::pack2(acc[0] / P, acc[1] / P); o.y = Cvt<T>::pack2(acc[2] / P, acc[3] / P);
        o.z = Cvt<T>::pack2(acc[4] / P, acc[5] / P); o.w = Cvt<T>::pack2(acc[6] / P, acc[7] / P);
        *reinterpret_cast<uint4*>(pooled + static_cast<long long>(f) * C + c0) = o;
    }
}

// pooled[f, :] = T(mean_p feats[f, p, :]) for externally supplied features (B2 hook: mm_projector(feats))
template <typename T>
__global__ void pool_kernel(const T* __restrict__ feats, T* __restrict__ pooled, int P, int C) {
    const int f = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    const T* src = feats + static_cast<long long>(f) * P * C + c;
    for (int p = 0; p < P; ++p) s += Cvt<T>::to_f(src[static_cast<long long>(p) * C]);
    pooled[static_cast<long long>(f) * C + c] = Cvt<T>::from_f(s / P);
}

// ------------------------------------------------------------------------------------------
// LLM: RoPE (rotate-half, fp32 cos/sin, hf modeling_mistral.py:51-81) on q in place and on k while
// appending k, v to the cache.  qkv rows: [q (Hq*D) | k (Hk*D) | v (Hk*D)].
// cache layout: K/V [Hk][max_ctx][D] per layer.  positions pos0 + row.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void rope_append_kernel(T* __restrict__ qkv, T* __restrict__ kc, T* __restrict__ vc, int rows, int Hq,
                                   int Hk, int D, int max_ctx, const int* pos0_ptr, int pos0_host, float theta) {
    pdl_trigger();   // programmatic dependent launch: the next kernel may start its prologue now ...
    pdl_wait();      // ... and this one touches its predecessor's outputs only from here on
    const int pos0 = pos0_ptr ? *pos0_ptr : pos0_host;
    const int half = D / 2;
    const int per_row = (Hq + Hk) * half + Hk * D;
    const long long total = static_cast<long long>(rows) * per_row;
    const int ld = (Hq + 2 * Hk) * D;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int row = static_cast<int>(i / per_row);
        int j = static_cast<int>(i % per_row);
        const int pos = pos0 + row;
        T* base = qkv + static_cast<long long>(row) * ld;
        if (j < (Hq + Hk) * half) {
            const int hh = j / half, d = j % half;
            const float inv = powf(theta, -2.0f * d / D);
            float sn, cs;
            sincosf(pos * inv, &sn, &cs);
            // HF casts cos/sin to the model dtype before the multiply (modeling_mistral.py apply_rotary_pos_emb)
            cs = rnd<T>(cs);
            sn = rnd<T>(sn);
            T* p = base + hh * D;  // q heads then k heads are contiguous
            const float x1 = Cvt<T>::to_f(p[d]), x2 = Cvt<T>::to_f(p[d + half]);
            const float y1 = rnd<T>(x1 * cs) + rnd<T>(-x2 * sn);
            const float y2 = rnd<T>(x2 * cs) + rnd<T>(x1 * sn);
            if (hh < Hq) {
                p[d] = Cvt<T>::from_f(y1);
                p[d + half] = Cvt<T>::from_f(y2);
            } else {
                T* dst = kc + (static_cast<long long>(hh - Hq) * max_ctx + pos) * D;
                dst[d] = Cvt<T>::from_f(y1);
                dst[d + half] = Cvt<T>::from_f(y2);
            }
        } else {
            j -= (Hq + Hk) * half;
            const int hh = j / D, d = j % D;
            vc[(static_cast<long long>(hh) * max_ctx + pos) * D + d] = base[(Hq + Hk) * D + hh * D + d];
        }
    }
}

// rows[i, :] = table[ids[i], :]   (embed_tokens); ids may come from the device (decode loop)
template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ table, const int* __restrict__ ids, T* __restrict__ out,
                                   int n, int C) {
    pdl_trigger();   // programmatic dependent launch: the next kernel may start its prologue now ...
    pdl_wait();      // ... and this one touches its predecessor's outputs only from here on
    const int r = blockIdx.x;
    if (r >= n) return;
    const T* src = table + static_cast<long long>(ids[r]) * C;
    for (int c = threadIdx.x * 8; c < C; c += blockDim.x * 8)
        *reinterpret_cast<uint4*>(out + static_cast<long long>(r) * C + c) = *reinterpret_cast<const uint4*>(src + c);
}

// RMSNorm rows (prefill / batched gate): h = nw * T(x * rsqrt(mean(x^2) + eps)).  One CTA of 128 threads per row: every thread's pieces of the
// row AND of the weight are requested up front (one memory round trip for the whole kernel; a warp walking the row in a loop paid one
// per 512 bytes: 12 us per launch for 11 rows of 4096), the row is normalised from registers.  Rows wider than 8192 fall back to a loop.
constexpr int kRmsRowsThreads = 128;
template <typename T>
// nsplit > 0: the row is still waiting for the split-K partials of the projection that feeds the residual stream -- part[z][row][c] fp32,
// z < nsplit, split_stride apart -- and this kernel folds them first (x = T(x + T(sum_z part)), fixed order, written back to x_io),
// instead of a reduce kernel of its own between the GEMM and the norm.
__global__ void __launch_bounds__(kRmsRowsThreads) rmsnorm_rows_kernel(const T* x, const T* __restrict__ nw, T* __restrict__ h, int rows,
                                                                       int C, float eps, const float* __restrict__ part, int nsplit,
                                                                       long long split_stride, T* x_io) {
    constexpr int MAXI = 8;                       // 16-byte pieces per thread held in registers
    __shared__ float red[kRmsRowsThreads / 32];
    pdl_trigger();
    pdl_wait();
    const int row = blockIdx.x, tid = threadIdx.x;
    if (row >= rows) return;
    const T* xr = x + static_cast<long long>(row) * C;
    T* hr = h + static_cast<long long>(row) * C;
    const bool fits = C <= MAXI * kRmsRowsThreads * 8;
    uint4 u[MAXI], g[MAXI];
    float s = 0.f;
    if (fits) {
#pragma unroll
        for (int i = 0; i < MAXI; ++i) {
            const int c = (tid + i * kRmsRowsThreads) * 8;
            if (c < C) { u[i] = *reinterpret_cast<const uint4*>(xr + c); g[i] = *reinterpret_cast<const uint4*>(nw + c); }
        }
        if (nsplit > 0) {
#pragma unroll
            for (int i = 0; i < MAXI; ++i) {
                const int c = (tid + i * kRmsRowsThreads) * 8;
                if (c < C) {
                    const float* pr = part + static_cast<long long>(row) * C + c;
                    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
                    for (int z = 0; z < nsplit; ++z) {
                        const float4 p0 = *reinterpret_cast<const float4*>(pr + z * split_stride), p1 = *reinterpret_cast<const float4*>(pr + z * split_stride + 4);
                        a[0] += p0.x; a[1] += p0.y; a[2] += p0.z; a[3] += p0.w; a[4] += p1.x; a[5] += p1.y; a[6] += p1.z; a[7] += p1.w;
                    }
                    const uint32_t w[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
                    uint32_t o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = Cvt<T>::unpack2(w[j]);
                        o[j] = Cvt<T>::pack2(f.x + rnd<T>(a[2 * j]), f.y + rnd<T>(a[2 * j + 1]));
                    }
                    u[i] = make_uint4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<uint4*>(x_io + static_cast<long long>(row) * C + c) = u[i];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < MAXI; ++i) {
            if ((tid + i * kRmsRowsThreads) * 8 < C) {
                const uint32_t w[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float2 f = Cvt<T>::unpack2(w[j]); s += f.x * f.x + f.y * f.y; }
            }
        }
    } else {
        for (int c = tid * 8; c < C; c += kRmsRowsThreads * 8) {
            const uint4 v = *reinterpret_cast<const uint4*>(xr + c);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float2 f = Cvt<T>::unpack2(w[j]); s += f.x * f.x + f.y * f.y; }
        }
    }
    s = warp_sum(s);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kRmsRowsThreads / 32; ++w) tot += red[w];      // fixed order
    const float r = rsqrtf(tot / C + eps);
    auto norm8 = [&](const uint4& xv, const uint4& gv) {
        const uint32_t w[4] = {xv.x, xv.y, xv.z, xv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = Cvt<T>::unpack2(w[j]), gg = Cvt<T>::unpack2(gw[j]);
            o[j] = Cvt<T>::pack2(gg.x * rnd<T>(f.x * r), gg.y * rnd<T>(f.y * r));
        }
        return make_uint4(o[0], o[1], o[2], o[3]);
    };
    if (fits) {
#pragma unroll
        for (int i = 0; i < MAXI; ++i) {
            const int c = (tid + i * kRmsRowsThreads) * 8;
            if (c < C) *reinterpret_cast<uint4*>(hr + c) = norm8(u[i], g[i]);
        }
    } else {
        for (int c = tid * 8; c < C; c += kRmsRowsThreads * 8)
            *reinterpret_cast<uint4*>(hr + c) = norm8(*reinterpret_cast<const uint4*>(xr + c), *reinterpret_cast<const uint4*>(nw + c));
    }
}

// SwiGLU rows (prefill path): m[r, j] = T(T(silu(gu[r, j])) * gu[r, F + j])
template <typename T>
__global__ void swiglu_rows_kernel(const T* __restrict__ gu, T* __restrict__ m, int rows, int F) {
    pdl_trigger();
    pdl_wait();
    const long long total = static_cast<long long>(rows) * F;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / F;
        const int j = static_cast<int>(i % F);
        const float g = Cvt<T>::to_f(gu[r * 2 * F + j]), u = Cvt<T>::to_f(gu[r * 2 * F + F + j]);
        const float sg = g / (1.0f + __expf(-g));
        m[i] = Cvt<T>::from_f(rnd<T>(sg) * u);
    }
}

// Consumer of a split-K GEMM on a few rows (batched gate): v = T(sum_z part[z][i]) in fixed order;
// out[i] = resid ? T(resid[i] + v) : v      (resid may alias out: in-place residual stream)
template <typename T>
__global__ void splitk_rows_kernel(const float* __restrict__ part, int nsplit, long long split_stride, const T* resid, T* out,
                                   long long total) {
    pdl_trigger();
    pdl_wait();
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float acc = 0.f;
#pragma unroll 4      // the partials are independent loads: in flight together, summed in fixed order
        for (int z = 0; z < nsplit; ++z) acc += part[z * split_stride + i];
        const float v = rnd<T>(acc);
        out[i] = Cvt<T>::from_f(resid != nullptr ? Cvt<T>::to_f(resid[i]) + v : v);
    }
}

// repeat_kv at L = 1 for a batch of rows: out[r, h*D + d] = v[r, (h / rep)*D + d]   (gate batched over frames)
template <typename T>
__global__ void gqa_expand_rows_kernel(const T* __restrict__ v, T* __restrict__ out, int rows, int Hq, int Hk, int D) {
    pdl_trigger();
    pdl_wait();
    const int rep = Hq / Hk, per_row = Hq * D / 8;
    const long long total = static_cast<long long>(rows) * per_row;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / per_row;
        const int c = static_cast<int>(i % per_row) * 8, hq = c / D, d = c % D;
        *reinterpret_cast<uint4*>(out + r * Hq * D + c) = *reinterpret_cast<const uint4*>(v + r * Hk * D + (hq / rep) * D + d);
    }
}

// ------------------------------------------------------------------------------------------
// Cognition sampling before the LLM (/root/reference/streammind/model/videollama2_arch.py:595-611):
// similarity_sampling keeps the top-k frame tokens by cosine similarity with the LAST token, in their original order.
//   sim_i = T(sum_d T(T(x_id / n_i) * T(x_Ld / n_L))),  n_i = max(T(sqrt(sum_d x_id^2)), eps)      (torch's
//   cosine_similarity: normalise, multiply, sum, every tensor materialised in T); sums in fp64, so the result does not
//   depend on the summation order and equals the oracle bit for bit.  One warp per row.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void cos_sim_rows_kernel(const T* __restrict__ x, int n, int d, float eps, float* __restrict__ sim) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const T* xi = x + static_cast<long long>(warp) * d;
    const T* xl = x + static_cast<long long>(n - 1) * d;
    double si = 0.0, sl = 0.0;
    for (int k = lane; k < d; k += 32) {
        const double a = Cvt<T>::to_f(xi[k]), b = Cvt<T>::to_f(xl[k]);
        si += a * a;
        sl += b * b;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { si += __shfl_xor_sync(0xffffffffu, si, o); sl += __shfl_xor_sync(0xffffffffu, sl, o); }
    const float ni = fmaxf(rnd<T>(static_cast<float>(sqrt(si))), eps), nl = fmaxf(rnd<T>(static_cast<float>(sqrt(sl))), eps);
    double acc = 0.0;
    for (int k = lane; k < d; k += 32) {
        const float a = rnd<T>(Cvt<T>::to_f(xi[k]) / ni), b = rnd<T>(Cvt<T>::to_f(xl[k]) / nl);
        acc += static_cast<double>(rnd<T>(a * b));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) sim[warp] = rnd<T>(static_cast<float>(acc));
}

// Indices of the k largest sim values in ascending index order; ties go to the lower index (the reference's unstable argsort
// leaves tie order unspecified).  One CTA; rank by counting (n <= a few thousand frame tokens).
__global__ void __launch_bounds__(1024) topk_keep_order_kernel(const float* __restrict__ sim, int n, int k, int* __restrict__ idx_out) {
    extern __shared__ int keep[];                  // [n]
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float si = sim[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const float sj = sim[j];
            rank += (sj > si || (sj == si && j < i)) ? 1 : 0;
        }
        keep[i] = rank < k ? 1 : 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (!keep[i]) continue;
        int pos = 0;
        for (int j = 0; j < i; ++j) pos += keep[j];
        idx_out[pos] = i;
    }
}

}  // namespace smb
