// Weight-streaming GEMV for the batch-1 chains on the path: projector step (PreNet / Mamba-1 step /
// PostNet), the event gate (4 Mistral layers at L = 1) and LLM decode.  HBM-bound: each weight byte is
// read exactly once with 16-byte coalesced no-allocate loads; the input vector (<= 28 KB) is staged
// in shared memory after its producer-side transform (RMSNorm / LayerNorm / leaky-ReLU / GQA expand).
//
//   y[n] = epilogue( sum_k W[n, k] * pro(x)[k] )           W: [N, K] row-major, model dtype T
//
// Decomposition: the grid splits N into equal contiguous row blocks (one per CTA, grid a multiple of
// the SM count); inside a CTA the (row, K-segment) items are dealt round-robin to 16 warps; each lane
// streams 16-byte chunks, accumulates in fp32, warp-reduces, and the per-segment partials are summed
// in a FIXED order by the epilogue thread of that row (deterministic: greedy decode is reproducible).
// The first weight chunks are issued before the input vector is staged, so the prologue (and, with
// programmatic dependent launch, the previous kernel's tail) hides under the first HBM round trip.
#pragma once
#include "ptx.cuh"

namespace smb {

enum GemvPro : int {
    PRO_PLAIN = 0,        // x = x0
    PRO_RMSNORM = 1,      // x = nw * T(x0 * rsqrt(mean(x0^2) + eps))           (hf MistralRMSNorm)
    PRO_LAYERNORM = 2,    // x = T(LN(x0) * nw + nb)
    PRO_LN_LEAKY = 3,     // x = leaky_relu(T(LN(x0) * nw + nb))
    PRO_GQA_EXPAND = 4,   // x[h*D + d] = x0[(h / rep)*D + d]                      (repeat_kv at L = 1)
};
enum GemvEpi : int {
    GEPI_STORE = 0,       // y = T(acc + bias)
    GEPI_LEAKY = 1,       // y = leaky_relu(T(acc + bias))
    GEPI_RESID = 2,       // resid[n] = T(resid[n] + T(acc + bias))   (in place residual stream)
    GEPI_ADD_TO = 3,      // y[n] = T(resid[n] + T(acc + bias))        (out of place)
    GEPI_F32 = 4,         // y(float) = float(T(acc + bias))   (hf: lm_head output in T, then .float())
    GEPI_SWIGLU = 5,      // NMAT = 2: y = T(T(silu(T(acc0))) * T(acc1))        (hf MistralMLP)
    GEPI_MAMBA_CONV = 6,  // rows < d_inner: roll conv window, y = T(silu(T(conv + cb))); rows >= d_inner: z
};

struct GemvArgs {
    const void* W0;
    const void* W1;  // second matrix sharing the input (NMAT = 2), else unused
    int N, K;
    int seg_len;  // K-segment per work item (multiple of 256)
    int pro, epi;
    const void* x0;
    const void* nw;
    const void* nb;
    float eps;
    int gqa_rep, head_dim;
    const void* bias;
    void* y;
    void* resid;
    // GEPI_MAMBA_CONV
    void* conv_state;      // T [d_inner][d_conv], rolling window (oldest first)
    const void* conv_w;    // T [d_inner][d_conv]
    const void* conv_b;    // T [d_inner]
    void* z_out;           // T [d_inner]
    int d_inner, d_conv;
    // NV > 1 (template): the same weights against NV input vectors (consecutive frames of a chunk / of the frames in
    // flight): every weight byte is streamed once for NV frames.  Element strides between the vectors:
    long long x_stride, y_stride, resid_stride, z_stride;
    int nv_host;   // host-side copy of NV (selects the template instance)
    long long conv_state_stride;   // GEPI_MAMBA_CONV: elements between the conv windows of the NV vectors (0: consecutive frames of ONE
                                   // stream walk one window in order; > 0: the vectors are frames of different streams)
};

#ifndef SMB_GEMV_UNR1
#define SMB_GEMV_UNR1 4
#endif
constexpr int kGemvThreads = 512;
constexpr int kGemvWarps = kGemvThreads / 32;
constexpr int kGemvMaxRowsPerCta = 256;
constexpr int kGemvMaxSegs = 8;

__device__ __forceinline__ float leaky_relu_f(float v) { return v > 0.0f ? v : 0.01f * v; }
__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }

template <typename T>
__device__ __forceinline__ float dot8(const uint4& w, const T* xs) {
    // xs: 8 consecutive T in shared memory (16-byte aligned)
    const uint4 xv = *reinterpret_cast<const uint4*>(xs);
    float2 w0 = Cvt<T>::unpack2(w.x), w1 = Cvt<T>::unpack2(w.y), w2 = Cvt<T>::unpack2(w.z), w3 = Cvt<T>::unpack2(w.w);
    float2 x0 = Cvt<T>::unpack2(xv.x), x1 = Cvt<T>::unpack2(xv.y), x2 = Cvt<T>::unpack2(xv.z),
           x3 = Cvt<T>::unpack2(xv.w);
    float s = w0.x * x0.x;
    s = fmaf(w0.y, x0.y, s);
    s = fmaf(w1.x, x1.x, s);
    s = fmaf(w1.y, x1.y, s);
    s = fmaf(w2.x, x2.x, s);
    s = fmaf(w2.y, x2.y, s);
    s = fmaf(w3.x, x3.x, s);
    s = fmaf(w3.y, x3.y, s);
    return s;
}

// Stage pro(x_v) (GemvPro) of input vector v into xs (shared memory, T[K]); all kGemvThreads threads take part.
template <typename T>
__device__ __forceinline__ void gemv_stage_vector(const GemvArgs& a, int v, T* xs, float* red) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const T* x0 = reinterpret_cast<const T*>(a.x0) + v * a.x_stride;
    const T* nw = reinterpret_cast<const T*>(a.nw);
    const T* nb = reinterpret_cast<const T*>(a.nb);
    if (v > 0) __syncthreads();                             // red[] of the previous vector has been consumed
    if (a.pro == PRO_PLAIN) {
        for (int k = threadIdx.x * 8; k < a.K; k += kGemvThreads * 8)
            *reinterpret_cast<uint4*>(xs + k) = *reinterpret_cast<const uint4*>(x0 + k);
    } else if (a.pro == PRO_GQA_EXPAND) {
        for (int k = threadIdx.x; k < a.K; k += kGemvThreads) {
            const int h = k / a.head_dim, d = k - h * a.head_dim;
            xs[k] = x0[(h / a.gqa_rep) * a.head_dim + d];
        }
    } else {
        // norms: pass 1 statistics (fp32), pass 2 normalise.  K <= 16384 -> <= 32 elements / thread
        float s1 = 0.f, s2 = 0.f;
        for (int k = threadIdx.x; k < a.K; k += kGemvThreads) {
            const float v = Cvt<T>::to_f(x0[k]);
            s1 += v;
            s2 += v * v;
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0) { red[warp] = s1; red[kGemvWarps + warp] = s2; }
        __syncthreads();
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int w = 0; w < kGemvWarps; ++w) { t1 += red[w]; t2 += red[kGemvWarps + w]; }
        const float inv_k = 1.0f / static_cast<float>(a.K);
        if (a.pro == PRO_RMSNORM) {
            const float r = rsqrtf(t2 * inv_k + a.eps);
            for (int k = threadIdx.x; k < a.K; k += kGemvThreads) {
                const float v = rnd<T>(Cvt<T>::to_f(x0[k]) * r);
                xs[k] = Cvt<T>::from_f(Cvt<T>::to_f(nw[k]) * v);
            }
        } else {
            const float mean = t1 * inv_k;
            // two-pass variance for accuracy (x is L2/L1 resident)
            float sq = 0.f;
            for (int k = threadIdx.x; k < a.K; k += kGemvThreads) {
                const float d = Cvt<T>::to_f(x0[k]) - mean;
                sq += d * d;
            }
            sq = warp_sum(sq);
            __syncthreads();
            if (lane == 0) red[warp] = sq;
            __syncthreads();
            float var = 0.f;
#pragma unroll
            for (int w = 0; w < kGemvWarps; ++w) var += red[w];
            const float r = rsqrtf(var * inv_k + a.eps);
            for (int k = threadIdx.x; k < a.K; k += kGemvThreads) {
                float v = (Cvt<T>::to_f(x0[k]) - mean) * r * Cvt<T>::to_f(nw[k]) + Cvt<T>::to_f(nb[k]);
                v = rnd<T>(v);
                if (a.pro == PRO_LN_LEAKY) v = leaky_relu_f(v);
                xs[k] = Cvt<T>::from_f(v);
            }
        }
    }
}

// Epilogue (GemvEpi) of output row n for input vector v: acc / accb are the complete fp32 dot products.
template <typename T, int NMAT>
__device__ __forceinline__ void gemv_finish_row(const GemvArgs& a, int n, int v, float acc, float accb) {
    const T* bias = reinterpret_cast<const T*>(a.bias);
    if (bias != nullptr) acc += Cvt<T>::to_f(bias[n]);
    T* y = reinterpret_cast<T*>(a.y) + v * a.y_stride;
    T* resid = reinterpret_cast<T*>(a.resid) + v * a.resid_stride;
    switch (a.epi) {
        case GEPI_STORE: y[n] = Cvt<T>::from_f(acc); break;
        case GEPI_LEAKY: y[n] = Cvt<T>::from_f(leaky_relu_f(rnd<T>(acc))); break;
        case GEPI_RESID: resid[n] = Cvt<T>::from_f(Cvt<T>::to_f(resid[n]) + rnd<T>(acc)); break;
        case GEPI_ADD_TO: y[n] = Cvt<T>::from_f(Cvt<T>::to_f(resid[n]) + rnd<T>(acc)); break;
        case GEPI_F32: (reinterpret_cast<float*>(a.y) + v * a.y_stride)[n] = rnd<T>(acc); break;  // logits leave lm_head in T
        case GEPI_SWIGLU: {
            const float g = rnd<T>(silu_f(rnd<T>(acc)));
            y[n] = Cvt<T>::from_f(g * rnd<T>(accb));
            break;
        }
        case GEPI_MAMBA_CONV: {
            if (n < a.d_inner) {
                // Mamba causal depth-wise conv as a rolling window (mamba_simple.py:215-221); the
                // reference's full-sequence conv1d materialises T(conv + bias) before SiLU (:168-169)
                T* st = reinterpret_cast<T*>(a.conv_state) + v * a.conv_state_stride + static_cast<size_t>(n) * a.d_conv;
                const T* cw = reinterpret_cast<const T*>(a.conv_w) + static_cast<size_t>(n) * a.d_conv;
                const float xn = rnd<T>(acc);
                float c = 0.f;
                for (int w = 0; w < a.d_conv - 1; ++w) {
                    const T sv = st[w + 1];
                    st[w] = sv;
                    c = fmaf(Cvt<T>::to_f(sv), Cvt<T>::to_f(cw[w]), c);
                }
                st[a.d_conv - 1] = Cvt<T>::from_f(xn);
                c = fmaf(xn, Cvt<T>::to_f(cw[a.d_conv - 1]), c);
                c = rnd<T>(c + Cvt<T>::to_f(reinterpret_cast<const T*>(a.conv_b)[n]));
                y[n] = Cvt<T>::from_f(silu_f(c));
            } else {
                (reinterpret_cast<T*>(a.z_out) + v * a.z_stride)[n - a.d_inner] = Cvt<T>::from_f(acc);
            }
            break;
        }
    }
}

template <typename T, int NMAT, int NV = 1>
__global__ void __launch_bounds__(kGemvThreads, NV == 1 ? 2 : 1) gemv_kernel(const GemvArgs a) {
    extern __shared__ __align__(16) uint8_t gemv_smem[];
    const int xpitch = (a.K + 7) & ~7;                                          // elements per staged vector
    T* xs = reinterpret_cast<T*>(gemv_smem);                                   // [NV][xpitch]
    float* part = reinterpret_cast<float*>(gemv_smem + ((NV * xpitch * 2 + 15) & ~15));  // [NMAT][NV][rows][nseg]
    __shared__ float red[kGemvWarps * 2];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows_per_cta = (a.N + gridDim.x - 1) / gridDim.x;
    const int r0 = blockIdx.x * rows_per_cta;
    const int r1 = min(a.N, r0 + rows_per_cta);
    const int nrows = max(0, r1 - r0);
    const int nseg = (a.K + a.seg_len - 1) / a.seg_len;
    const int nitems = nrows * nseg;
    const T* W0 = reinterpret_cast<const T*>(a.W0);
    const T* W1 = reinterpret_cast<const T*>(a.W1);

    // ---- issue the first item's weight loads before touching the input vector
    constexpr int UNR = NV == 1 ? (NMAT == 1 ? SMB_GEMV_UNR1 : 4) : 8;  // 16-byte chunks in flight per matrix per lane per batch (NV > 1: one CTA per SM)
    uint4 wbuf[NMAT][UNR];
    int item = warp;
    auto issue = [&](int it, int batch) {
        const int row = r0 + it / nseg, seg = it % nseg;
        const int kbeg = seg * a.seg_len;
        const int kend = min(a.K, kbeg + a.seg_len);
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int k = kbeg + (batch * UNR + u) * 256 + lane * 8;
            if (k < kend) {
                wbuf[0][u] = ldg_stream(W0 + static_cast<size_t>(row) * a.K + k);
                if (NMAT == 2) wbuf[1][u] = ldg_stream(W1 + static_cast<size_t>(row) * a.K + k);
            }
        }
    };
    pdl_trigger();
    if (item < nitems) issue(item, 0);   // weights never depend on the previous kernel
    pdl_wait();

    // ---- stage pro(x) in shared memory (one vector after the other)
#pragma unroll 1
    for (int v = 0; v < NV; ++v) gemv_stage_vector<T>(a, v, xs + v * xpitch, red);
    __syncthreads();

    // ---- stream the items
    for (; item < nitems; item += kGemvWarps) {
        const int seg = item % nseg;
        const int kbeg = seg * a.seg_len;
        const int kend = min(a.K, kbeg + a.seg_len);
        const int nbatch = (kend - kbeg + UNR * 256 - 1) / (UNR * 256);
        float acc0[NV], acc1[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) { acc0[v] = 0.f; acc1[v] = 0.f; }
        for (int b = 0; b < nbatch; ++b) {
            if (b > 0) issue(item, b);
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int k = kbeg + (b * UNR + u) * 256 + lane * 8;
                if (k < kend) {
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        acc0[v] += dot8<T>(wbuf[0][u], xs + v * xpitch + k);
                        if (NMAT == 2) acc1[v] += dot8<T>(wbuf[1][u], xs + v * xpitch + k);
                    }
                }
            }
        }
        const int next = item + kGemvWarps;
        if (next < nitems) issue(next, 0);  // keep HBM requests in flight across the reduction
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            acc0[v] = warp_sum(acc0[v]);
            if (NMAT == 2) acc1[v] = warp_sum(acc1[v]);
        }
        if (lane == 0) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                part[(0 * NV + v) * nitems + item] = acc0[v];
                if (NMAT == 2) part[(1 * NV + v) * nitems + item] = acc1[v];
            }
        }
    }
    __syncthreads();

    // ---- epilogue: one thread per row, fixed-order sum of the K-segment partials
    for (int r = threadIdx.x; r < nrows; r += kGemvThreads) {
        const int n = r0 + r;
#pragma unroll 1
        for (int v = 0; v < NV; ++v) {       // vectors in order: the Mamba conv window rolls once per frame
        float acc = 0.f, accb = 0.f;
        for (int s = 0; s < nseg; ++s) {
            acc += part[(0 * NV + v) * nitems + r * nseg + s];
            if (NMAT == 2) accb += part[(1 * NV + v) * nitems + r * nseg + s];
        }
        gemv_finish_row<T, NMAT>(a, n, v, acc, accb);
        }   // v
    }
}

// ---------------------------------------------------------------------------------------------
// Mamba-1 single-step selective scan fused with dt_proj (K = dt_rank is one 16-byte chunk per lane
// for dt_rank = 256): one warp per channel d.
//   dt = softplus(T(W_dt[d,:] . xdb[:R]) + b_dt[d]);  h[d,n] = h[d,n]*exp(dt*A[d,n]) + dt*B[n]*x[d]
//   y[d] = T((sum_n h[d,n]*C[n] + D[d]*x[d]) * silu(z[d]))
// (selective_scan_fn semantics: /root/reference/streammind/model/mamba_ssm/ops/selective_scan_interface.py:91-157)
// ---------------------------------------------------------------------------------------------
struct ScanArgs {
    const void* W_dt;   // T [d_inner][dt_rank]
    const void* b_dt;   // T [d_inner]
    const void* A_log;  // T [d_inner][d_state]
    const void* D;      // T [d_inner]
    const void* xdb;    // T [dt_rank + 2*d_state]   (dt | B | C)
    const void* x;      // T [d_inner]  (conv + SiLU output)
    const void* z;      // T [d_inner]
    float* state;       // fp32 [d_inner][d_state]
    void* y;            // T [d_inner]
    int d_inner, dt_rank, d_state;
    int nv;             // consecutive frames processed in order (xdb / x / z / y strided per frame)
    long long xdb_stride, x_stride, z_stride, y_stride;
    long long state_stride;   // 0: the nv frames belong to one stream (state carried in a register); > 0: floats between the states
                              // of the nv streams the frames belong to (multi-stream batching)
};

template <typename T>
__global__ void __launch_bounds__(256) mamba_scan_step_kernel(const ScanArgs a) {
    extern __shared__ __align__(16) uint8_t scan_smem[];
    pdl_trigger();
    pdl_wait();
    T* xdb = reinterpret_cast<T*>(scan_smem);                 // [nv][nxp]
    const int nx = a.dt_rank + 2 * a.d_state, nxp = (nx + 7) & ~7;
    const int nv = max(1, a.nv);
    for (int i = threadIdx.x; i < nx * nv; i += blockDim.x) {
        const int v = i / nx, j = i % nx;
        xdb[v * nxp + j] = reinterpret_cast<const T*>(a.xdb)[v * a.xdb_stride + j];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int d = blockIdx.x * wpb + warp; d < a.d_inner; d += gridDim.x * wpb) {
        const T* w = reinterpret_cast<const T*>(a.W_dt) + static_cast<size_t>(d) * a.dt_rank;
        const float bdt = Cvt<T>::to_f(reinterpret_cast<const T*>(a.b_dt)[d]);
        const float Dd = Cvt<T>::to_f(reinterpret_cast<const T*>(a.D)[d]);
        // this lane's state element (d_state <= 32) stays in a register across the frames of the batch
        const bool has_n = lane < a.d_state;
        float* sp = a.state + static_cast<size_t>(d) * a.d_state + lane;
        float hst = has_n ? *sp : 0.f;
        const float A = has_n ? -__expf(Cvt<T>::to_f(reinterpret_cast<const T*>(a.A_log)[static_cast<size_t>(d) * a.d_state + lane])) : 0.f;
        for (int v = 0; v < nv; ++v) {
            if (a.state_stride != 0) {          // frame v belongs to its own stream: swap the state in and out
                if (v > 0 && has_n) sp[(v - 1) * a.state_stride] = hst;
                hst = has_n ? sp[v * a.state_stride] : 0.f;
            }
            const T* xv_db = xdb + v * nxp;
            float acc = 0.f;
            for (int k = lane * 8; k < a.dt_rank; k += 256) {
                if (k + 8 <= a.dt_rank) {
                    acc += dot8<T>(ldg_stream(w + k), xv_db + k);
                } else {
                    for (int j = k; j < a.dt_rank; ++j) acc += Cvt<T>::to_f(w[j]) * Cvt<T>::to_f(xv_db[j]);
                }
            }
            acc = warp_sum(acc);
            const float dtl = rnd<T>(acc) + bdt;
            const float dt = dtl > 20.0f ? dtl : log1pf(__expf(dtl));  // F.softplus, threshold 20
            const float xv = Cvt<T>::to_f(reinterpret_cast<const T*>(a.x)[v * a.x_stride + d]);
            float contrib = 0.f;
            if (has_n) {
                const float Bn = Cvt<T>::to_f(xv_db[a.dt_rank + lane]);
                const float Cn = Cvt<T>::to_f(xv_db[a.dt_rank + a.d_state + lane]);
                hst = hst * __expf(dt * A) + dt * Bn * xv;
                contrib = hst * Cn;
            }
            contrib = warp_sum(contrib);
            if (lane == 0) {
                const float zv = Cvt<T>::to_f(reinterpret_cast<const T*>(a.z)[v * a.z_stride + d]);
                const float yv = (contrib + Dd * xv) * silu_f(zv);
                reinterpret_cast<T*>(a.y)[v * a.y_stride + d] = Cvt<T>::from_f(yv);
            }
        }
        if (has_n) sp[a.state_stride != 0 ? (nv - 1) * a.state_stride : 0] = hst;
    }
}

}  // namespace smb
