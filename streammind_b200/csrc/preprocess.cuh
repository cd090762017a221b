// Frame preprocessing on the device: uint8 HWC frames -> pad to square (virtual) -> Pillow's 8-bit bicubic
// resample to the tower's input size (two passes with a uint8 intermediate, fixed-point coefficients) -> rescale /
// normalise / round to the model dtype through a 3 x 256 table.  Replaces the per-frame PIL + numpy work of
// mm_utils.py:446-464 (process_video, aspect_ratio 'pad') -> expand2square (mm_utils.py:257-268) ->
// CLIPImageProcessor.preprocess (transformers 4.44.2) -> PIL.Image.resize (Pillow 9.4.0, Resample.c).
// Byte / integer work, HBM-bound: per 1080p frame 6.2 MB read, 1.9 MB intermediate written and read, 0.68 MB written.
#pragma once
#include <cstdint>
#include <cmath>
#include <vector>

namespace smb {

constexpr int kPreBits = 32 - 8 - 2;     // Resample.c PRECISION_BITS

struct PreArgs {
    const uint8_t* src;      // [n, H, W, 3]
    uint8_t* tmp;            // [n, S, out, 3]: the padded square after the horizontal pass
    const int* bounds;       // [out, 2] first tap, tap count (same table for both passes: the padded image is square)
    const int* kk;           // [out, ksize] fixed-point taps (vertical pass: one row per block, broadcast reads)
    const int* kk_t;         // [ksize, out] the same taps transposed (horizontal pass: neighbouring threads, neighbouring ints)
    int H, W, S, pad_x, pad_y, out, ksize;
    int bg0, bg1, bg2;       // background colour of expand2square
};

__device__ __forceinline__ int pre_clip8(int ss) {
    const int v = ss >> kPreBits;          // arithmetic shift, as Resample.c's clip8
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: CTA = one row of the (virtual) padded square.  The source row is staged in shared memory with coalesced
// 32-bit loads (byte loads when the row is not 4-byte aligned), then every thread produces output columns from it.
__global__ void __launch_bounds__(128) preprocess_h_kernel(const PreArgs a) {
    extern __shared__ __align__(16) uint8_t pre_row[];
    const int y = blockIdx.x, f = blockIdx.y;
    uint8_t* o = a.tmp + (static_cast<size_t>(f) * a.S + y) * a.out * 3;
    const int sy = y - a.pad_y;
    if (sy < 0 || sy >= a.H) {       // a padding row: every tap sees the background colour, and taps that sum to 1 +- ksize * 2^-23 return it
        for (int xx = threadIdx.x; xx < a.out; xx += blockDim.x) {
            o[xx * 3] = static_cast<uint8_t>(a.bg0); o[xx * 3 + 1] = static_cast<uint8_t>(a.bg1); o[xx * 3 + 2] = static_cast<uint8_t>(a.bg2);
        }
        return;
    }
    const uint8_t* row = a.src + (static_cast<size_t>(f) * a.H + sy) * a.W * 3;
    const int nbytes = a.W * 3;
    if (((reinterpret_cast<uintptr_t>(row) | static_cast<uintptr_t>(nbytes)) & 3) == 0) {
        const uint32_t* src32 = reinterpret_cast<const uint32_t*>(row);
        uint32_t* dst32 = reinterpret_cast<uint32_t*>(pre_row);
        for (int i = threadIdx.x; i < nbytes / 4; i += blockDim.x) dst32[i] = __ldg(src32 + i);
    } else {
        for (int i = threadIdx.x; i < nbytes; i += blockDim.x) pre_row[i] = row[i];
    }
    __syncthreads();
    for (int xx = threadIdx.x; xx < a.out; xx += blockDim.x) {
        const int xmin = a.bounds[2 * xx], n = a.bounds[2 * xx + 1];
        const int* k = a.kk_t + xx;
        int s0 = 1 << (kPreBits - 1), s1 = s0, s2 = s0;
        for (int x = 0; x < n; ++x) {
            const int sx = xmin + x - a.pad_x;
            int v0 = a.bg0, v1 = a.bg1, v2 = a.bg2;
            if (sx >= 0 && sx < a.W) { v0 = pre_row[sx * 3]; v1 = pre_row[sx * 3 + 1]; v2 = pre_row[sx * 3 + 2]; }
            const int w = k[static_cast<size_t>(x) * a.out];
            s0 += v0 * w; s1 += v1 * w; s2 += v2 * w;
        }
        o[xx * 3] = static_cast<uint8_t>(pre_clip8(s0)); o[xx * 3 + 1] = static_cast<uint8_t>(pre_clip8(s1)); o[xx * 3 + 2] = static_cast<uint8_t>(pre_clip8(s2));
    }
}

// vertical pass + rescale / normalise / round (table) + channels-first store: thread = (output column, output row)
template <typename T>
__global__ void __launch_bounds__(128) preprocess_v_kernel(const PreArgs a, const T* __restrict__ lut /*[3][256]*/, T* __restrict__ out /*[n,3,out,out]*/) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x, yy = blockIdx.y, f = blockIdx.z;
    if (xx >= a.out) return;
    const int ymin = a.bounds[2 * yy], n = a.bounds[2 * yy + 1];
    const int* k = a.kk + static_cast<size_t>(yy) * a.ksize;
    int s0 = 1 << (kPreBits - 1), s1 = s0, s2 = s0;
    const uint8_t* p = a.tmp + ((static_cast<size_t>(f) * a.S + ymin) * a.out + xx) * 3;
    for (int y = 0; y < n; ++y, p += static_cast<size_t>(a.out) * 3) {
        const int w = k[y];
        s0 += p[0] * w; s1 += p[1] * w; s2 += p[2] * w;
    }
    const size_t plane = static_cast<size_t>(a.out) * a.out;
    T* o = out + static_cast<size_t>(f) * 3 * plane + static_cast<size_t>(yy) * a.out + xx;
    o[0] = lut[pre_clip8(s0)];
    o[plane] = lut[256 + pre_clip8(s1)];
    o[2 * plane] = lut[512 + pre_clip8(s2)];
}

// the same pass with 32-bit loads (out * 3 a multiple of 4, e.g. 336): thread = 4 consecutive bytes of the interleaved row
template <typename T>
__global__ void __launch_bounds__(128) preprocess_v4_kernel(const PreArgs a, const T* __restrict__ lut, T* __restrict__ out) {
    const int words = a.out * 3 / 4;
    const int wi = blockIdx.x * blockDim.x + threadIdx.x, yy = blockIdx.y, f = blockIdx.z;
    if (wi >= words) return;
    const int ymin = a.bounds[2 * yy], n = a.bounds[2 * yy + 1];
    const int* k = a.kk + static_cast<size_t>(yy) * a.ksize;
    int s[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) s[j] = 1 << (kPreBits - 1);
    const uint32_t* p = reinterpret_cast<const uint32_t*>(a.tmp + (static_cast<size_t>(f) * a.S + ymin) * a.out * 3) + wi;
    for (int y = 0; y < n; ++y, p += words) {
        const uint32_t v = *p;
        const int w = k[y];
        s[0] += static_cast<int>(v & 255u) * w;
        s[1] += static_cast<int>((v >> 8) & 255u) * w;
        s[2] += static_cast<int>((v >> 16) & 255u) * w;
        s[3] += static_cast<int>(v >> 24) * w;
    }
    const size_t plane = static_cast<size_t>(a.out) * a.out;
    T* o = out + static_cast<size_t>(f) * 3 * plane + static_cast<size_t>(yy) * a.out;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = wi * 4 + j, x = i / 3, c = i - 3 * x;
        o[c * plane + x] = lut[c * 256 + pre_clip8(s[j])];
    }
}

// Resample.c precompute_coeffs (bicubic, box = the whole axis) + normalize_coeffs_8bpc, in double like Pillow
inline double pre_bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}
inline int pre_build_table(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk) {
    const double scale = static_cast<double>(static_cast<float>(in_size) - 0.0f) / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 2.0 * filterscale;
    const int ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
    bounds.assign(static_cast<size_t>(out_size) * 2, 0);
    kk.assign(static_cast<size_t>(out_size) * ksize, 0);
    std::vector<double> w(ksize);
    const double ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        int xmin = static_cast<int>(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = static_cast<int>(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            w[x] = pre_bicubic((x + xmin - center + 0.5) * ss);
            ww += w[x];
        }
        for (int x = 0; x < xmax; ++x) {
            if (ww != 0.0) w[x] /= ww;
            kk[static_cast<size_t>(xx) * ksize + x] = w[x] < 0 ? static_cast<int>(-0.5 + w[x] * (1 << kPreBits))
                                                                : static_cast<int>(0.5 + w[x] * (1 << kPreBits));
        }
        bounds[2 * xx] = xmin;
        bounds[2 * xx + 1] = xmax;
    }
    return ksize;
}

}  // namespace smb
