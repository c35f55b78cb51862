// Debug views of the reference: the search visualisations
// (lane_tracker.py:689-771) and the fixed-point bilinear resize behind the split view (utils.py:88).
// These are not on the per-frame hot path; they exist so that process(visualize_search / split_view) returns what
// the reference returns, rendered on the device from the buffers the tracker already holds there.
#include "lt_common.cuh"

// cv2.addWeighted(a, 1, b, beta, 0) on one uint8 value: float32, round half to even, saturate
__device__ __forceinline__ uint8_t add_weighted_u8(uint32_t a, uint32_t b, float beta) {
    float f = rintf(__fadd_rn((float)a, __fmul_rn((float)b, beta)));
    return (uint8_t)fminf(fmaxf(f, 0.f), 255.f);
}

// Base image: the binary mask on three channels; in the sliding-window view the search windows are blended in
// first (template = 255 per side, the uint8 sum of both sides wraps to 254 where they overlap; weight 0.5).
// rects: [n][5] = {row0, row1, col0, col1, side}, half-open, already resolved with Python's slice rules.
__global__ void __launch_bounds__(256)
k_vis_base(const uint8_t* __restrict__ mask, const int* __restrict__ rects, int nrect, int W, int H,
           uint8_t* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const uint32_t v = mask[(size_t)y * W + x];
    bool in_l = false, in_r = false;
    for (int i = 0; i < nrect; ++i) {
        const int* r = rects + 5 * i;
        if (y >= r[0] && y < r[1] && x >= r[2] && x < r[3]) { if (r[4]) in_r = true; else in_l = true; }
    }
    const uint32_t t = ((in_l ? 255u : 0u) + (in_r ? 255u : 0u)) & 255u;
    uint8_t* o = out + ((size_t)y * W + x) * 3;
    o[0] = (uint8_t)v;
    o[1] = nrect ? add_weighted_u8(v, t, 0.5f) : (uint8_t)v;
    o[2] = (uint8_t)v;
}

// out[y, x] = colour for every pixel of a packed list (y << 16 | x + 32768)
__global__ void k_vis_scatter(const uint32_t* __restrict__ px, int n, int W, int H, uint32_t rgb, uint8_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t v = px[i];
    const int y = (int)(v >> 16), x = (int)(v & 0xFFFFu) - 32768;
    if ((unsigned)y >= (unsigned)H || (unsigned)x >= (unsigned)W) return;
    uint8_t* o = out + ((size_t)y * W + x) * 3;
    o[0] = rgb & 255; o[1] = (rgb >> 8) & 255; o[2] = (rgb >> 16) & 255;
}

// graph points of a polynomial (get_poly_points output: xs[i] on row H - n + i)
__global__ void k_vis_scatter_poly(const int* __restrict__ xs, const int* __restrict__ count, int W, int H, uint32_t rgb,
                                   uint8_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, n = *count;
    if (i >= n) return;
    const int y = H - n + i, x = xs[i];
    if ((unsigned)y >= (unsigned)H || (unsigned)x >= (unsigned)W) return;
    uint8_t* o = out + ((size_t)y * W + x) * 3;
    o[0] = rgb & 255; o[1] = (rgb >> 8) & 255; o[2] = (rgb >> 16) & 255;
}

// band view: addWeighted(output, 1, window_img, 0.3, 0) where window_img is (0,255,0) inside either band polygon
__global__ void __launch_bounds__(256)
k_vis_band_blend(const int2* __restrict__ rows_l, const int2* __restrict__ rows_r, int W, int H, uint8_t* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int2 a = rows_l[y], b = rows_r[y];
    if ((x >= a.x && x <= a.y) || (x >= b.x && x <= b.y)) {
        uint8_t* o = out + ((size_t)y * W + x) * 3;
        o[1] = add_weighted_u8(o[1], 255u, 0.3f);
    }
}

int lt_launch_vis_base(const uint8_t* d_mask, const int* d_rects, int nrect, int W, int H, uint8_t* d_out, cudaStream_t st) {
    k_vis_base<<<dim3(lt_div_up(W, 256), H), 256, 0, st>>>(d_mask, d_rects, nrect, W, H, d_out);
    LT_LAUNCH_CHECK();
    return 0;
}
int lt_launch_vis_scatter(const uint32_t* d_px, int n, int W, int H, uint32_t rgb, uint8_t* d_out, cudaStream_t st) {
    if (n <= 0) return 0;
    k_vis_scatter<<<lt_div_up(n, 256), 256, 0, st>>>(d_px, n, W, H, rgb, d_out);
    LT_LAUNCH_CHECK();
    return 0;
}
int lt_launch_vis_scatter_poly(const int* d_xs, const int* d_count, int W, int H, uint32_t rgb, uint8_t* d_out, cudaStream_t st) {
    k_vis_scatter_poly<<<lt_div_up(H, 256), 256, 0, st>>>(d_xs, d_count, W, H, rgb, d_out);
    LT_LAUNCH_CHECK();
    return 0;
}
int lt_launch_vis_band_blend(const int2* rows_l, const int2* rows_r, int W, int H, uint8_t* d_out, cudaStream_t st) {
    k_vis_band_blend<<<dim3(lt_div_up(W, 256), H), 256, 0, st>>>(rows_l, rows_r, W, H, d_out);
    LT_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------
// cv2.resize(img, dsize) for uint8, INTER_LINEAR: OpenCV's fixed-point path.  Horizontal taps carry 11-bit weights
// rint(w * 2048) computed in float32 from (d + 0.5) * scale - 0.5 (weight forced to 0 where the tap pair leaves the
// row); the vertical pass clamps row indices instead and combines as
//     (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2
// ---------------------------------------------------------------------------

struct ResizeTap { int i0, i1, w0, w1; };

__device__ __forceinline__ ResizeTap resize_tap(int d, int sn, double scale, bool vertical) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (!vertical && (s < 0 || s >= sn - 1)) f = 0.f;
    ResizeTap t;
    t.w1 = (int)rintf(__fmul_rn(f, 2048.f));
    t.w0 = (int)rintf(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    t.i0 = min(max(s, 0), sn - 1);
    t.i1 = min(max(s + 1, 0), sn - 1);
    return t;
}

__global__ void __launch_bounds__(256)
k_resize_linear(const uint8_t* __restrict__ src, int sw, int sh, int cn, size_t src_pitch, uint8_t* __restrict__ dst,
                int dw, int dh, size_t dst_pitch, double scale_x, double scale_y) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw) return;
    const ResizeTap tx = resize_tap(x, sw, scale_x, false), ty = resize_tap(y, sh, scale_y, true);
    const uint8_t* r0 = src + (size_t)ty.i0 * src_pitch;
    const uint8_t* r1 = src + (size_t)ty.i1 * src_pitch;
    for (int c = 0; c < cn; ++c) {
        const int s0 = r0[tx.i0 * cn + c] * tx.w0 + r0[tx.i1 * cn + c] * tx.w1;
        const int s1 = r1[tx.i0 * cn + c] * tx.w0 + r1[tx.i1 * cn + c] * tx.w1;
        const int v = (((ty.w0 * (s0 >> 4)) >> 16) + ((ty.w1 * (s1 >> 4)) >> 16) + 2) >> 2;
        dst[(size_t)y * dst_pitch + x * cn + c] = (uint8_t)min(max(v, 0), 255);
    }
}

int lt_launch_resize_linear(const uint8_t* d_src, int sw, int sh, int cn, size_t src_pitch, uint8_t* d_dst, int dw, int dh,
                            size_t dst_pitch, cudaStream_t st) {
    k_resize_linear<<<dim3(lt_div_up(dw, 256), dh), 256, 0, st>>>(d_src, sw, sh, cn, src_pitch, d_dst, dw, dh, dst_pitch,
                                                                  (double)sw / dw, (double)sh / dh);
    LT_LAUNCH_CHECK();
    return 0;
}
