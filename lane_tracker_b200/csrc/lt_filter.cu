// filter_lane_points (lane_tracker.py:183-240) on pair-packed planes:
//   ellipse top-hat 29 (R) / 55 (Lab b)  -> cross ("bilateral") threshold  \
//   or box-mean adaptive threshold on the raw planes                        > OR -> open 5x5 -> bit mask
//
// The ellipse erosion/dilation is the dominant cost of the whole path.  It is computed by
// row-span decomposition:  out(y,x) = op_{dy} Hop_{hw[dy]}(y+dy, x)  where
// Hop_w is the horizontal window min/max of half-width w.  Per source row the CTA builds
// power-of-two window tables (4, 8, 16, 32) in shared memory; every thread owns one packed
// column (two image strips in the two u16 lanes), derives the <=17 distinct Hop_w values of
// the row from two table reads each, and folds them into a K-deep register pipeline
// A[j] = op(A[j+1], Hop_{hw[j]}) whose head is a finished output row.  All min/max are
// single VIMNMX(3).U16x2 instructions.
#include "lt_common.cuh"

// ---------------------------------------------------------------------------
// structuring elements: cv2.getStructuringElement(MORPH_ELLIPSE,(k,k)) row half-widths
// ---------------------------------------------------------------------------

template <int K> struct Ellipse;

template <> struct Ellipse<55> {
    static constexpr int R = 27, ND = 17;
    __host__ __device__ static constexpr int hw(int j) {
        constexpr int t[55] = {0, 7, 10, 12, 14, 16, 17, 18, 19, 20, 21, 22, 22, 23, 24, 24, 25, 25, 25,
                               26, 26, 26, 27, 27, 27, 27, 27, 27, 27, 27, 27, 27, 27, 26, 26, 26, 25,
                               25, 25, 24, 24, 23, 22, 22, 21, 20, 19, 18, 17, 16, 14, 12, 10, 7, 0};
        return t[j];
    }
    __host__ __device__ static constexpr int uniq(int i) {
        constexpr int t[17] = {0, 7, 10, 12, 14, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27};
        return t[i];
    }
};

template <> struct Ellipse<29> {
    static constexpr int R = 14, ND = 9;
    __host__ __device__ static constexpr int hw(int j) {
        constexpr int t[29] = {0, 5, 7, 9, 10, 11, 11, 12, 13, 13, 13, 14, 14, 14, 14,
                               14, 14, 14, 13, 13, 13, 12, 11, 11, 10, 9, 7, 5, 0};
        return t[j];
    }
    __host__ __device__ static constexpr int uniq(int i) {
        constexpr int t[9] = {0, 5, 7, 9, 10, 11, 12, 13, 14};
        return t[i];
    }
};

template <int K> __host__ __device__ constexpr int ell_uidx(int w) {
    int r = 0;
    for (int i = 0; i < Ellipse<K>::ND; ++i)
        if (Ellipse<K>::uniq(i) == w) r = i;
    return r;
}

// ---------------------------------------------------------------------------
// the morphology kernel
// ---------------------------------------------------------------------------

constexpr int MORPH_TW = 288;       // packed columns per CTA (= threads)
constexpr int MORPH_RB = 8;         // source rows per table build

template <bool IS_MAX> __device__ __forceinline__ uint32_t op2(uint32_t a, uint32_t b) {
    return IS_MAX ? __vmaxu2(a, b) : __vminu2(a, b);
}
template <bool IS_MAX> __device__ __forceinline__ uint32_t op3(uint32_t a, uint32_t b, uint32_t c) {
    return IS_MAX ? __vimax3_u16x2(a, b, c) : __vimin3_u16x2(a, b, c);
}

// One staged element: packed pixel pair at plane row r, packed column gx (may lie outside the plane).
template <bool IS_MAX>
__device__ __forceinline__ uint32_t stage_elem(const uint32_t* __restrict__ src, const LtDims& d, int r, int gx) {
    constexpr uint32_t PADL = IS_MAX ? 0u : 0xFFFFu;
    constexpr uint32_t PAD2 = PADL | (PADL << 16);
    if ((unsigned)r >= (unsigned)d.bv_h) return PAD2;
    const uint32_t* row = src + (size_t)r * d.p2;
    if (gx >= 0 && gx + d.p2 < d.bv_w) return __ldg(&row[gx]);        // interior: both lanes real pixels
    uint32_t lo = PADL, hi = PADL;
    if (gx < 0) {
        if (gx + d.p2 >= 0) hi = __ldg(&row[gx + d.p2]) & 0xFFFFu;     // image col gx+p2 lives in the low strip
    } else if (gx < d.p2) {
        lo = __ldg(&row[gx]) & 0xFFFFu;                               // hi lane: col >= bv_w -> outside
    } else {
        if (gx < d.bv_w && gx - d.p2 < d.p2) lo = __ldg(&row[gx - d.p2]) >> 16;   // image col gx lives in the high strip
    }
    return lo | (hi << 16);
}

template <int K, bool IS_MAX, bool TOPHAT>
__global__ void __launch_bounds__(MORPH_TW, 2)
k_morph(const uint32_t* __restrict__ src_all, uint32_t* __restrict__ dst_all, const uint32_t* __restrict__ orig_all,
        LtDims d, int band_rows, size_t stream_stride, const int* __restrict__ list, const int* __restrict__ count) {
    using E = Ellipse<K>;
    constexpr int R = E::R;
    constexpr int TW = MORPH_TW, RB = MORPH_RB;
    constexpr int TE = TW + 2 * R;          // staged elements per row
    constexpr int TEA = TE + 32;            // + slack that always holds PAD
    constexpr bool HAS32 = (2 * R + 1) >= 32;
    constexpr int NTAB = HAS32 ? 5 : 4;
    constexpr uint32_t PADL = IS_MAX ? 0u : 0xFFFFu;
    constexpr uint32_t PAD2 = PADL | (PADL << 16);
    constexpr int NPF = (RB * TE + TW - 1) / TW;

    int slot = blockIdx.z;
    if (count != nullptr && slot >= *count) return;
    int s = list ? list[slot] : slot;

    extern __shared__ uint32_t smem[];
    uint32_t* T0 = smem;
    uint32_t* T4 = T0 + RB * TEA;
    uint32_t* T8 = T4 + RB * TEA;
    uint32_t* T16 = T8 + RB * TEA;
    uint32_t* T32 = T16 + RB * TEA;   // only touched when HAS32

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW;
    const int yb0 = blockIdx.y * band_rows;
    const int yb1 = min(yb0 + band_rows, d.bv_h);
    const uint32_t* src = src_all + (size_t)s * stream_stride;
    uint32_t* dst = dst_all + (size_t)s * stream_stride;
    const uint32_t* orig = TOPHAT ? orig_all + (size_t)s * stream_stride : nullptr;

    for (int i = tid; i < NTAB * RB * TEA; i += TW) smem[i] = PAD2;

    const int r_begin = yb0 - R;
    const int r_end = yb1 + R;      // exclusive
    const int nblk = (r_end - r_begin + RB - 1) / RB;

    uint32_t pf[NPF];
#pragma unroll
    for (int q = 0; q < NPF; ++q) {
        int e = tid + q * TW;
        int rr = e / TE, i = e - rr * TE;
        pf[q] = (e < RB * TE) ? stage_elem<IS_MAX>(src, d, r_begin + rr, x0 + i - R) : PAD2;
    }

    uint32_t A[K];
#pragma unroll
    for (int j = 0; j < K; ++j) A[j] = PAD2;

    const int gx = x0 + tid;                       // this thread's packed column
    const bool col_ok = gx < d.p2;
    const uint32_t lane_mask = (gx + d.p2 < d.bv_w) ? 0xFFFFFFFFu : 0x0000FFFFu;
    __syncthreads();

    for (int blk = 0; blk < nblk; ++blk) {
        const int rb0 = r_begin + blk * RB;
        // publish the staged rows, then start fetching the next block
#pragma unroll
        for (int q = 0; q < NPF; ++q) {
            int e = tid + q * TW;
            int rr = e / TE, i = e - rr * TE;
            if (e < RB * TE) T0[rr * TEA + i] = pf[q];
        }
        __syncthreads();
        if (blk + 1 < nblk) {
#pragma unroll
            for (int q = 0; q < NPF; ++q) {
                int e = tid + q * TW;
                int rr = e / TE, i = e - rr * TE;
                pf[q] = (e < RB * TE) ? stage_elem<IS_MAX>(src, d, rb0 + RB + rr, x0 + i - R) : PAD2;
            }
        }
        // window tables: T4 -> (T8, T16) -> T32
        for (int e = tid; e < RB * TE; e += TW) {
            int rr = e / TE, i = e - rr * TE;
            const uint32_t* t = T0 + rr * TEA + i;
            T4[rr * TEA + i] = op2<IS_MAX>(op3<IS_MAX>(t[0], t[1], t[2]), t[3]);
        }
        __syncthreads();
        for (int e = tid; e < RB * TE; e += TW) {
            int rr = e / TE, i = e - rr * TE;
            const uint32_t* t = T4 + rr * TEA + i;
            uint32_t v8 = op2<IS_MAX>(t[0], t[4]);
            T8[rr * TEA + i] = v8;
            T16[rr * TEA + i] = op3<IS_MAX>(v8, t[8], t[12]);
        }
        __syncthreads();
        if (HAS32) {
            for (int e = tid; e < RB * TE; e += TW) {
                int rr = e / TE, i = e - rr * TE;
                const uint32_t* t = T16 + rr * TEA + i;
                T32[rr * TEA + i] = op2<IS_MAX>(t[0], t[16]);
            }
            __syncthreads();
        }
        // walk the RB rows of this block
#pragma unroll 1
        for (int rr = 0; rr < RB; ++rr) {
            const int r = rb0 + rr;
            if (r >= r_end) break;
            const int y = r - R;                    // output row completed by source row r
            uint32_t og = 0;
            const bool emit = col_ok && y >= yb0;   // y < yb1 is implied by r < r_end
            if (TOPHAT && emit) og = __ldg(&orig[(size_t)y * d.p2 + gx]);
            const int base = rr * TEA + tid + R;    // index of this thread's column in the tables
            uint32_t H[E::ND];
#pragma unroll
            for (int u = 0; u < E::ND; ++u) {
                constexpr int dummy = 0; (void)dummy;
                const int w = E::uniq(u);
                const int len = 2 * w + 1;
                if (w == 0) {
                    H[u] = T0[base];
                } else if (len >= 32) {
                    H[u] = op2<IS_MAX>(T32[base - w], T32[base + w - 31]);
                } else if (len >= 16) {
                    H[u] = op2<IS_MAX>(T16[base - w], T16[base + w - 15]);
                } else {
                    H[u] = op2<IS_MAX>(T8[base - w], T8[base + w - 7]);
                }
            }
#pragma unroll
            for (int j = 0; j < K - 1; ++j) A[j] = op2<IS_MAX>(A[j + 1], H[ell_uidx<K>(E::hw(j))]);
            A[K - 1] = H[ell_uidx<K>(E::hw(K - 1))];
            if (emit) {
                uint32_t v = A[0];
                if (TOPHAT) v = og - v;             // open <= src per lane: no borrow between lanes
                dst[(size_t)y * d.p2 + gx] = v & lane_mask;
            }
        }
        __syncthreads();
    }
}

template <int K, bool IS_MAX, bool TOPHAT>
static int launch_morph(lt_handle* h, const uint32_t* src, uint32_t* dst, const uint32_t* orig, int n,
                        const int* list, const int* count, int bands, cudaStream_t st) {
    constexpr int R = Ellipse<K>::R;
    constexpr int TEA = MORPH_TW + 2 * R + 32;
    constexpr int NTAB = (2 * R + 1 >= 32) ? 5 : 4;
    size_t smem = (size_t)NTAB * MORPH_RB * TEA * sizeof(uint32_t);
    static bool attr_done = false;
    if (!attr_done) {
        LT_CUDA(cudaFuncSetAttribute(k_morph<K, IS_MAX, TOPHAT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
        attr_done = true;
    }
    const LtDims& d = h->d;
    int band_rows = lt_div_up(d.bv_h, bands);
    dim3 g(lt_div_up(d.p2, MORPH_TW), lt_div_up(d.bv_h, band_rows), n);
    k_morph<K, IS_MAX, TOPHAT><<<g, MORPH_TW, smem, st>>>(src, dst, orig, d, band_rows, h->stream_plane, list, count);
    LT_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------
// cross ("bilateral") threshold, bilateral_adaptive_threshold (lane_tracker.py:14-83)
//   pass <=> (L < t and R < t) or (U < t and D < t),  t = k*p - C*k, zero-padded side sums
// horizontal half: one warp per row, exclusive prefix sums of the linearised row in smem
// vertical half:   one thread per packed column, running sums in packed u16x2 registers
// ---------------------------------------------------------------------------

constexpr int ROWK_WARPS = 8;

// exclusive prefix sums E[0..W] of one image row into shared memory (one warp)
__device__ __forceinline__ void warp_row_prefix(const uint32_t* __restrict__ prow, const LtDims& d,
                                                uint32_t* lin, uint32_t* E, int lane) {
    for (int x = lane; x < d.p2; x += 32) {
        uint32_t v = __ldg(&prow[x]);
        lin[x] = v & 0xFFFFu;
        if (x + d.p2 < d.bv_w) lin[x + d.p2] = v >> 16;
    }
    __syncwarp();
    uint32_t carry = 0;
    for (int b = 0; b < d.bv_w; b += 32) {
        int x = b + lane;
        uint32_t v = (x < d.bv_w) ? lin[x] : 0u, inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += n;
        }
        if (x < d.bv_w) E[x] = carry + inc - v;
        carry += __shfl_sync(0xFFFFFFFFu, inc, 31);
    }
    if (lane == 0) E[d.bv_w] = carry;
    __syncwarp();
}

__global__ void __launch_bounds__(ROWK_WARPS * 32)
k_cross_h(const uint32_t* __restrict__ plane_all, uint32_t* __restrict__ bits_all, LtDims d, int k, int C,
          int accumulate, size_t plane_stride, size_t bits_stride, const int* __restrict__ list,
          const int* __restrict__ count) {
    int slot = blockIdx.y;
    if (count != nullptr && slot >= *count) return;
    int s = list ? list[slot] : slot;
    extern __shared__ uint32_t smem[];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int y = blockIdx.x * ROWK_WARPS + warp;
    if (y >= d.bv_h) return;
    int wpad = (d.bv_w + 32) & ~31;
    uint32_t* lin = smem + (size_t)warp * 2 * wpad;
    uint32_t* E = lin + wpad;
    warp_row_prefix(plane_all + (size_t)s * plane_stride + (size_t)y * d.p2, d, lin, E, lane);
    uint32_t* brow = bits_all + (size_t)s * bits_stride + (size_t)y * d.mwords;
    const int Ck = C * k;
    for (int wd = 0; wd < d.mwords; ++wd) {
        int c = wd * 32 + lane;
        bool pass = false;
        if (c < d.bv_w) {
            int t = k * (int)lin[c] - Ck;
            int L = (int)(E[c] - E[max(c - k, 0)]);
            int Rs = (int)(E[min(c + k + 1, d.bv_w)] - E[c + 1]);
            pass = (L < t) && (Rs < t);
        }
        uint32_t b = __ballot_sync(0xFFFFFFFFu, pass);
        if (lane == 0) brow[wd] = accumulate ? (brow[wd] | b) : b;
    }
}

__global__ void __launch_bounds__(32)
k_cross_v(const uint32_t* __restrict__ plane_all, uint32_t* __restrict__ bits_all, LtDims d, int k, int C,
          int band_rows, size_t plane_stride, size_t bits_stride, const int* __restrict__ list,
          const int* __restrict__ count) {
    int slot = blockIdx.z;
    if (count != nullptr && slot >= *count) return;
    int s = list ? list[slot] : slot;
    int lane = threadIdx.x;
    int x = blockIdx.x * 32 + lane;              // packed column; p2 is a multiple of 32
    int yb0 = blockIdx.y * band_rows, yb1 = min(yb0 + band_rows, d.bv_h);
    const uint32_t* P = plane_all + (size_t)s * plane_stride + x;
    uint32_t* bits = bits_all + (size_t)s * bits_stride;
    const bool hi_ok = x + d.p2 < d.bv_w;
    auto ld = [&](int r) -> uint32_t { return ((unsigned)r < (unsigned)d.bv_h) ? __ldg(&P[(size_t)r * d.p2]) : 0u; };
    uint32_t U = 0, D = 0;
    for (int i = 1; i <= k; ++i) { U += ld(yb0 - i); D += ld(yb0 + i); }
    const int Ck = C * k;
    uint32_t p = ld(yb0);
    for (int y = yb0; y < yb1; ++y) {
        uint32_t pn = ld(y + 1);
        int tl = k * (int)(p & 0xFFFFu) - Ck, th = k * (int)(p >> 16) - Ck;
        bool pl = ((int)(U & 0xFFFFu) < tl) && ((int)(D & 0xFFFFu) < tl);
        bool ph = hi_ok && ((int)(U >> 16) < th) && ((int)(D >> 16) < th);
        uint32_t bl = __ballot_sync(0xFFFFFFFFu, pl), bh = __ballot_sync(0xFFFFFFFFu, ph);
        if (lane == 0) {
            uint32_t* brow = bits + (size_t)y * d.mwords;
            brow[blockIdx.x] |= bl;
            brow[blockIdx.x + (d.p2 >> 5)] |= bh;
        }
        U = U + p - ld(y - k);                    // lanes stay in [0, 65535]: add first, then subtract
        D = D + ld(y + k + 1) - pn;
        p = pn;
    }
}

// ---------------------------------------------------------------------------
// cv2.adaptiveThreshold(MEAN_C, THRESH_BINARY, block, -c) (lane_tracker.py:217-218)
// box sum with replicated border = row sums (k_box_h) then running column sums (k_box_v)
// ---------------------------------------------------------------------------

__global__ void __launch_bounds__(ROWK_WARPS * 32)
k_box_h(const uint32_t* __restrict__ plane_all, uint32_t* __restrict__ hs_all, LtDims d, int half,
        size_t plane_stride, const int* __restrict__ list, const int* __restrict__ count) {
    int slot = blockIdx.y;
    if (count != nullptr && slot >= *count) return;
    int s = list ? list[slot] : slot;
    extern __shared__ uint32_t smem[];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int y = blockIdx.x * ROWK_WARPS + warp;
    if (y >= d.bv_h) return;
    int wpad = (d.bv_w + 32) & ~31;
    uint32_t* lin = smem + (size_t)warp * 2 * wpad;
    uint32_t* E = lin + wpad;
    warp_row_prefix(plane_all + (size_t)s * plane_stride + (size_t)y * d.p2, d, lin, E, lane);
    uint32_t* hrow = hs_all + (size_t)s * plane_stride + (size_t)y * d.p2;
    const int W = d.bv_w;
    const uint32_t first = lin[0], last = lin[W - 1];
    auto rowsum = [&](int c) -> uint32_t {
        uint32_t v = E[min(c + half + 1, W)] - E[max(c - half, 0)];
        v += (uint32_t)max(half - c, 0) * first + (uint32_t)max(c + half - (W - 1), 0) * last;
        return v;
    };
    for (int x = lane; x < d.p2; x += 32) {
        uint32_t lo = rowsum(x), hi = (x + d.p2 < W) ? rowsum(x + d.p2) : 0u;
        hrow[x] = lo | (hi << 16);
    }
}

__global__ void __launch_bounds__(32)
k_box_v(const uint32_t* __restrict__ plane_all, const uint32_t* __restrict__ hs_all, uint32_t* __restrict__ bits_all,
        LtDims d, int half, int c, int accumulate, int band_rows, size_t plane_stride, size_t bits_stride,
        const int* __restrict__ list, const int* __restrict__ count) {
    int slot = blockIdx.z;
    if (count != nullptr && slot >= *count) return;
    int s = list ? list[slot] : slot;
    int lane = threadIdx.x;
    int x = blockIdx.x * 32 + lane;
    int yb0 = blockIdx.y * band_rows, yb1 = min(yb0 + band_rows, d.bv_h);
    const uint32_t* P = plane_all + (size_t)s * plane_stride + x;
    const uint32_t* Hs = hs_all + (size_t)s * plane_stride + x;
    uint32_t* bits = bits_all + (size_t)s * bits_stride;
    const bool hi_ok = x + d.p2 < d.bv_w;
    auto ldh = [&](int r) -> uint32_t { r = max(0, min(d.bv_h - 1, r)); return __ldg(&Hs[(size_t)r * d.p2]); };
    uint32_t Sl = 0, Sh = 0;
    for (int dy = -half; dy <= half; ++dy) { uint32_t v = ldh(yb0 + dy); Sl += v & 0xFFFFu; Sh += v >> 16; }
    const uint32_t n = (uint32_t)(2 * half + 1) * (uint32_t)(2 * half + 1);
    for (int y = yb0; y < yb1; ++y) {
        uint32_t p = __ldg(&P[(size_t)y * d.p2]);
        int ml = (int)((2u * Sl + n) / (2u * n)), mh = (int)((2u * Sh + n) / (2u * n));
        bool pl = ((int)(p & 0xFFFFu) - ml) > c;
        bool ph = hi_ok && (((int)(p >> 16) - mh) > c);
        uint32_t bl = __ballot_sync(0xFFFFFFFFu, pl), bh = __ballot_sync(0xFFFFFFFFu, ph);
        if (lane == 0) {
            uint32_t* brow = bits + (size_t)y * d.mwords;
            int wl = blockIdx.x, wh = blockIdx.x + (d.p2 >> 5);
            brow[wl] = accumulate ? (brow[wl] | bl) : bl;
            brow[wh] = accumulate ? (brow[wh] | bh) : bh;
        }
        uint32_t a = ldh(y + half + 1), b = ldh(y - half);
        Sl += (a & 0xFFFFu) - (b & 0xFFFFu);
        Sh += (a >> 16) - (b >> 16);
    }
}

// mask_noise (lane_tracker.py:221-231): merged &= ~inRange(b, thresh, 255) | cross(b, k_noise, C_noise)
__global__ void __launch_bounds__(32)
k_noise_combine(const uint32_t* __restrict__ planeB_all, const uint32_t* __restrict__ noise_bits_all,
                uint32_t* __restrict__ merged_all, LtDims d, int thresh, size_t plane_stride, size_t bits_stride,
                const int* __restrict__ list, const int* __restrict__ count) {
    int slot = blockIdx.z;
    if (count != nullptr && slot >= *count) return;
    int s = list ? list[slot] : slot;
    int lane = threadIdx.x, x = blockIdx.x * 32 + lane, y = blockIdx.y;
    uint32_t p = __ldg(&planeB_all[(size_t)s * plane_stride + (size_t)y * d.p2 + x]);
    uint32_t il = __ballot_sync(0xFFFFFFFFu, (int)(p & 0xFFFFu) >= thresh);
    uint32_t ih = __ballot_sync(0xFFFFFFFFu, (int)(p >> 16) >= thresh);
    if (lane == 0) {
        size_t o = (size_t)s * bits_stride + (size_t)y * d.mwords;
        int wl = blockIdx.x, wh = blockIdx.x + (d.p2 >> 5);
        merged_all[o + wl] &= (~il | noise_bits_all[o + wl]);
        merged_all[o + wh] &= (~ih | noise_bits_all[o + wh]);
    }
}

// ---------------------------------------------------------------------------
// open 5x5 ellipse on the bit mask (lane_tracker.py:238): rows [0,2,2,2,0]
// ---------------------------------------------------------------------------

constexpr int OPEN_ROWS = 32;   // output rows per CTA

__device__ __forceinline__ uint32_t shl_bits(uint32_t prev, uint32_t cur, int n) {   // bit x <- bit x-n
    return (cur << n) | (prev >> (32 - n));
}
__device__ __forceinline__ uint32_t shr_bits(uint32_t cur, uint32_t next, int n) {   // bit x <- bit x+n
    return (cur >> n) | (next << (32 - n));
}

__global__ void __launch_bounds__(256)
k_open5(const uint32_t* __restrict__ in_all, uint32_t* __restrict__ out_all, LtDims d, size_t bits_stride,
        const int* __restrict__ list, const int* __restrict__ count) {
    int slot = blockIdx.y;
    if (count != nullptr && slot >= *count) return;
    int s = list ? list[slot] : slot;
    extern __shared__ uint32_t smem[];
    const int mw = d.mwords;
    const int y0 = blockIdx.x * OPEN_ROWS;
    const int nin = OPEN_ROWS + 8, ner = OPEN_ROWS + 4;
    uint32_t* M = smem;                 // rows y0-4 .. y0+OPEN_ROWS+3, erode padding (ones) applied
    uint32_t* Er = smem + nin * mw;     // eroded rows y0-2 .. y0+OPEN_ROWS+1, dilate padding (zeros)
    const uint32_t* in = in_all + (size_t)s * bits_stride;
    uint32_t* out = out_all + (size_t)s * bits_stride;
    for (int e = threadIdx.x; e < nin * mw; e += blockDim.x) {
        int r = e / mw, w = e - r * mw, y = y0 - 4 + r;
        uint32_t valid = (w * 32 + 32 <= d.bv_w) ? 0xFFFFFFFFu : (w * 32 >= d.bv_w ? 0u : ((1u << (d.bv_w - w * 32)) - 1u));
        uint32_t v = ((unsigned)y < (unsigned)d.bv_h) ? (__ldg(&in[(size_t)y * mw + w]) & valid) : 0xFFFFFFFFu;
        M[e] = v | ~valid;              // outside the image never blocks an erosion
    }
    __syncthreads();
    for (int e = threadIdx.x; e < ner * mw; e += blockDim.x) {
        int r = e / mw, w = e - r * mw, y = y0 - 2 + r;
        uint32_t res = 0;
        if ((unsigned)y < (unsigned)d.bv_h) {
            const uint32_t* c = M + (r + 2) * mw;     // row y in M
            res = c[w - 2 * mw] & c[w + 2 * mw];
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const uint32_t* q = c + dy * mw;
                uint32_t cur = q[w], prev = w > 0 ? q[w - 1] : 0xFFFFFFFFu, next = w + 1 < mw ? q[w + 1] : 0xFFFFFFFFu;
                res &= cur & shl_bits(prev, cur, 1) & shl_bits(prev, cur, 2) & shr_bits(cur, next, 1) & shr_bits(cur, next, 2);
            }
            uint32_t valid = (w * 32 + 32 <= d.bv_w) ? 0xFFFFFFFFu : (w * 32 >= d.bv_w ? 0u : ((1u << (d.bv_w - w * 32)) - 1u));
            res &= valid;
        }
        Er[e] = res;                    // rows outside the image: zeros (ignored by the dilation)
    }
    __syncthreads();
    for (int e = threadIdx.x; e < OPEN_ROWS * mw; e += blockDim.x) {
        int r = e / mw, w = e - r * mw, y = y0 + r;
        if (y >= d.bv_h) continue;
        const uint32_t* c = Er + (r + 2) * mw;
        uint32_t res = c[w - 2 * mw] | c[w + 2 * mw];
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const uint32_t* q = c + dy * mw;
            uint32_t cur = q[w], prev = w > 0 ? q[w - 1] : 0u, next = w + 1 < mw ? q[w + 1] : 0u;
            res |= cur | shl_bits(prev, cur, 1) | shl_bits(prev, cur, 2) | shr_bits(cur, next, 1) | shr_bits(cur, next, 2);
        }
        uint32_t valid = (w * 32 + 32 <= d.bv_w) ? 0xFFFFFFFFu : (w * 32 >= d.bv_w ? 0u : ((1u << (d.bv_w - w * 32)) - 1u));
        out[(size_t)y * mw + w] = res & valid;
    }
}

// ---------------------------------------------------------------------------
// layout converters (API boundary / tests)
// ---------------------------------------------------------------------------

__global__ void k_mask_to_u8(const uint32_t* __restrict__ bits, uint8_t* __restrict__ out, LtDims d, size_t bits_stride) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, s = blockIdx.z;
    if (x >= d.bv_w) return;
    uint32_t w = __ldg(&bits[(size_t)s * bits_stride + (size_t)y * d.mwords + (x >> 5)]);
    out[((size_t)s * d.bv_h + y) * d.bv_w + x] = ((w >> (x & 31)) & 1u) ? 255 : 0;
}

__global__ void k_u8_to_mask(const uint8_t* __restrict__ in, uint32_t* __restrict__ bits, LtDims d, size_t bits_stride) {
    int wd = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    int y = blockIdx.y, s = blockIdx.z;
    if (wd >= d.mwords) return;
    int x = wd * 32 + lane;
    bool on = x < d.bv_w && __ldg(&in[((size_t)s * d.bv_h + y) * d.bv_w + x]) != 0;
    uint32_t b = __ballot_sync(0xFFFFFFFFu, on);
    if (lane == 0) bits[(size_t)s * bits_stride + (size_t)y * d.mwords + wd] = b;
}

__global__ void k_plane_to_u8(const uint32_t* __restrict__ plane, uint8_t* __restrict__ out, LtDims d, size_t plane_stride) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, s = blockIdx.z;
    if (x >= d.bv_w) return;
    uint32_t v = __ldg(&plane[(size_t)s * plane_stride + (size_t)y * d.p2 + (x >= d.p2 ? x - d.p2 : x)]);
    out[((size_t)s * d.bv_h + y) * d.bv_w + x] = (uint8_t)((x >= d.p2 ? (v >> 16) : v) & 255u);
}

int lt_launch_mask_to_u8(lt_handle* h, const uint32_t* bits, uint8_t* d_mask, int n, cudaStream_t st) {
    dim3 g(lt_div_up(h->d.bv_w, 256), h->d.bv_h, n);
    k_mask_to_u8<<<g, 256, 0, st>>>(bits, d_mask, h->d, h->stream_mask);
    LT_LAUNCH_CHECK();
    return 0;
}
int lt_launch_u8_to_mask(lt_handle* h, const uint8_t* d_mask, uint32_t* bits, int n, cudaStream_t st) {
    dim3 g(lt_div_up(h->d.mwords, 8), h->d.bv_h, n);
    k_u8_to_mask<<<g, 256, 0, st>>>(d_mask, bits, h->d, h->stream_mask);
    LT_LAUNCH_CHECK();
    return 0;
}
int lt_launch_plane_to_u8(lt_handle* h, const uint32_t* plane, uint8_t* d_dst, int n, cudaStream_t st) {
    dim3 g(lt_div_up(h->d.bv_w, 256), h->d.bv_h, n);
    k_plane_to_u8<<<g, 256, 0, st>>>(plane, d_dst, h->d, h->stream_plane);
    LT_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------
// the whole filter for one attempt
// ---------------------------------------------------------------------------

static int launch_cross(lt_handle* h, const uint32_t* plane, uint32_t* bits, int k, int C, int accumulate, int n,
                        const int* list, const int* count, cudaStream_t st) {
    const LtDims& d = h->d;
    int wpad = (d.bv_w + 32) & ~31;
    size_t smem = (size_t)ROWK_WARPS * 2 * wpad * sizeof(uint32_t);
    static bool attr_done = false;
    if (!attr_done && smem > 48 * 1024) {
        LT_CUDA(cudaFuncSetAttribute(k_cross_h, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    dim3 gh(lt_div_up(d.bv_h, ROWK_WARPS), n);
    k_cross_h<<<gh, ROWK_WARPS * 32, smem, st>>>(plane, bits, d, k, C, accumulate, h->stream_plane, h->stream_mask,
                                                  list, count);
    LT_LAUNCH_CHECK();
    int band_rows = 64;
    dim3 gv(d.p2 / 32, lt_div_up(d.bv_h, band_rows), n);
    k_cross_v<<<gv, 32, 0, st>>>(plane, bits, d, k, C, band_rows, h->stream_plane, h->stream_mask, list, count);
    LT_LAUNCH_CHECK();
    return 0;
}

static int launch_box(lt_handle* h, const uint32_t* plane, uint32_t* hs, uint32_t* bits, int block, int c,
                      int accumulate, int n, const int* list, const int* count, cudaStream_t st) {
    const LtDims& d = h->d;
    int wpad = (d.bv_w + 32) & ~31;
    size_t smem = (size_t)ROWK_WARPS * 2 * wpad * sizeof(uint32_t);
    static bool attr_done = false;
    if (!attr_done && smem > 48 * 1024) {
        LT_CUDA(cudaFuncSetAttribute(k_box_h, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    int half = block / 2;
    dim3 gh(lt_div_up(d.bv_h, ROWK_WARPS), n);
    k_box_h<<<gh, ROWK_WARPS * 32, smem, st>>>(plane, hs, d, half, h->stream_plane, list, count);
    LT_LAUNCH_CHECK();
    int band_rows = 64;
    dim3 gv(d.p2 / 32, lt_div_up(d.bv_h, band_rows), n);
    k_box_v<<<gv, 32, 0, st>>>(plane, hs, bits, d, half, c, accumulate, band_rows, h->stream_plane, h->stream_mask,
                               list, count);
    LT_LAUNCH_CHECK();
    return 0;
}

int lt_launch_filter(lt_handle* h, int n, const LtAttemptParams& p, const int* list, const int* count,
                     cudaStream_t st) {
    const LtDims& d = h->d;
    int rc;
    if (p.filter_type == 0) {
        // bands: enough CTAs to fill 148 SMs x 2 while keeping the 2R-row warm-up per band small
        int tiles = lt_div_up(d.p2, MORPH_TW);
        int bands = 1;
        while (bands < 8 && n * tiles * bands < 296) ++bands;
        if ((rc = launch_morph<55, false, false>(h, h->planeB, h->tmpB, nullptr, n, list, count, bands, st))) return rc;
        lt_prof_mark(h, ST_ERODE55, st);
        if ((rc = launch_morph<29, false, false>(h, h->planeR, h->tmpR, nullptr, n, list, count, bands, st))) return rc;
        lt_prof_mark(h, ST_ERODE29, st);
        if ((rc = launch_morph<55, true, true>(h, h->tmpB, h->topB, h->planeB, n, list, count, bands, st))) return rc;
        lt_prof_mark(h, ST_TOPHAT55, st);
        if ((rc = launch_morph<29, true, true>(h, h->tmpR, h->topR, h->planeR, n, list, count, bands, st))) return rc;
        lt_prof_mark(h, ST_TOPHAT29, st);
        if ((rc = launch_cross(h, h->topR, h->merged, p.ksize_r, p.C_r, 0, n, list, count, st))) return rc;
        lt_prof_mark(h, ST_CROSS_R, st);
        if ((rc = launch_cross(h, h->topB, h->merged, p.ksize_b, p.C_b, 1, n, list, count, st))) return rc;
        lt_prof_mark(h, ST_CROSS_B, st);
    } else {
        if ((rc = launch_box(h, h->planeR, h->tmpR, h->merged, p.ksize_r, p.C_r, 0, n, list, count, st))) return rc;
        if ((rc = launch_box(h, h->planeB, h->tmpB, h->merged, p.ksize_b, p.C_b, 1, n, list, count, st))) return rc;
        lt_prof_mark(h, ST_BOX, st);
    }
    if (p.mask_noise) {
        if ((rc = launch_cross(h, h->planeB, h->mask, p.ksize_noise, p.C_noise, 0, n, list, count, st))) return rc;
        dim3 g(d.p2 / 32, d.bv_h, n);
        k_noise_combine<<<g, 32, 0, st>>>(h->planeB, h->mask, h->merged, d, p.noise_thresh, h->stream_plane,
                                          h->stream_mask, list, count);
        LT_LAUNCH_CHECK();
        lt_prof_mark(h, ST_NOISE, st);
    }
    size_t smem = (size_t)(2 * OPEN_ROWS + 12) * d.mwords * sizeof(uint32_t);
    dim3 go(lt_div_up(d.bv_h, OPEN_ROWS), n);
    k_open5<<<go, 256, smem, st>>>(h->merged, h->mask, d, h->stream_mask, list, count);
    LT_LAUNCH_CHECK();
    lt_prof_mark(h, ST_OPEN5, st);
    return 0;
}
