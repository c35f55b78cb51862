// Row half-widths of cv2.getStructuringElement(MORPH_ELLIPSE, (k, k)) for the two structuring elements of
// filter_lane_points (lane_tracker.py:203-204): shared by the morphology kernels.
#pragma once

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

