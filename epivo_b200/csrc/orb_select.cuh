// KeyPointsFilter::retainBest (OpenCV features2d/src/keypoint.cpp) on packed keys, restating the two libstdc++ algorithms it
// calls -- std::nth_element (bits/stl_algo.h __introselect: median-of-three Hoare partition, heap-select after 2 lg n
// rounds, insertion sort of the last <= 3) and std::partition (bidirectional form) -- so that the kept keypoints come
// out in exactly OpenCV's order.  Compiles for the device (orb.cu runs it on one thread per image and level) and for the
// host (tests/cpp/orb_select_host.cpp checks it against the real std:: algorithms).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define EPV_HD __host__ __device__
#else
#define EPV_HD
#endif

// keys: response << 24 | candidate index.  comp(a, b) = KeypointResponseGreater: a.response > b.response.
EPV_HD inline bool kgt(uint32_t a, uint32_t b) { return (a >> 24) > (b >> 24); }
EPV_HD inline void kswap(uint32_t* k, int i, int j) {
    const uint32_t t = k[i];
    k[i] = k[j];
    k[j] = t;
}

// libstdc++ std::__adjust_heap + __push_heap on k[0..len) with comp = kgt
EPV_HD inline void orb_adjust_heap(uint32_t* k, int hole, int len, uint32_t value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (kgt(k[child], k[child - 1])) --child;
        k[hole] = k[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        k[hole] = k[child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && kgt(k[parent], value)) {
        k[hole] = k[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    k[hole] = value;
}

// libstdc++ std::__heap_select(first, middle, last, comp) on k[first..last)
EPV_HD inline void orb_heap_select(uint32_t* k, int first, int middle, int last) {
    uint32_t* h = k + first;
    const int len = middle - first;
    if (len >= 2)
        for (int parent = (len - 2) / 2;; --parent) {
            orb_adjust_heap(h, parent, len, h[parent]);
            if (parent == 0) break;
        }
    for (int i = middle; i < last; ++i)
        if (kgt(k[i], k[first])) {
            const uint32_t v = k[i];
            k[i] = k[first];
            orb_adjust_heap(h, 0, len, v);
        }
}

// libstdc++ std::nth_element (__introselect) on k[0..n), then KeyPointsFilter::retainBest's std::partition of the
// tail by response >= the n_points-th response.  Returns the number kept.  One thread.
EPV_HD inline int orb_retain_best(uint32_t* k, int n, int n_points) {
    int first = 0, last = n;
    const int nth = n_points - 1;
    int depth = 0;                                                       // 2 * std::__lg(n)
    for (int t = n; t > 1; t >>= 1) depth += 2;
    bool done = false;
    while (last - first > 3) {
        if (depth == 0) {
            orb_heap_select(k, first, nth + 1, last);
            kswap(k, first, nth);
            done = true;
            break;
        }
        --depth;
        const int mid = first + (last - first) / 2;
        const int a = first + 1, b = mid, c = last - 1;                  // __move_median_to_first(first, a, b, c)
        if (kgt(k[a], k[b])) {
            if (kgt(k[b], k[c])) kswap(k, first, b);
            else if (kgt(k[a], k[c])) kswap(k, first, c);
            else kswap(k, first, a);
        } else if (kgt(k[a], k[c])) kswap(k, first, a);
        else if (kgt(k[b], k[c])) kswap(k, first, c);
        else kswap(k, first, b);
        int lo = first + 1, hi = last;                                   // __unguarded_partition(first + 1, last, first)
        const uint32_t piv = k[first];
        for (;;) {
            while (kgt(k[lo], piv)) ++lo;
            --hi;
            while (kgt(piv, k[hi])) --hi;
            if (!(lo < hi)) break;
            kswap(k, lo, hi);
            ++lo;
        }
        if (lo <= nth) first = lo;
        else last = lo;
    }
    if (!done)                                                           // __insertion_sort(first, last)
        for (int i = first + 1; i < last; ++i) {
            const uint32_t v = k[i];
            if (kgt(v, k[first])) {
                for (int j = i; j > first; --j) k[j] = k[j - 1];
                k[first] = v;
            } else {
                int j = i;
                while (kgt(v, k[j - 1])) {
                    k[j] = k[j - 1];
                    --j;
                }
                k[j] = v;
            }
        }
    const uint32_t amb = k[n_points - 1] >> 24;                          // std::partition, bidirectional form
    int f = n_points, l = n;
    for (;;) {
        while (f != l && (k[f] >> 24) >= amb) ++f;
        if (f == l) break;
        --l;
        while (f != l && !((k[l] >> 24) >= amb)) --l;
        if (f == l) break;
        kswap(k, f, l);
        ++f;
    }
    return f;
}
