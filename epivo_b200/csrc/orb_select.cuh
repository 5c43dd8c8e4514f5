// KeyPointsFilter::retainBest (OpenCV features2d/src/keypoint.cpp) on packed keys, restating the two libstdc++ algorithms it
// calls -- std::nth_element (bits/stl_algo.h __introselect: median-of-three Hoare partition, heap-select after 2 lg n
// rounds, insertion sort of the last <= 3) and std::partition (bidirectional form) -- so that the kept keypoints come
// out in exactly OpenCV's order.  Two forms: the algorithms as libstdc++ writes them, for one thread (orb_retain_best: the
// readable statement, and the host tests' second opinion), and the same result computed by a block of threads
// (orb_retain_best_block, below: what orb.cu's orb_select_kernel runs, one block per image and level).  Both compile
// for the device and for the host (tests/cpp/orb_select_host.cpp checks them against the real std:: algorithms).
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

// ---- the same result with a block of threads ----------------------------------------------------------------------------
// A Hoare partition pass is data-parallel: positions the left scan has passed are never revisited, so its i-th stop is
// the i-th element of [first+1, last) that is NOT greater than the pivot (in the untouched array), the right scan's i-th
// stop is the i-th element of [first, last) from the right that the pivot is NOT greater than, pass i swaps the two while
// the left one is still left of the right one, and the pass returns min(L[t], R[t-1]) at the first t where they have met
// (the scans then stop on the elements the previous swap put there).  std::partition pairs the i-th failing element from
// the left with the i-th passing one from the right in the same way and returns first + (number passing).  So every pass
// is two ordered compactions, a search for t and |t| independent swaps; the median-of-three, the <= 3-element insertion
// sort and the (never reached) heap-select stay on one thread.  Exec supplies the threads:
//   tid(), nthreads(), sync(), compact2(m, fa, fb, A, B, &nA, &nB) -- A gets every c in [0, m) with fa(c), ascending, B
//   likewise with fb --, and imin(int*, v) (atomic minimum on a word all threads see).
struct OrbSelectShared {       // one per block, visible to all its threads
    int first, last, depth, done, nA, nB, t, kept;
};

template <class Exec>
EPV_HD inline int orb_retain_best_block(Exec& ex, uint32_t* k, int n, int n_points, uint32_t* A, uint32_t* B, OrbSelectShared* sh) {
    const int nth = n_points - 1;
    if (ex.tid() == 0) {
        sh->first = 0;
        sh->last = n;
        sh->depth = 0;
        for (int t = n; t > 1; t >>= 1) sh->depth += 2;
        sh->done = 0;
    }
    ex.sync();
    for (;;) {
        const int first = sh->first, last = sh->last;
        if (last - first <= 3 || sh->done) break;
        ex.sync();                                       // everyone has read the range before thread 0 changes it
        if (ex.tid() == 0) {
            if (sh->depth == 0) {
                orb_heap_select(k, first, nth + 1, last);
                kswap(k, first, nth);
                sh->done = 1;
            } else {
                --sh->depth;
                const int mid = first + (last - first) / 2;
                const int a = first + 1, b = mid, c = last - 1;          // __move_median_to_first(first, a, b, c)
                if (kgt(k[a], k[b])) {
                    if (kgt(k[b], k[c])) kswap(k, first, b);
                    else if (kgt(k[a], k[c])) kswap(k, first, c);
                    else kswap(k, first, a);
                } else if (kgt(k[a], k[c])) kswap(k, first, a);
                else if (kgt(k[b], k[c])) kswap(k, first, c);
                else kswap(k, first, b);
            }
        }
        ex.sync();
        if (sh->done) break;
        const uint32_t pv = k[first] >> 24;
        const int m = last - first;
        // A: stops of the left scan (first+1+c, ascending); B: stops of the right scan (last-1-c, descending)
        ex.compact2(m, [&](int c) { return c < m - 1 && (k[first + 1 + c] >> 24) <= pv; },
                    [&](int c) { return (k[last - 1 - c] >> 24) >= pv; }, A, B, &sh->nA, &sh->nB);
        const int nA = sh->nA, nB = sh->nB, lim = nA < nB ? nA : nB;
        if (ex.tid() == 0) sh->t = lim;
        ex.sync();
        for (int i = ex.tid(); i < lim; i += ex.nthreads())
            if (!(first + 1 + (int)A[i] < last - 1 - (int)B[i])) { ex.imin(&sh->t, i); break; }   // monotone: the first failure per thread suffices
        ex.sync();
        const int t = sh->t;
        for (int i = ex.tid(); i < t; i += ex.nthreads()) kswap(k, first + 1 + (int)A[i], last - 1 - (int)B[i]);
        if (ex.tid() == 0) {
            int cut = t < nA ? first + 1 + (int)A[t] : last;
            if (t >= 1) {
                const int r = last - 1 - (int)B[t - 1];
                if (r < cut) cut = r;
            }
            if (cut <= nth) sh->first = cut;
            else sh->last = cut;
        }
        ex.sync();
    }
    ex.sync();
    if (ex.tid() == 0 && !sh->done) {                                    // __insertion_sort(first, last)
        const int first = sh->first, last = sh->last;
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
    }
    ex.sync();
    // std::partition of [n_points, n) by response >= k[n_points-1]'s
    const uint32_t amb = k[n_points - 1] >> 24;
    const int m = n - n_points;
    ex.compact2(m, [&](int c) { return !((k[n_points + c] >> 24) >= amb); }, [&](int c) { return (k[n - 1 - c] >> 24) >= amb; },
                A, B, &sh->nA, &sh->nB);
    const int nA = sh->nA, nB = sh->nB, lim = nA < nB ? nA : nB;
    for (int i = ex.tid(); i < lim; i += ex.nthreads()) {
        const int l = n_points + (int)A[i], r = n - 1 - (int)B[i];
        if (!(l < r)) break;
        kswap(k, l, r);
    }
    ex.sync();
    return n_points + nB;
}

// one host "thread": what the block form computes, for tests/cpp/orb_select_host.cpp
struct OrbSelectHostExec {
    int tid() const { return 0; }
    int nthreads() const { return 1; }
    void sync() const {}
    void imin(int* p, int v) const { if (v < *p) *p = v; }
    template <class FA, class FB>
    void compact2(int m, FA fa, FB fb, uint32_t* A, uint32_t* B, int* nA, int* nB) const {
        int a = 0, b = 0;
        for (int c = 0; c < m; ++c) {
            if (fa(c)) A[a++] = (uint32_t)c;
            if (fb(c)) B[b++] = (uint32_t)c;
        }
        *nA = a;
        *nB = b;
    }
};
